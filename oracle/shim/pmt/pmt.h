/* Build shim for oracle/_ref (see gnuradio/block.h): the handful of pmt functions that
 * lib/ldpc_decoder_bb_impl.cc uses to publish its "llr_pdu" message, enough for the harness to read
 * the message back.  Not GNU Radio's pmt. */
#ifndef ORACLE_SHIM_PMT_H
#define ORACLE_SHIM_PMT_H
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace pmt {
struct node;
typedef std::shared_ptr<node> pmt_t;
struct node {
    enum kind_t { NIL, SYMBOL, LONG, UINT64, DICT, PAIR, U8VECTOR } kind = NIL;
    std::string sym;
    long l = 0;
    uint64_t u = 0;
    std::map<std::string, pmt_t> dict;
    pmt_t car, cdr;
    std::vector<uint8_t> bytes;
};
inline pmt_t mp(const char* s)
{
    auto n = std::make_shared<node>();
    n->kind = node::SYMBOL;
    n->sym = s;
    return n;
}
inline pmt_t from_long(long v)
{
    auto n = std::make_shared<node>();
    n->kind = node::LONG;
    n->l = v;
    return n;
}
inline pmt_t from_uint64(uint64_t v)
{
    auto n = std::make_shared<node>();
    n->kind = node::UINT64;
    n->u = v;
    return n;
}
inline pmt_t make_dict()
{
    auto n = std::make_shared<node>();
    n->kind = node::DICT;
    return n;
}
inline pmt_t dict_add(const pmt_t& d, const pmt_t& key, const pmt_t& value)
{
    auto n = std::make_shared<node>(*d); /* dictionaries are immutable in pmt: a new one is returned */
    n->dict[key->sym] = value;
    return n;
}
inline pmt_t cons(const pmt_t& a, const pmt_t& b)
{
    auto n = std::make_shared<node>();
    n->kind = node::PAIR;
    n->car = a;
    n->cdr = b;
    return n;
}
} // namespace pmt
#endif
