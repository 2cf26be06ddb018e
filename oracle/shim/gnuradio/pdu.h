/* Build shim for oracle/_ref (see block.h): gr::pdu::make_pdu_vector for byte vectors. */
#ifndef ORACLE_SHIM_GR_PDU_H
#define ORACLE_SHIM_GR_PDU_H
#include <pmt/pmt.h>
namespace gr {
namespace types {
enum vector_type { byte_t, short_t, int_t, float_t, complex_t };
}
namespace pdu {
inline pmt::pmt_t make_pdu_vector(types::vector_type, const uint8_t* buf, size_t items)
{
    auto n = std::make_shared<pmt::node>();
    n->kind = pmt::node::U8VECTOR;
    n->bytes.assign(buf, buf + items);
    return n;
}
} // namespace pdu
} // namespace gr
#endif
