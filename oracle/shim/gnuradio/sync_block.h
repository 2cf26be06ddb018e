/* Build shim for oracle/_ref (see block.h): gr::sync_block with a directly callable work(). */
#ifndef ORACLE_SHIM_GR_SYNC_BLOCK_H
#define ORACLE_SHIM_GR_SYNC_BLOCK_H
#include <gnuradio/block.h>
namespace gr {
class sync_block : public block
{
public:
    sync_block() {}
    sync_block(const std::string& n, io_signature::sptr i, io_signature::sptr o) : block(n, i, o) {}
    virtual int work(int, gr_vector_const_void_star&, gr_vector_void_star&) { return 0; }
};
} // namespace gr
#endif
