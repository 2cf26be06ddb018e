/* Build shim for oracle/_ref: GNU Radio is not installed in this image.  Just enough of
 * gr::block for the reference's bbdeheader_bb_impl.cc / bbdescrambler_bb_impl.cc to compile
 * UNMODIFIED and be driven by oracle/ref_bb_harness.cc: no scheduler, no buffers -- the
 * harness calls work()/general_work() directly on caller-owned arrays. */
#ifndef ORACLE_SHIM_GR_BLOCK_H
#define ORACLE_SHIM_GR_BLOCK_H
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include <gnuradio/io_signature.h>
#include <gnuradio/logger.h>
#include <pmt/pmt.h>

typedef std::vector<int> gr_vector_int;
typedef std::vector<const void*> gr_vector_const_void_star;
typedef std::vector<void*> gr_vector_void_star;

namespace gr {
class block
{
public:
    block() : d_logger(std::make_shared<logger>()), d_debug_logger(d_logger) {}
    block(const std::string&, io_signature::sptr, io_signature::sptr) : d_logger(std::make_shared<logger>()), d_debug_logger(d_logger) {}
    virtual ~block() {}
    virtual void forecast(int, gr_vector_int&) {}
    virtual int general_work(int, gr_vector_int&, gr_vector_const_void_star&, gr_vector_void_star&) { return 0; }
    void set_output_multiple(int m) { shim_output_multiple = m; }
    void set_relative_rate(double) {}
    void consume_each(int n) { shim_consumed += n; }
    void message_port_register_out(const pmt::pmt_t&) {}
    void message_port_pub(const pmt::pmt_t&, const pmt::pmt_t& msg) { shim_msgs.push_back(msg); } // kept for the harness
    int shim_output_multiple = 1;
    long shim_consumed = 0;
    std::vector<pmt::pmt_t> shim_msgs;

protected:
    std::shared_ptr<logger> d_logger, d_debug_logger;
};
} // namespace gr

namespace gnuradio {
template <class T>
std::shared_ptr<T> get_initial_sptr(T* p)
{
    return std::shared_ptr<T>(p);
}
} // namespace gnuradio
#endif
