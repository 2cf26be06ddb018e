// dvbs2rx_b200_blocks.h -- host-side mirror of the three gr-dvbs2rx blocks on the FEC decode path,
// written above the C ABI of libdvbs2_b200.so.
//
// GNU Radio is not available in this build environment, so these classes do not derive from
// gr::block; they keep the reference's class names, factory signatures, stream contracts
// (item sizes, output multiples, relative rates, forecast), general_work argument meaning and
// return values, counters and error behaviour, so that the body of each reference *_impl.cc can
// be replaced by a call into them (INTEGRATION.md shows the patch):
//   ldpc_decoder_bb        <- include/gnuradio/dvbs2rx/ldpc_decoder_bb.h:38-51, lib/ldpc_decoder_bb_impl.{h,cc}
//   bch_decoder_bb         <- include/gnuradio/dvbs2rx/bch_decoder_bb.h:39-55,  lib/bch_decoder_bb_impl.{h,cc}
//   xfecframe_demapper_cb  <- include/gnuradio/dvbs2rx/xfecframe_demapper_cb.h:36-44, lib/xfecframe_demapper_cb_impl.{h,cc}
//   bbdescrambler_bb       <- include/gnuradio/dvbs2rx/bbdescrambler_bb.h, lib/bbdescrambler_bb_impl.{h,cc}
//   bbdeheader_bb          <- include/gnuradio/dvbs2rx/bbdeheader_bb.h,    lib/bbdeheader_bb_impl.{h,cc}
// (apsk_points(): the EN 302 307-1 16APSK / 32APSK constellations, the C++ twin of dvbs2rx_b200/apsk.py)
// and ldpc_cuda::ldpc_dec_init / ldpc_dec_decode, the pair that slots in beside the reference's ISA
// namespaces (lib/ldpc_decoder_bb_impl.cc:34-52) behind `int (*decode)(void*, int8_t*, int)`.
#pragma once
#include <array>
#include <complex>
#include <cstdint>
#include <functional>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <vector>

#include "dvb_config.h"

struct dvbs2b200_code;

typedef std::complex<float> gr_complex;
typedef std::vector<int> gr_vector_int;
typedef std::vector<const void*> gr_vector_const_void_star;
typedef std::vector<void*> gr_vector_void_star;

namespace gr {
namespace dvbs2rx {

// Frames handed to the GPU per launch by the LDPC block.  The reference's d_simd_size is 32 (AVX2);
// a B200 wants thousands of frames in flight, so the block asks the scheduler for a multiple of the
// reference's batch and keeps the reference's batch semantics inside it (term_group = 32).
constexpr int kRefSimdSize = 32;

// CUDA device the blocks created from now on live on (default: DVBS2RX_B200_DEVICE from the environment, else 0).
// A flowgraph that wants several GPUs creates one chain of blocks per device, or calls dvbs2b200_multi_* directly.
void set_device(int device);

// What the reference publishes on "llr_pdu" (lib/ldpc_decoder_bb_impl.cc:362-367,422-429):
// meta {simd_size, frame_cnt} + posterior LLRs of one SIMD batch.
struct llr_pdu {
    long simd_size;
    uint64_t frame_cnt;
    const int8_t* llr; // [simd_size][n_ldpc]
    size_t n_llr;
};

class ldpc_decoder_bb
{
public:
    typedef std::shared_ptr<ldpc_decoder_bb> sptr;
    static sptr make(dvb_standard_t standard, dvb_framesize_t framesize, dvb_code_rate_t rate,
                     dvb_constellation_t constellation, dvb_outputmode_t outputmode, dvb_infomode_t infomode,
                     int max_trials, int debug_level = 0);
    ~ldpc_decoder_bb();

    void forecast(int noutput_items, gr_vector_int& ninput_items_required);
    int general_work(int noutput_items, gr_vector_int& ninput_items, gr_vector_const_void_star& input_items,
                     gr_vector_void_star& output_items);
    unsigned int get_average_trials() { return d_total_trials / d_batch_cnt; } // per batch, as the reference

    // stream contract (what the reference passes to set_output_multiple / set_relative_rate)
    int output_multiple() const { return d_output_multiple; }
    double relative_rate() const { return d_relative_rate; }
    int consumed() const { return d_consumed; } // what general_work passed to consume_each
    // out-port "llr_pdu": called once per 32-frame batch, in order
    void set_llr_pdu_handler(std::function<void(const llr_pdu&)> h) { d_pdu_handler = std::move(h); }
    // frames per launch = batches_per_call * 32; 1 reproduces the reference's granularity
    void set_batches_per_call(int n);

private:
    ldpc_decoder_bb() {}
    dvbs2b200_code* d_code = nullptr;
    unsigned int d_nldpc, d_nldpc_bytes, d_kldpc, d_kldpc_bytes, d_output_mode;
    uint64_t d_frame_cnt = 0, d_batch_cnt = 0;
    unsigned int d_total_trials = 0;
    int d_max_trials;
    int d_simd_size = kRefSimdSize;
    int d_output_multiple = 0, d_consumed = 0;
    double d_relative_rate = 0;
    std::vector<int8_t> d_post;
    std::vector<int32_t> d_ret;
    std::function<void(const llr_pdu&)> d_pdu_handler;
};

class bch_decoder_bb
{
public:
    typedef std::shared_ptr<bch_decoder_bb> sptr;
    static sptr make(dvb_standard_t standard, dvb_framesize_t framesize, dvb_code_rate_t rate,
                     dvb_outputmode_t outputmode, int debug_level = 0);
    ~bch_decoder_bb();
    void forecast(int noutput_items, gr_vector_int& ninput_items_required);
    int general_work(int noutput_items, gr_vector_int& ninput_items, gr_vector_const_void_star& input_items,
                     gr_vector_void_star& output_items);
    uint64_t get_frame_count() { return d_frame_cnt; }
    uint64_t get_error_count() { return d_frame_error_cnt; }
    int output_multiple() const { return (int)d_k_bytes; }
    int consumed() const { return d_consumed; }

private:
    bch_decoder_bb() {}
    dvbs2b200_code* d_code = nullptr;
    unsigned int d_k_bytes, d_n_bytes;
    uint64_t d_frame_cnt = 0, d_frame_error_cnt = 0;
    int d_consumed = 0;
    std::vector<int32_t> d_corr;
};

class xfecframe_demapper_cb
{
public:
    typedef std::shared_ptr<xfecframe_demapper_cb> sptr;
    // throws std::runtime_error("Unsupported constellation") like lib/xfecframe_demapper_cb_impl.cc:70-72 for anything
    // but QPSK / 8PSK -- unless apsk_opt_in is set: then MOD_16APSK / MOD_32APSK are demapped by the table-driven
    // max-log demapper (dvbs2b200_demap_table; EN 302 307-1 constellations of the code rate, tables UNPINNED by any
    // reference).  The block has no SNR estimator for them: the noise level comes from set_es_n0_db().
    static sptr make(dvb_framesize_t framesize, dvb_code_rate_t rate, dvb_constellation_t constellation, bool apsk_opt_in = false);
    void set_es_n0_db(float es_n0_db);
    ~xfecframe_demapper_cb();
    void forecast(int noutput_items, gr_vector_int& ninput_items_required);
    int general_work(int noutput_items, gr_vector_int& ninput_items, gr_vector_const_void_star& input_items,
                     gr_vector_void_star& output_items);
    // in-port "llr_pdu" (lib/xfecframe_demapper_cb_impl.cc:188-318)
    void handle_llr_pdu(const llr_pdu& pdu);
    float get_snr() { return d_snr; }
    int output_multiple() const { return (int)d_fecframe_len; }
    int consumed() const { return d_consumed; }

private:
    xfecframe_demapper_cb() {}
    dvbs2b200_code* d_code = nullptr;
    dvb_constellation_t d_constellation;
    dvb_code_rate_t d_rate;
    bool d_waiting_first_llr = true;
    unsigned int d_fecframe_len, d_xfecframe_len, d_bits;
    unsigned int d_rowaddr0 = 0, d_rowaddr1 = 0, d_rowaddr2 = 0;
    float d_snr = 0, d_N0 = 1, d_precision = 4;
    uint64_t d_frame_cnt = 0;
    int d_consumed = 0;
    std::mutex d_mutex;
    static constexpr size_t kPool = 64; // XFECFRAME_POOL_SIZE
    std::array<std::vector<gr_complex>, kPool> d_pool;
    std::array<uint64_t, kPool> d_saved;
    size_t d_idx = 0;
    std::vector<float> d_n0_per_frame, d_snr_per_frame;
    bool d_table = false;            // APSK (opt-in): table-driven demapper
    std::vector<float> d_points;     // [2^bits][2]
    std::vector<int> d_row_offsets;  // [bits]
    std::vector<gr_complex> d_gather_iq;
    std::vector<int8_t> d_gather_llr;
};

// bbdescrambler_bb <- include/gnuradio/dvbs2rx/bbdescrambler_bb.h, lib/bbdescrambler_bb_impl.{h,cc} (a gr::sync_block)
class bbdescrambler_bb
{
public:
    typedef std::shared_ptr<bbdescrambler_bb> sptr;
    static sptr make(dvb_standard_t standard, dvb_framesize_t framesize, dvb_code_rate_t rate);
    ~bbdescrambler_bb();
    int work(int noutput_items, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items);
    int output_multiple() const { return (int)kbch_bytes; }

private:
    bbdescrambler_bb() {}
    dvbs2b200_code* d_code = nullptr;
    unsigned int kbch_bytes;
};

// bbdeheader_bb <- include/gnuradio/dvbs2rx/bbdeheader_bb.h:27-72, lib/bbdeheader_bb_impl.{h,cc}.
// The stream state (synchronised?, partial TS packet) lives in the handle's device memory.
class bbdeheader_bb
{
public:
    typedef std::shared_ptr<bbdeheader_bb> sptr;
    static sptr make(dvb_standard_t standard, dvb_framesize_t framesize, dvb_code_rate_t rate, int debug_level = 0);
    ~bbdeheader_bb();
    void forecast(int noutput_items, gr_vector_int& ninput_items_required);
    int general_work(int noutput_items, gr_vector_int& ninput_items, gr_vector_const_void_star& input_items,
                     gr_vector_void_star& output_items);
    uint64_t get_packet_count() { return counters(0); }
    uint64_t get_error_count() { return counters(1); }
    uint64_t get_bbframe_count() { return counters(2); }
    uint64_t get_bbframe_drop_count() { return counters(3); }
    uint64_t get_bbframe_gap_count() { return counters(4); }
    int output_multiple() const { return (int)(d_max_dfl / 8); }
    int consumed() const { return d_consumed; }
    // input is BCH output that has not been through bbdescrambler_bb: descramble on the fly (one block less)
    void set_scrambled_input(bool on) { d_scrambled = on; }

private:
    bbdeheader_bb() {}
    uint64_t counters(int which);
    dvbs2b200_code* d_code = nullptr;
    unsigned int d_kbch_bytes, d_max_dfl;
    int d_consumed = 0;
    bool d_scrambled = false;
};

} // namespace dvbs2rx
} // namespace gr

// The seam of lib/ldpc_decoder_bb_impl.h:41: `code` is [simd][N] int8, overwritten with posterior
// LLRs; returns trials left (>= 0) or -1; `buffer` (the ISA paths' scratch) is unused.
// 16APSK / 32APSK points [2^bits][2] of a code rate (index = label, first bit = MSB, unit energy); false if the
// standard defines no such MODCOD
bool apsk_points(gr::dvbs2rx::dvb_constellation_t constellation, gr::dvbs2rx::dvb_code_rate_t rate, std::vector<float>& points);

namespace ldpc_cuda {
int ldpc_dec_init(int standard, int framesize, int rate, int simd_size);
int ldpc_dec_decode(void* buffer, int8_t* code, int trials);
void ldpc_dec_shutdown();
} // namespace ldpc_cuda
