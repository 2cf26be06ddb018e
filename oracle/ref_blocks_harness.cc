/*
 * ref_blocks_harness.cc -- C entry points over the reference's ldpc_decoder_bb_impl.cc and bch_decoder_bb_impl.cc
 * (TEST INFRASTRUCTURE).  Compiled twice by oracle/Makefile:
 *   oracle/_ref/libref_blocks.so      the two files UNMODIFIED (CPU: AVX2 / SSE4.1 / generic decoder, bch_codec)
 *   oracle/_ref/libpatched_blocks.so  the same two files with patches/0001-ldpc-cuda-seam.patch and
 *                                     patches/0002-bch-cuda.patch applied and -DDVBS2RX_WITH_B200, linked against
 *                                     libdvbs2_b200.so: the reference's own constructor and general_work bodies run,
 *                                     with `decode == &ldpc_cuda::ldpc_dec_decode` and dvbs2b200_bch_decode inside.
 * GNU Radio, pmt and cpu_features are not installed in this image; oracle/shim/ stands in for the few members the
 * two files use (gr::block, io_signature, logger, pmt dictionaries / pairs, pdu::make_pdu_vector, GetX86Info).
 * tests/test_gpu_refblocks.py drives both libraries with the same input and compares every output.
 */
#include <cstdint>
#include <cstring>
#include <vector>

#include "bch_decoder_bb_impl.h"
#include "ldpc_decoder_bb_impl.h"

using namespace gr::dvbs2rx;

extern "C" {

void* blk_ldpc_create(int standard, int framesize, int rate, int constellation, int outputmode, int infomode, int max_trials)
{
    try {
        return new ldpc_decoder_bb::sptr(ldpc_decoder_bb::make((dvb_standard_t)standard, (dvb_framesize_t)framesize, (dvb_code_rate_t)rate,
                                                               (dvb_constellation_t)constellation, (dvb_outputmode_t)outputmode,
                                                               (dvb_infomode_t)infomode, max_trials, 0));
    } catch (...) {
        return nullptr;
    }
}
void blk_ldpc_destroy(void* h) { delete static_cast<ldpc_decoder_bb::sptr*>(h); }
int blk_ldpc_output_multiple(void* h) { return (*static_cast<ldpc_decoder_bb::sptr*>(h))->shim_output_multiple; }
/* general_work producing noutput_items bytes; returns its return value */
int blk_ldpc_work(void* h, const int8_t* in, int ninput_items, int noutput_items, uint8_t* out)
{
    ldpc_decoder_bb::sptr& blk = *static_cast<ldpc_decoder_bb::sptr*>(h);
    gr_vector_int ninput(1, ninput_items);
    gr_vector_const_void_star ins(1, in);
    gr_vector_void_star outs(1, out);
    return blk->general_work(noutput_items, ninput, ins, outs);
}
long blk_ldpc_consumed(void* h) { return (*static_cast<ldpc_decoder_bb::sptr*>(h))->shim_consumed; }
unsigned blk_ldpc_average_trials(void* h) { return (*static_cast<ldpc_decoder_bb::sptr*>(h))->get_average_trials(); }
int blk_ldpc_pdu_count(void* h) { return (int)(*static_cast<ldpc_decoder_bb::sptr*>(h))->shim_msgs.size(); }
/* message idx of the "llr_pdu" port: meta {simd_size, frame_cnt} and the payload; returns the payload size */
long blk_ldpc_pdu(void* h, int idx, uint8_t* payload, long cap, long* simd_size, uint64_t* frame_cnt)
{
    ldpc_decoder_bb::sptr& blk = *static_cast<ldpc_decoder_bb::sptr*>(h);
    if (idx < 0 || idx >= (int)blk->shim_msgs.size())
        return -1;
    const pmt::pmt_t& m = blk->shim_msgs[idx];
    *simd_size = m->car->dict.at("simd_size")->l;
    *frame_cnt = m->car->dict.at("frame_cnt")->u;
    const std::vector<uint8_t>& b = m->cdr->bytes;
    if ((long)b.size() <= cap)
        memcpy(payload, b.data(), b.size());
    return (long)b.size();
}

void* blk_bch_create(int standard, int framesize, int rate, int outputmode)
{
    try {
        return new bch_decoder_bb::sptr(
            bch_decoder_bb::make((dvb_standard_t)standard, (dvb_framesize_t)framesize, (dvb_code_rate_t)rate, (dvb_outputmode_t)outputmode, 0));
    } catch (...) {
        return nullptr;
    }
}
void blk_bch_destroy(void* h) { delete static_cast<bch_decoder_bb::sptr*>(h); }
int blk_bch_output_multiple(void* h) { return (*static_cast<bch_decoder_bb::sptr*>(h))->shim_output_multiple; }
int blk_bch_work(void* h, const uint8_t* in, int ninput_items, int noutput_items, uint8_t* out)
{
    bch_decoder_bb::sptr& blk = *static_cast<bch_decoder_bb::sptr*>(h);
    gr_vector_int ninput(1, ninput_items);
    gr_vector_const_void_star ins(1, in);
    gr_vector_void_star outs(1, out);
    return blk->general_work(noutput_items, ninput, ins, outs);
}
long blk_bch_consumed(void* h) { return (*static_cast<bch_decoder_bb::sptr*>(h))->shim_consumed; }
uint64_t blk_bch_frame_count(void* h) { return (*static_cast<bch_decoder_bb::sptr*>(h))->get_frame_count(); }
uint64_t blk_bch_error_count(void* h) { return (*static_cast<bch_decoder_bb::sptr*>(h))->get_error_count(); }
}
