/*
 * ref_bb_harness.cc -- C entry points over the reference's UNMODIFIED bbdescrambler_bb_impl.cc and
 * bbdeheader_bb_impl.cc (TEST INFRASTRUCTURE, part of oracle/_ref/libdvbs2_ref.so).
 *
 * GNU Radio is not installed in this image; oracle/shim/gnuradio/{block,sync_block,io_signature,
 * logger}.h stand in for the handful of gr::block members those two files use, so that their
 * work()/general_work() bodies run exactly as written, on caller-owned arrays.
 */
#include <cstdint>
#include <cstring>
#include <vector>

#include "bbdeheader_bb_impl.h"
#include "bbdescrambler_bb_impl.h"
#include "pl_descrambler.h"

using namespace gr::dvbs2rx;

extern "C" {

int ref_bb_descramble(int standard, int framesize, int rate, const uint8_t* in, int nbytes, uint8_t* out)
{
    try {
        auto blk = bbdescrambler_bb::make((dvb_standard_t)standard, (dvb_framesize_t)framesize, (dvb_code_rate_t)rate);
        gr_vector_const_void_star ins(1, in);
        gr_vector_void_star outs(1, out);
        return blk->work(nbytes, ins, outs);
    } catch (...) {
        return -1;
    }
}

void* ref_bbdeheader_create(int standard, int framesize, int rate)
{
    try {
        return new bbdeheader_bb::sptr(
            bbdeheader_bb::make((dvb_standard_t)standard, (dvb_framesize_t)framesize, (dvb_code_rate_t)rate, 0));
    } catch (...) {
        return nullptr;
    }
}
void ref_bbdeheader_destroy(void* h) { delete static_cast<bbdeheader_bb::sptr*>(h); }

/* general_work on `n_in` input bytes with room for `n_out` output bytes; returns bytes produced */
int ref_bbdeheader_work(void* h, const uint8_t* in, int n_in, uint8_t* out, int n_out)
{
    bbdeheader_bb::sptr& blk = *static_cast<bbdeheader_bb::sptr*>(h);
    gr_vector_int ninput(1, n_in);
    gr_vector_const_void_star ins(1, in);
    gr_vector_void_star outs(1, out);
    return blk->general_work(n_out, ninput, ins, outs);
}
long ref_bbdeheader_consumed(void* h) { return (*static_cast<bbdeheader_bb::sptr*>(h))->shim_consumed; }
/* lib/pl_descrambler.cc (compiled unmodified over oracle/shim/volk): the scrambling code R_n of the first n payload
 * symbols of a Gold code, read back from the descrambling sequence (1 -> 0, -j -> 1, -1 -> 2, +j -> 3) */
int ref_pl_rn(int gold_code, uint8_t* rn, int n)
{
    pl_descrambler d(gold_code);
    std::vector<gr_complex> ones(n, gr_complex(1.0f, 0.0f));
    d.descramble(ones.data(), (uint16_t)n);
    const gr_complex* y = d.get_payload();
    for (int i = 0; i < n; ++i) {
        if (y[i] == gr_complex(1, 0))
            rn[i] = 0;
        else if (y[i] == gr_complex(0, -1))
            rn[i] = 1;
        else if (y[i] == gr_complex(-1, 0))
            rn[i] = 2;
        else if (y[i] == gr_complex(0, 1))
            rn[i] = 3;
        else
            return -1;
    }
    return 0;
}
void ref_bbdeheader_counters(void* h, uint64_t* out5)
{
    bbdeheader_bb::sptr& blk = *static_cast<bbdeheader_bb::sptr*>(h);
    out5[0] = blk->get_packet_count();
    out5[1] = blk->get_error_count();
    out5[2] = blk->get_bbframe_count();
    out5[3] = blk->get_bbframe_drop_count();
    out5[4] = blk->get_bbframe_gap_count();
}
}
