/*
 * ref_harness.cc -- C entry points over the UNMODIFIED reference sources (TEST INFRASTRUCTURE).
 *
 * Built by oracle/Makefile into oracle/_ref/libdvbs2_ref.so from the sources where they lie
 * under $(REF) (= /root/reference); nothing of the reference is copied into this repository.
 * It links the reference's own translation units
 *     lib/ldpc_decoder/ldpc_decoder_{avx2,sse41,generic}.cc, lib/gf.cc, lib/bch.cc
 * and includes lib/ldpc_decoder/*.hh, lib/dvb_*_tables.hh, lib/psk.hh.  Used to
 *   (1) pin oracle/dvbs2_oracle.c and generate tests/golden/ (tools/gen_golden.py),
 *   (2) time the reference CPU path (bench.py cpu_baseline / --impl reference).
 */
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "algorithms.hh"
#include "layered_decoder.hh"
#include "ldpc.hh"

#include "dvb_s2_tables.hh"
#include "dvb_s2x_tables.hh"
#include "dvb_t2_tables.hh"

#include "bch.h"
#include "gf.h"
#include "psk.hh"

/* the reference's ISA entry points (lib/ldpc_decoder_bb_impl.cc:34-52) */
namespace ldpc_avx2 {
void ldpc_dec_init(LDPCInterface* it);
int ldpc_dec_decode(void* buffer, int8_t* code, int trials);
} // namespace ldpc_avx2
namespace ldpc_sse41 {
void ldpc_dec_init(LDPCInterface* it);
int ldpc_dec_decode(void* buffer, int8_t* code, int trials);
} // namespace ldpc_sse41
namespace ldpc_generic {
void ldpc_dec_init(LDPCInterface* it);
int ldpc_dec_decode(void* buffer, int8_t* code, int trials);
} // namespace ldpc_generic

#define S2_B(X) X(DVB_S2_TABLE_B1) X(DVB_S2_TABLE_B2) X(DVB_S2_TABLE_B3) X(DVB_S2_TABLE_B4) \
    X(DVB_S2_TABLE_B5) X(DVB_S2_TABLE_B6) X(DVB_S2_TABLE_B7) X(DVB_S2_TABLE_B8)             \
    X(DVB_S2_TABLE_B9) X(DVB_S2_TABLE_B10) X(DVB_S2_TABLE_B11)
#define S2_C(X) X(DVB_S2_TABLE_C1) X(DVB_S2_TABLE_C2) X(DVB_S2_TABLE_C3) X(DVB_S2_TABLE_C4) \
    X(DVB_S2_TABLE_C5) X(DVB_S2_TABLE_C6) X(DVB_S2_TABLE_C7) X(DVB_S2_TABLE_C8)             \
    X(DVB_S2_TABLE_C9) X(DVB_S2_TABLE_C10)
#define S2X_B(X) X(DVB_S2X_TABLE_B1) X(DVB_S2X_TABLE_B2) X(DVB_S2X_TABLE_B3)                 \
    X(DVB_S2X_TABLE_B4) X(DVB_S2X_TABLE_B5) X(DVB_S2X_TABLE_B6) X(DVB_S2X_TABLE_B7)          \
    X(DVB_S2X_TABLE_B8) X(DVB_S2X_TABLE_B9) X(DVB_S2X_TABLE_B10) X(DVB_S2X_TABLE_B11)        \
    X(DVB_S2X_TABLE_B12) X(DVB_S2X_TABLE_B13) X(DVB_S2X_TABLE_B14) X(DVB_S2X_TABLE_B15)      \
    X(DVB_S2X_TABLE_B16) X(DVB_S2X_TABLE_B17) X(DVB_S2X_TABLE_B18) X(DVB_S2X_TABLE_B19)      \
    X(DVB_S2X_TABLE_B20) X(DVB_S2X_TABLE_B21) X(DVB_S2X_TABLE_B22) X(DVB_S2X_TABLE_B23)      \
    X(DVB_S2X_TABLE_B24)
#define S2X_C(X) X(DVB_S2X_TABLE_C1) X(DVB_S2X_TABLE_C2) X(DVB_S2X_TABLE_C3)                 \
    X(DVB_S2X_TABLE_C4) X(DVB_S2X_TABLE_C5) X(DVB_S2X_TABLE_C6) X(DVB_S2X_TABLE_C7)          \
    X(DVB_S2X_TABLE_C8) X(DVB_S2X_TABLE_C9) X(DVB_S2X_TABLE_C10)
#define T2(X) X(DVB_T2_TABLE_A3) X(DVB_T2_TABLE_B3)
#define ALL_TABLES(X) S2_B(X) S2_C(X) S2X_B(X) S2X_C(X) T2(X)

static LDPCInterface* make_table(const char* name)
{
#define X(T)                   \
    if (!strcmp(name, #T))     \
        return new LDPC<T>();
    ALL_TABLES(X)
#undef X
    return nullptr;
}

typedef SIMD<int8_t, 32> simd32_t;
typedef OffsetMinSumAlgorithm<simd32_t, NormalUpdate<simd32_t>, 2> alg32_t; /* ldpc_decoder_avx2.cc:13-21 */
typedef LDPCDecoder<simd32_t, alg32_t> dec32_t;

using gr::dvbs2rx::bch_codec;
using gr::dvbs2rx::bitset256_t;
using gr::dvbs2rx::galois_field;
using gr::dvbs2rx::gf2m_poly;

struct RefBch {
    std::unique_ptr<galois_field<uint32_t>> gf;
    std::unique_ptr<bch_codec<uint32_t, bitset256_t>> codec;
    int t;
};

extern "C" {

int ref_simd_width(int isa) { return isa == 2 ? 32 : 16; }

/* isa: 0 generic, 1 sse4.1, 2 avx2.  Returns N or -1. */
int ref_ldpc_init(const char* table, int isa)
{
    LDPCInterface* it = make_table(table);
    if (!it)
        return -1;
    int n = it->code_len();
    if (isa == 2)
        ldpc_avx2::ldpc_dec_init(it);
    else if (isa == 1)
        ldpc_sse41::ldpc_dec_init(it);
    else
        ldpc_generic::ldpc_dec_init(it);
    delete it;
    return n;
}

/* One SIMD batch through the reference entry point; code = [simd][N], in place. */
int ref_ldpc_decode(int isa, int n, int8_t* code, int trials)
{
    int simd = ref_simd_width(isa);
    void* buf = aligned_alloc(simd, (size_t)simd * n);
    int r;
    if (isa == 2)
        r = ldpc_avx2::ldpc_dec_decode(buf, code, trials);
    else if (isa == 1)
        r = ldpc_sse41::ldpc_dec_decode(buf, code, trials);
    else
        r = ldpc_generic::ldpc_dec_decode(buf, code, trials);
    free(buf);
    return r;
}

/* Throughput harness: `threads` workers, each with its own AVX2 LDPCDecoder instance (the
 * ISA entry points share one global object, lib/ldpc_decoder/ldpc_decoder_avx2.cc:21), each
 * decoding batches of 32 frames in place exactly as lib/ldpc_decoder_bb_impl.cc:406-419 does.
 * frames must be a multiple of 32.  ret[b] = return value of batch b.  Returns seconds spent
 * inside the decode calls (wall clock over the parallel region). */
double ref_ldpc_decode_mt(const char* table, int8_t* code, int frames, int trials, int threads,
                          int* ret)
{
    LDPCInterface* it = make_table(table);
    if (!it || frames % 32)
        return -1.0;
    const int n = it->code_len();
    const int batches = frames / 32;
    if (threads < 1)
        threads = 1;
    std::vector<std::unique_ptr<dec32_t>> dec(threads);
    std::vector<void*> buf(threads);
    for (int t = 0; t < threads; ++t) {
        dec[t].reset(new dec32_t());
        dec[t]->init(it);
        buf[t] = aligned_alloc(32, (size_t)32 * n);
    }
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&, t]() {
            for (int b = t; b < batches; b += threads) {
                int r = (*dec[t])(buf[t], code + (size_t)b * 32 * n, trials);
                if (ret)
                    ret[b] = r;
            }
        });
    for (auto& th : pool)
        th.join();
    auto t1 = std::chrono::steady_clock::now();
    for (int t = 0; t < threads; ++t)
        free(buf[t]);
    delete it;
    return std::chrono::duration<double>(t1 - t0).count();
}

/* ---- BCH: lib/bch_decoder_bb_impl.cc:43-71 construction ---------------------------------- */
void* ref_bch_create(uint32_t prim_poly, int t, int n)
{
    try {
        RefBch* b = new RefBch();
        b->gf.reset(new galois_field<uint32_t>(prim_poly));
        b->codec.reset(new bch_codec<uint32_t, bitset256_t>(b->gf.get(), (uint8_t)t, (uint32_t)n));
        b->t = t;
        return b;
    } catch (...) {
        return nullptr;
    }
}
void ref_bch_destroy(void* h) { delete (RefBch*)h; }
int ref_bch_k(void* h) { return ((RefBch*)h)->codec->get_k(); }
uint32_t ref_gf_alpha(void* h, uint32_t i) { return ((RefBch*)h)->gf->get_alpha_i(i); }
uint32_t ref_gf_min_poly(void* h, uint32_t i)
{
    RefBch* b = (RefBch*)h;
    return b->gf->get_min_poly(b->gf->get_alpha_i(i)).get_poly();
}
int ref_bch_genpoly(void* h, uint8_t* g, int cap)
{
    const auto& p = ((RefBch*)h)->codec->get_gen_poly();
    for (int i = 0; i <= p.degree() && i < cap; ++i)
        g[i] = p.get_poly()[i];
    return p.degree();
}
void ref_bch_encode(void* h, const uint8_t* msg, uint8_t* cw) { ((RefBch*)h)->codec->encode(msg, cw); }
/* -100: the reference threw */
int ref_bch_decode(void* h, const uint8_t* cw, uint8_t* msg)
{
    try {
        return ((RefBch*)h)->codec->decode(cw, msg);
    } catch (...) {
        return -100;
    }
}
int ref_bch_syndrome(void* h, const uint8_t* cw, uint32_t* synd)
{
    auto s = ((RefBch*)h)->codec->syndrome(cw);
    for (size_t i = 0; i < s.size(); ++i)
        synd[i] = s[i];
    return (int)s.size();
}
int ref_bch_err_loc_poly(void* h, const uint32_t* synd, uint32_t* sigma)
{
    RefBch* b = (RefBch*)h;
    std::vector<uint32_t> s(synd, synd + 2 * b->t);
    auto p = b->codec->err_loc_polynomial(s);
    for (int i = 0; i <= p.degree(); ++i)
        sigma[i] = p.get_poly()[i];
    return p.degree();
}
/* frames decoded back to back on `threads` workers (one codec is const and shared);
 * returns seconds inside decode */
double ref_bch_decode_mt(void* h, const uint8_t* cw, uint8_t* msg, int frames, int nbytes,
                         int kbytes, int threads, int* ret)
{
    RefBch* b = (RefBch*)h;
    if (threads < 1)
        threads = 1;
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&, t]() {
            for (int f = t; f < frames; f += threads) {
                int r;
                try {
                    r = b->codec->decode(cw + (size_t)f * nbytes, msg + (size_t)f * kbytes);
                } catch (...) {
                    r = -100;
                }
                if (ret)
                    ret[f] = r;
            }
        });
    for (auto& th : pool)
        th.join();
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

/* ---- 8PSK soft demap + deinterleave: lib/psk.hh:143-150 driven exactly as
 *      lib/xfecframe_demapper_cb_impl.cc:148,155-176 does -------------------------------- */
void ref_demap_8psk(const float* iq, int n_syms, float n0, int r0, int r1, int r2, int8_t* out)
{
    PhaseShiftKeying<8, gr_complex, int8_t> mod;
    Modulation<gr_complex, int8_t>* m = &mod;
    float precision = 4.0 / n0;
    std::vector<int8_t> soft((size_t)3 * n_syms);
    const gr_complex* in = reinterpret_cast<const gr_complex*>(iq);
    for (int j = 0; j < n_syms; ++j)
        m->soft(soft.data() + 3 * j, in[j], precision);
    int idx = 0;
    for (int j = 0; j < n_syms; ++j) {
        out[r0 + j] = soft[idx++];
        out[r1 + j] = soft[idx++];
        out[r2 + j] = soft[idx++];
    }
}
/* ---- 8PSK SNR estimates: lib/psk.hh hard()/map() driven as the block does.  llr == NULL: the initial
 *      estimate from sliced symbols (lib/xfecframe_demapper_cb_impl.cc:128-142); else the post-decoder one
 *      from posterior LLR signs, re-interleaved through the row offsets (:267-302).  Returns sp / np. */
float ref_snr_8psk(const float* iq, int n_syms, const int8_t* llr, int r0, int r1, int r2)
{
    PhaseShiftKeying<8, gr_complex, int8_t> mod;
    Modulation<gr_complex, int8_t>* m = &mod;
    const gr_complex* in = reinterpret_cast<const gr_complex*>(iq);
    float sp = 0, np = 0;
    int8_t tmp[3];
    for (int j = 0; j < n_syms; ++j) {
        if (llr) {
            tmp[0] = llr[r0 + j] < 0 ? -1 : 1;
            tmp[1] = llr[r1 + j] < 0 ? -1 : 1;
            tmp[2] = llr[r2 + j] < 0 ? -1 : 1;
        } else {
            m->hard(tmp, in[j]);
        }
        gr_complex s = m->map(tmp);
        gr_complex e = in[j] - s;
        sp += std::norm(s);
        np += std::norm(e);
    }
    if (!(np > 0))
        np = 1e-12;
    return sp / np;
}
/* QPSK through the generic PSK class (lib/psk.hh:87-91); the block itself uses VOLK */
void ref_demap_qpsk_psk4(const float* iq, int n_syms, float precision, int8_t* out)
{
    PhaseShiftKeying<4, gr_complex, int8_t> mod;
    const gr_complex* in = reinterpret_cast<const gr_complex*>(iq);
    for (int j = 0; j < n_syms; ++j)
        mod.soft(out + 2 * j, in[j], precision);
}

} // extern "C"
