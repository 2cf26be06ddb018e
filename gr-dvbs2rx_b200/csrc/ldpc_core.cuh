// ldpc_core.cuh -- the arithmetic of the layered offset-min-sum decoder, one check-node PAIR per thread.
//
// Host/device code: the sm_100a kernel (ldpc_kernel.cu) calls these functions from its threads, and
// tools/ldpc_emul.cc calls the very same functions thread by thread on the CPU (intrinsics replaced by the
// scalar definitions below), so the arithmetic can be checked against the oracle without a GPU.
//
// Arithmetic contract (bit-exact with the reference CPU path):
//   lib/ldpc_decoder/layered_decoder.hh:50-79   one check node: v->c = L - msg, finalp, L = v->c + c->v
//   lib/ldpc_decoder/algorithms.hh:151-207      OffsetMinSumAlgorithm<int8>: beta = 1, two minima, sign product
//                                               with zero counted as +, stored messages clamped to [-32, 31]
//
// Representation (what makes the pair step short):
//   * posteriors live in shared memory as BIASED bytes u = L + 128, in the pair-interleaved order of
//     code_tables.h: nodes p and p+180 of a layer find their operands of a link in ONE 16-bit word, so a thread
//     runs both nodes in the halves of a register (s16x2: node p = low half "A", node p+180 = high half "B");
//   * saturating int8 arithmetic in the biased domain is ONE instruction: x + 128 = relu(min(u + (-old), 255))
//     (VIADDMNMX.S16x2.RELU);
//   * the c->v messages of a check node are a function of {min0, min1, argmin, one sign per link}.  State per
//     check-node pair and layer: one word of clamped minima (bytes c0A c1A c0B c1B, c = min(m, 32)) and a 2-bit
//     field (sign | argmin << 1) per link and node, 8 links per word, bytes [A links 0-3 | B 0-3 | A 4-7 | B 4-7].
//     A field IS a byte index into a 4-entry candidate table {-min(c0,31), +c0, -min(c1,31), +c1}: one PRMT with
//     sign replication selects -(old message) of both nodes, asymmetric clamp included;
//   * the same trick produces the new messages: {+min0, -min0, +min1, -min1} selected by the NEW fields, which are
//     the next iteration's state -- no per-link compare, no per-link sign mask;
//   * address and selector arithmetic of a link is four IMADs on ready-made table operands (FMA pipe), the
//     min/max/select work is on the ALU pipe: the two pipes issue side by side.
#pragma once
#include <stdint.h>

#include "code_tables.h"

#if defined(__CUDACC__)
#define LDPC_HD __host__ __device__ __forceinline__
#else
#define LDPC_HD inline
#endif

namespace dvbs2b200 {
namespace core {

constexpr int kPairs = 180; // check-node pairs per layer

// ---- primitives ------------------------------------------------------------------------------------
LDPC_HD uint32_t h2(int x) { return (uint32_t)(uint16_t)x * 0x00010001u; }

#if defined(__CUDA_ARCH__)
LDPC_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}
LDPC_HD uint32_t imad(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
LDPC_HD uint32_t mulhi(uint32_t a, uint32_t b)
{
    uint32_t r;
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
LDPC_HD uint32_t vmin2(uint32_t a, uint32_t b) { return __vmins2(a, b); }
LDPC_HD uint32_t vmax2(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }
LDPC_HD uint32_t vmax3(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_s16x2(a, b, c); }
LDPC_HD uint32_t vmin3(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_s16x2(a, b, c); }
LDPC_HD uint32_t vadd2(uint32_t a, uint32_t b) { return __vadd2(a, b); }
// relu(min(a + b, c)) per half
LDPC_HD uint32_t vaddmin_relu(uint32_t a, uint32_t b, uint32_t c) { return __viaddmin_s16x2_relu(a, b, c); }
// max(a + b, c) per half
LDPC_HD uint32_t vaddmax(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2(a, b, c); }
// relu(max(a + b, c)) per half
LDPC_HD uint32_t vaddmax_relu(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2_relu(a, b, c); }
LDPC_HD int popc(uint32_t x) { return __popc(x); }
#else
inline uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    const uint64_t pool = (uint64_t)a | ((uint64_t)b << 32);
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        const uint32_t n = (sel >> (4 * i)) & 0xfu;
        uint32_t byte = (uint32_t)((pool >> (8 * (n & 7u))) & 0xffu);
        if (n & 8u)
            byte = (byte & 0x80u) ? 0xffu : 0x00u;
        r |= byte << (8 * i);
    }
    return r;
}
inline uint32_t imad(uint32_t a, uint32_t b, uint32_t c) { return a * b + c; }
inline uint32_t mulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline int16_t lo16(uint32_t x) { return (int16_t)(x & 0xffffu); }
inline int16_t hi16(uint32_t x) { return (int16_t)(x >> 16); }
inline uint32_t pk16(int a, int b) { return ((uint32_t)a & 0xffffu) | ((uint32_t)b << 16); }
inline int imin(int a, int b) { return a < b ? a : b; }
inline int imax(int a, int b) { return a > b ? a : b; }
inline uint32_t vmin2(uint32_t a, uint32_t b) { return pk16(imin(lo16(a), lo16(b)), imin(hi16(a), hi16(b))); }
inline uint32_t vmax2(uint32_t a, uint32_t b) { return pk16(imax(lo16(a), lo16(b)), imax(hi16(a), hi16(b))); }
inline uint32_t vmax3(uint32_t a, uint32_t b, uint32_t c) { return vmax2(vmax2(a, b), c); }
inline uint32_t vmin3(uint32_t a, uint32_t b, uint32_t c) { return vmin2(vmin2(a, b), c); }
inline uint32_t vadd2(uint32_t a, uint32_t b) { return pk16((int16_t)(lo16(a) + lo16(b)), (int16_t)(hi16(a) + hi16(b))); }
inline uint32_t vaddmin_relu(uint32_t a, uint32_t b, uint32_t c) { return vmax2(vmin2(vadd2(a, b), c), 0u); }
inline uint32_t vaddmax(uint32_t a, uint32_t b, uint32_t c) { return vmax2(vadd2(a, b), c); }
inline uint32_t vaddmax_relu(uint32_t a, uint32_t b, uint32_t c) { return vmax2(vmax2(vadd2(a, b), c), 0u); }
inline int popc(uint32_t x) { return __builtin_popcount(x); }
#endif

// ---- operands of a link ----------------------------------------------------------------------------
// thread constants
struct ThreadConst {
    uint32_t p;    // check-node pair
    uint32_t two;  // 2, 4, 0xffffffff: opaque to the compiler, so that the multiply-highs and multiply-adds below stay
    uint32_t four; // on the FMA pipe instead of becoming shifts / adds on the (busier) ALU pipe
    uint32_t neg1;
    uint32_t c30, c16; // 1 << 30, 1 << 16: multiply-high by them = shift right by 2 / by 16
    uint32_t c32;      // 32 (a literal 32 would become a shift-add on the ALU pipe)
};
struct LinkOp {
    uint32_t adr;  // byte offset of the halfword in L
    uint32_t g2;   // 1: node A is the LOW byte of the halfword
};
// data link through circulant e (code_tables.h: EdgeRec)
LDPC_HD LinkOp data_link(const EdgeRec& e, const ThreadConst& tc)
{
    const uint32_t t = imad(tc.p, tc.neg1, e.w0);
    LinkOp o;
    o.g2 = mulhi(t, tc.two);             // bit 31
    const uint32_t v = mulhi(t, tc.four); // 2 * g2 + ge
    // hi + 2p - 360 ge,  ge = v - 2 g2
    o.adr = imad(v, 0xfffffe98u /* -360 */, imad(o.g2, 720u, imad(tc.p, tc.two, (uint32_t)e.hi)));
    return o;
}
// unpack selector: bytes of the halfword, zero extended, node A -> low half
LDPC_HD uint32_t sel_unpack(uint32_t g2) { return imad(g2, 0xffu, 0x4041u); } // g2 = 1: 0x4140 (A low byte), 0: 0x4041
// pack selector: low bytes of the halves back into memory order (only the low 16 bits of the result are stored)
LDPC_HD uint32_t sel_pack(uint32_t g2) { return imad(g2, 0x1eu, 0x4402u); }   // g2 = 1: 0x4420, 0: 0x4402

// ---- state ------------------------------------------------------------------------------------------
template <int NW>
struct PairState {      // decoded, ready for the links
    uint32_t candA, candB; // bytes {-min(c0,31), +c0, -min(c1,31), +c1} of node A / B
    uint32_t W[NW];        // 2-bit fields, state layout
};
LDPC_HD constexpr int field_shift(int d) { return 16 * ((d & 7) >> 2) + 2 * (d & 3); }

// word w of a register array, w not a compile-time constant (no dynamic register indexing)
template <int NW>
LDPC_HD uint32_t pick_word(const uint32_t (&W)[NW], int w)
{
    uint32_t r = W[0];
#pragma unroll
    for (int k = 1; k < NW; ++k)
        r = (w == k) ? W[k] : r;
    return r;
}
template <int NW>
LDPC_HD void or_word(uint32_t (&W)[NW], int w, uint32_t v)
{
#pragma unroll
    for (int k = 0; k < NW; ++k)
        W[k] |= (NW == 1 || w == k) ? v : 0u;
}

// candidate tables from the word of clamped minima (bytes c0A c1A c0B c1B)
LDPC_HD void decode_minima(uint32_t Cw, uint32_t& candA, uint32_t& candB)
{
    const uint32_t P0 = prmt(Cw, 0u, 0x4240u), P1 = prmt(Cw, 0u, 0x4341u); // halves [c0A, c0B], [c1A, c1B]
    const uint32_t N0 = vaddmax(~P0, h2(1), h2(-31)), N1 = vaddmax(~P1, h2(1), h2(-31)); // -min(c, 31)
    const uint32_t X = prmt(N0, P0, 0x6240u), Y = prmt(N1, P1, 0x6240u);
    candA = prmt(X, Y, 0x5410u);
    candB = prmt(X, Y, 0x7632u);
}

// The field word of link d shifted down so that the link's fields sit at bits 0-1 (node A) and 8-9 (node B).  The
// links of a word are visited in order, so each shift is one multiply-high (FMA pipe) of the previous one.
LDPC_HD uint32_t field_word(uint32_t W, uint32_t prev, int d, const ThreadConst& tc)
{
    const int r = d & 7;
    return r == 0 ? W : (r == 4 ? mulhi(W, tc.c16) : mulhi(prev, tc.c30));
}
// x + 128 of a link: u = unpacked biased posteriors, wsh = field_word() of the link
LDPC_HD uint32_t link_x(uint32_t u, uint32_t wsh, uint32_t candA, uint32_t candB)
{
    const uint32_t negold = prmt(candA, candB, imad(wsh & 0x0303u, 0x11u, 0xc480u)); // -(old message), sign extended
    return vaddmin_relu(u, negold, h2(255));
}

// running minima and sign bits over the links of a check-node pair
template <int NW>
struct Acc {
    uint32_t k0, k1;    // two smallest keys per half; key = max(|x| - 1, 0) * 32 + link
    // "x of link d is >= 0" bits, collected with LEFT shifts only (IMAD.SHL, FMA pipe): links with (d & 7) < 4 in
    // lo[], the others in hi[], both at bits 8, 10, 12, 14 of the node's half; sgn_word() merges them
    uint32_t lo[NW], hi[NW];
};
template <int NW>
LDPC_HD void acc_init(Acc<NW>& a)
{
    a.k0 = a.k1 = h2(0x7fff);
#pragma unroll
    for (int w = 0; w < NW; ++w)
        a.lo[w] = a.hi[w] = 0u;
}
// accumulator layout: bit 2*(d & 7) of a half of word d >> 3 (node A = low half)
template <int NW>
LDPC_HD uint32_t sgn_word(const Acc<NW>& a, int w) { return ((a.lo[w] >> 8) & 0x00ff00ffu) | a.hi[w]; }
// magnitude max(|x| - 1, 0) of xb = x + 128 (x = -128 gives 127; the cap at 126 is applied to the minima):
// relu(max(xb - 129, 127 - xb)); 255 - xb needs no borrow between the halves, so it is one IMAD
LDPC_HD uint32_t link_mag(uint32_t xb, const ThreadConst& tc)
{
    return vaddmax_relu(xb, h2(-129), vadd2(imad(xb, tc.neg1, 0x00ff00ffu), h2(-128)));
}
template <int NW>
LDPC_HD void link_sign(Acc<NW>& a, uint32_t xb, int d)
{
    const int pos = 8 + 2 * (d & 3); // bit 7 of each half (x >= 0) moves up by pos - 7
    const uint32_t moved = imad(xb, 1u << (pos - 7), 0u) & (0x00010001u << pos);
    or_word(a.lo, d >> 3, (d & 7) < 4 ? moved : 0u);
    or_word(a.hi, d >> 3, (d & 7) < 4 ? 0u : moved);
}
template <int NW>
LDPC_HD void link_merge(Acc<NW>& a, uint32_t xb, int d, const ThreadConst& tc)
{
    const uint32_t key = imad(link_mag(xb, tc), tc.c32, h2(d));
    a.k1 = vmin2(a.k1, vmax2(a.k0, key));
    a.k0 = vmin2(a.k0, key);
    link_sign(a, xb, d);
}

// what the links are updated with, and the state that is stored
template <int NW>
struct Final {
    uint32_t cand2A, cand2B; // bytes {+min0, -min0, +min1, -min1}
    uint32_t W[NW];          // new fields, state layout
    uint32_t Cw;             // new clamped minima
};
// deg_live: links that exist in this layer (a dead slot contributes nothing to acc)
template <int NW>
LDPC_HD void finalize(const Acc<NW>& a, int deg_live, Final<NW>& f)
{
    // order statistics commute with the monotone cap: min(|x|, 127) - 1 <= 126 is applied to the two minima
    const uint32_t min0 = vmin2((a.k0 >> 5) & 0x07ff07ffu, h2(126)), min1 = vmin2((a.k1 >> 5) & 0x07ff07ffu, h2(126));
    const uint32_t n0 = vadd2(~min0, h2(1)), n1 = vadd2(~min1, h2(1));
    const uint32_t X = prmt(min0, n0, 0x6240u), Y = prmt(min1, n1, 0x6240u);
    f.cand2A = prmt(X, Y, 0x5410u);
    f.cand2B = prmt(X, Y, 0x7632u);
    f.Cw = prmt(vmin2(min0, h2(32)), vmin2(min1, h2(32)), 0x6240u);
    // signs: new message of link d is negative iff an odd number of the OTHER links are negative (zero counts as +)
    int nnA = 0, nnB = 0;
    uint32_t sg[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        sg[w] = sgn_word(a, w);
        nnA += popc(sg[w] & 0x00005555u);
        nnB += popc(sg[w] & 0x55550000u);
    }
    const uint32_t flip = (((deg_live + 1 + nnA) & 1) ? 0x00005555u : 0u) | (((deg_live + 1 + nnB) & 1) ? 0x55550000u : 0u);
    const uint32_t argA = a.k0 & 31u, argB = (a.k0 >> 16) & 31u;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        uint32_t acc = sg[w] ^ flip;
        const uint32_t hotA = 2u << (2 * (argA & 7u)), hotB = 0x20000u << (2 * (argB & 7u));
        if (NW == 1 || (int)(argA >> 3) == w)
            acc |= hotA;
        if (NW == 1 || (int)(argB >> 3) == w)
            acc |= hotB;
        f.W[w] = prmt(acc, 0u, 0x3120u); // accumulator layout -> state layout
    }
}
// new posterior + 128 of a link; wsh = field_word() of the link in the NEW field words
template <int NW>
LDPC_HD uint32_t link_new(uint32_t xb, uint32_t wsh, const Final<NW>& f)
{
    const uint32_t msg = prmt(f.cand2A, f.cand2B, imad(wsh & 0x0303u, 0x11u, 0xc480u));
    return vaddmin_relu(xb, msg, h2(255));
}
// syndrome contribution of new posteriors ub: sign parity in bit 7 of each half (inverted: bias), zero detect in bit 15
LDPC_HD void syndrome_acc(uint32_t ub, uint32_t& syn, uint32_t& zer)
{
    syn ^= ub;
    const uint32_t z = ub ^ h2(128);
    zer |= vadd2(z, h2(-1)) & ~z; // bit 15 of a half set iff that half of z is 0
}

// ---- single check node (scalar), for the serial phase of conflict layers ---------------------------------
// -(old message) of link d of the node in half hs
template <int NW>
LDPC_HD int negold_scalar(const PairState<NW>& s, int d, int hs)
{
    const uint32_t f = (pick_word(s.W, d >> 3) >> (field_shift(d) + 8 * hs)) & 3u;
    return (int)(int8_t)((hs ? s.candB : s.candA) >> (8 * f));
}
LDPC_HD int clamp255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }
LDPC_HD int mag_scalar(int xb)
{
    const int a = xb - 129, b = 127 - xb;
    const int m = a > b ? a : b;
    return m > 0 ? m : 0;
}

// byte address of node j's operand through circulant e (pair-interleaved layout)
LDPC_HD int node_operand_addr(const EdgeRec& e, int j)
{
    int ap, ra, gbase;
    unpack_edge(e, ap, ra, gbase);
    int s = j - ap - kPairs * ra; // (j - shift) mod 360
    s += (s < 0) ? 360 : 0;
    return gbase + 2 * (s >= kPairs ? s - kPairs : s) + (s >= kPairs);
}

} // namespace core
} // namespace dvbs2b200
