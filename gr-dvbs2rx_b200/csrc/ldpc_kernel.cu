// ldpc_kernel.cu -- layered offset-min-sum LDPC decoder for DVB-S2/S2X/T2 on sm_100a.
//
// Arithmetic contract (bit-exact with the reference CPU path):
//   lib/ldpc_decoder/layered_decoder.hh:27-160  schedule, iteration control
//   lib/ldpc_decoder/algorithms.hh:151-207      OffsetMinSumAlgorithm<int8>, beta = 1,
//                                               stored messages clamped to [-32, 31]
//   lib/ldpc_decoder_bb_impl.cc:432-442         hard decision + MSB-first packing
//
// Design (B200-first, not a translation of the SIMD-across-frames CPU code):
//   * one FECFRAME per CTA of 192 threads, three CTAs per SM, persistent over the batch;
//   * the frame's N int8 posteriors stay in shared memory for the whole decode, in a
//     "pair-interleaved" order (code_tables.h) chosen so that check nodes p and p+180 of a layer
//     read/write ONE aligned 16-bit word per link: each thread runs TWO check nodes in the
//     halves of a 32-bit register with the native s16x2 min/max/add instructions
//     (VIMNMX.S16x2, VIADD.16x2) -- the 360-bit quasi-cyclic rotation costs one PRMT;
//   * check->variable messages are never stored per edge.  A check node's `deg` int8 messages
//     are a function of {min0, min1, argmin, one sign bit per link}: one 32-bit word (64-bit
//     for deg > 15) per check node, lossless w.r.t. the reference's clamp-on-store.  These words
//     live in a per-CTA global scratch that stays L2 resident (57 MB for 444 CTAs at rate 1/2),
//     are read with one coalesced 8-byte load per thread per layer, prefetched a layer ahead;
//   * HBM traffic is the compulsory one: the soft input is read once, the packed hard
//     decisions (and optionally the posteriors) are written once;
//   * the code's tables (layers, circulants with precomputed PRMT selectors, step list) are
//     staged into shared memory with one TMA bulk copy per CTA;
//   * layers whose circulants share a 360-bit group are order sensitive in the reference (it
//     visits check nodes serially), so they run as precomputed wavefront steps of single check
//     nodes (code_tables.cc:build_schedule) -- same result, bit for bit;
//   * the syndrome test after an iteration is skipped when the check nodes of the final step,
//     whose posteriors are final by then, already prove the frame is still bad.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "kernels.h"

// This file is compiled twice (__graft_entry__.py): DVBS2_LEGACY_WAVEFRONT=0 builds the kernels whose
// order-sensitive layers run as "split" steps, =1 the ones that run them as wavefront steps of whole check
// nodes (with the tensor-memory state variant).  One binary with both paths inlined costs the hot
// conflict-free pair step registers and instruction-cache hits, so they are separate kernels and the host
// picks per code (code_tables.cc:choose_split).
#ifndef DVBS2_LEGACY_WAVEFRONT
#define DVBS2_LEGACY_WAVEFRONT 0
#endif
// Split build: kernels for codes with at least this many data links per check node are compiled for two
// resident CTAs per SM (168 registers, no spills) instead of three (96 registers, spilling)
#ifndef DVBS2_TWO_CTA_FROM
#define DVBS2_TWO_CTA_FROM 11
#endif
#if DVBS2_LEGACY_WAVEFRONT
#define LDPC_SYM(x) x##_wavefront
#else
#define LDPC_SYM(x) x##_split
#endif

namespace dvbs2b200 {

namespace {

constexpr int kPairs = 180; // check-node pairs per layer
constexpr int kTmemColsDev = 128; // = code_tables.h kTmemCols

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- s16x2 helpers: two check nodes per register ----------------------------------------------
__device__ __forceinline__ uint32_t h2(int x) { return (uint32_t)(uint16_t)x * 0x00010001u; }
__device__ __forceinline__ uint32_t vmin(uint32_t a, uint32_t b) { return __vmins2(a, b); }
__device__ __forceinline__ uint32_t vmax(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }
__device__ __forceinline__ uint32_t vadd(uint32_t a, uint32_t b) { return __vadd2(a, b); }
__device__ __forceinline__ uint32_t vsub(uint32_t a, uint32_t b) { return __vsub2(a, b); }
// PRMT with the full 4-bit selectors: bit 3 of a nibble replicates the selected byte's sign
// (__byte_perm masks the selector with 0x7777, so it cannot express the sign-extending forms)
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}
// 0xFFFF in every half whose bit 15 is set
__device__ __forceinline__ uint32_t signmask(uint32_t x) { return prmt(x, 0, 0xBB99); }
// bitwise select: mask ? a : b  (one LOP3)
__device__ __forceinline__ uint32_t bsel(uint32_t mask, uint32_t a, uint32_t b) { return (a & mask) | (b & ~mask); }
// a * b + c on the FMA pipe (IMAD), keeps shifts/adds off the ALU pipe
__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

// L2 eviction policies: the check-node state is re-read every iteration and must stay L2
// resident (evict_last); the soft input streams through once (evict_first).
__device__ __forceinline__ uint64_t policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint32_t ldg_hint(const uint32_t* p, uint64_t pol)
{
    uint32_t r;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ uint2 ldg_hint(const uint2* p, uint64_t pol)
{
    uint2 r;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(r.x), "=r"(r.y) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ uint4 ldg_hint(const uint4* p, uint64_t pol)
{
    uint4 r;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void stg_hint(uint32_t* p, uint32_t v, uint64_t pol)
{
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void stg_hint(uint2* p, uint2 v, uint64_t pol)
{
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.u32 [%0], {%1,%2}, %3;" ::"l"(p), "r"(v.x), "r"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ void stg_hint(uint4* p, uint4 v, uint64_t pol)
{
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w), "l"(pol)
                 : "memory");
}

// Tensor memory used as plain scratch (SASS: LDTM / STTM): one 32-bit cell per lane and column; a
// warp reaches the 32 TMEM lanes of its quarter (warp id % 4).  ~30 cycles for a load + store pair
// against ~700 for the L2 round trip (tools/ubench/tmem_test.cu).
__device__ __forceinline__ uint32_t tmem_ld(uint32_t taddr)
{
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    return r;
}
__device__ __forceinline__ void tmem_st(uint32_t taddr, uint32_t v)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

struct LayerView {
    uint32_t edge_begin;
    int cnt;
    int nshared; // links into groups that carry another circulant of the layer (sorted last)
};
__device__ __forceinline__ LayerView load_layer(const uint2* layers, int i)
{
    const uint2 r = layers[i];
    LayerView v;
    v.edge_begin = r.x;
    v.cnt = (int)(r.y & 0xffffu);
    v.nshared = (int)(r.y >> 16);
    return v;
}

// byte address of element s of a 360-bit group / of parity bit c in the pair-interleaved layout
__device__ __forceinline__ int data_addr(int group_base, int s) { return group_base + 2 * (s >= kPairs ? s - kPairs : s) + (s >= kPairs); }
__device__ __forceinline__ int parity_addr(int K, int half, int c) { return K + 2 * (c >= half ? c - half : c) + (c >= half); }

// Compressed check-node state word.
//   bits 0..5 min(min0, 32)   bits 6..11 min(min1, 32)   bits 12..16 argmin link
//   !WIDE: sign bit of link d at bit 17 + d (deg <= 15);  WIDE: sign bits in a second word.
// Links: 0 = own parity bit, 1 = previous parity bit of the zig-zag, 2.. = data bits.

// ---- two check nodes (p, p + 180) of a conflict-free layer ---------------------------------------
// lib/ldpc_decoder/layered_decoder.hh:57-76 + algorithms.hh:170-206, both nodes at once.
template <int CNT_MAX, bool UNIFORM, bool WIDE, bool SELF_CHECK>
__device__ __forceinline__ int process_pair(int8_t* __restrict__ L, const uint2* __restrict__ edges, const LayerView& lv, int layer,
                                            int p, int K, int q, uint32_t wA, uint32_t wB, uint32_t sA, uint32_t sB,
                                            uint32_t* __restrict__ msg_out, uint64_t pol)
{
    constexpr int DEG_MAX = CNT_MAX + 2;
    const bool first = (layer == 0 && p == 0); // check 0 has no previous parity link; check 180*q does
    const int half = kPairs * q;
    // ---- decode the stored state into s16x2 ----
    const uint32_t m0 = (wA & 63u) | ((wB & 63u) << 16);
    const uint32_t m1 = ((wA >> 6) & 63u) | (((wB >> 6) & 63u) << 16);
    const uint32_t x01 = m0 ^ m1;
    const uint32_t argA = (wA >> 12) & 31u, argB = (wB >> 12) & 31u;
    // one-hot argmin and inverted sign bits: link d of node A at bit d, of node B at bit 16 + d
    uint32_t hot_lo, hot_hi = 0, nsg_lo, nsg_hi = 0;
    if (!WIDE) {
        hot_lo = (1u << argA) | (0x10000u << argB);
        nsg_lo = ~((wA >> 17) | ((wB >> 17) << 16));
    } else {
        hot_lo = ((1u << argA) & 0xffffu) | (((1u << argB) & 0xffffu) << 16);
        hot_hi = ((1u << argA) >> 16) | ((1u << argB) & 0xffff0000u);
        nsg_lo = ~((sA & 0xffffu) | (sB << 16));
        nsg_hi = ~((sA >> 16) | (sB & 0xffff0000u));
    }

    int adr[DEG_MAX];
    uint32_t sel[DEG_MAX];
    uint32_t v[DEG_MAX], mag[DEG_MAX];
    const int c = q * p + layer;
    adr[0] = K + 2 * c;
    sel[0] = 0x9180u | (0x4420u << 16);
    // the word of parity bit 180q - 1 holds node B's previous parity bit in its LOW byte
    adr[1] = first ? K + 2 * (half - 1) : K + 2 * c - 2;
    sel[1] = first ? (0x8091u | (0x4402u << 16)) : sel[0];
    const uint32_t pkey = ((uint32_t)p << 16) | 0xffffu; // e.x > pkey  <=>  p < a'
#pragma unroll
    for (int d = 0; d < CNT_MAX; ++d) {
        if (UNIFORM || d < lv.cnt) {
            const uint2 e = edges[lv.edge_begin + d];
            const bool lt = e.x > pkey;
            adr[d + 2] = (int)(e.x & 0xffffu) + 2 * p - (lt ? 0 : 360);
            sel[d + 2] = lt ? (e.y ^ 0x00221111u) : e.y;
        } else {
            adr[d + 2] = 0;
            sel[d + 2] = 0;
        }
    }

    // candidates for -(old message): old >= 0 -> -min(m, 31), old < 0 -> +m; the argmin link takes m1
    const uint32_t n0p = vsub(0u, vmin(m0, h2(31))), n1p = vsub(0u, vmin(m1, h2(31)));
    const uint32_t nx01p = n0p ^ n1p;
    uint32_t k0 = h2(0x7fff), k1 = h2(0x7fff), sx = 0;
#pragma unroll
    for (int d = 0; d < DEG_MAX; ++d) {
        const bool live = UNIFORM || (d < 2) || (d - 2 < lv.cnt);
        if (live) {
            const uint32_t raw = *reinterpret_cast<const uint16_t*>(L + adr[d]);
            const uint32_t l = prmt(raw, 0, sel[d]); // PRMT reads selector bits 15:0 only
            // stored message: +-(d == argmin ? min1 : min0), clamped to [-32, 31]
            const uint32_t hot = (!WIDE || d < 16) ? hot_lo : hot_hi;
            const uint32_t nsg = (!WIDE || d < 16) ? nsg_lo : nsg_hi;
            const uint32_t im = signmask(imad(hot, 1u << (15 - (d & 15)), 0u)); // the shift as IMAD: FMA pipe, not ALU
            const uint32_t nm = signmask(imad(nsg, 1u << (15 - (d & 15)), 0u)); // 0xFFFF where the old message was >= 0
            const uint32_t xa = vadd(l, n0p ^ (nx01p & im));      // l - old, old >= 0
            const uint32_t xb = vadd(l, m0 ^ (x01 & im));         // l - old, old < 0
            uint32_t x = vmin(vmax(bsel(nm, xa, xb), h2(-128)), h2(127)); // vqsub
            if (d == 1 && first)
                x &= 0xffff0000u; // node A (check 0) has no such link: neutral value
            v[d] = x;
            sx ^= x;
            // |x| saturated to 127, minus beta = 1, floored at 0:  max(x - 1, ~x, 0), capped at 126
            const uint32_t mg = vmax(vmax(vadd(x, h2(-1)), ~x), 0u); // (the cap at 126 for x = -128 is applied to the two minima)
            mag[d] = mg;
            uint32_t key = imad(mg, 32u, h2(d));
            if (d == 1 && first)
                key |= 0x00007fffu; // never the minimum
            k1 = vmin(k1, vmax(k0, key));
            k0 = vmin(k0, key);
        } else {
            v[d] = 0;
            mag[d] = 0;
        }
    }
    // order statistics commute with the monotone cap: min(|x|, 127) - 1 <= 126 taken once per minimum, not per link
    const uint32_t min0 = vmin((k0 >> 5) & 0x07ff07ffu, h2(126));
    const uint32_t min1 = vmin((k1 >> 5) & 0x07ff07ffu, h2(126));
    // new posterior = v +- m, m = min over the OTHER links = min0 + min1 - min(mag, min1):
    //   v + m = (v + s01 + 1) + ~t,   v - m = (v - s01) + t,   t = min(mag, min1)
    const uint32_t s01 = vadd(min0, min1);
    const uint32_t s01p1 = vadd(s01, h2(1)), ns01 = vsub(0u, s01);
    uint32_t newsg_lo = 0, newsg_hi = 0, syn = 0, zer = 0;
#pragma unroll
    for (int dd = 0; dd < DEG_MAX; ++dd) {
        const int d = DEG_MAX - 1 - dd; // descending: the sign bits are shifted in from the top
        const bool live = UNIFORM || (d < 2) || (d - 2 < lv.cnt);
        if (!WIDE || d < 16)
            newsg_lo = vadd(newsg_lo, newsg_lo);
        else
            newsg_hi = vadd(newsg_hi, newsg_hi);
        if (live) {
            const uint32_t t = vmin(mag[d], min1);
            const uint32_t np = vadd(vadd(v[d], s01p1), ~t);
            const uint32_t nn = vadd(vadd(v[d], ns01), t);
            const uint32_t ng = signmask(sx ^ v[d]); // product of the other signs, zero counts as +
            const uint32_t nl = vmin(vmax(bsel(ng, nn, np), h2(-128)), h2(127)); // vqadd
            const uint32_t packed = prmt(nl, 0, sel[d] >> 16);
            if (d == 1 && first)
                L[adr[d]] = (int8_t)(nl >> 16); // only node B's link exists (low byte of that word)
            else
                *reinterpret_cast<uint16_t*>(L + adr[d]) = (uint16_t)packed;
            // ng is 0xFFFF = -1 in a half whose new message is negative: subtracting it shifts a 1 in
            if (!WIDE || d < 16)
                newsg_lo = vsub(newsg_lo, ng);
            else
                newsg_hi = vsub(newsg_hi, ng);
            if (SELF_CHECK) {
                uint32_t nlc = nl;
                if (d == 1 && first)
                    nlc = (nl & 0xffff0000u) | 1u; // absent link: positive, non-zero
                syn ^= nlc;
                zer |= vsub(nlc, h2(1)) & ~nlc; // bit 15 of a half set iff that half is 0
            }
        }
    }
    // ---- encode the new state ----
    const uint32_t c0 = vmin(min0, h2(32)), c1 = vmin(min1, h2(32));
    uint32_t nA = (c0 & 0xffffu) | ((c1 & 0xffffu) << 6) | ((k0 & 31u) << 12);
    uint32_t nB = (c0 >> 16) | ((c1 >> 16) << 6) | (((k0 >> 16) & 31u) << 12);
    if (!WIDE) {
        nA |= (newsg_lo & 0x7fffu) << 17;
        nB |= ((newsg_lo >> 16) & 0x7fffu) << 17;
        stg_hint(reinterpret_cast<uint2*>(msg_out), make_uint2(nA, nB), pol);
    } else {
        const uint32_t tA = (newsg_lo & 0xffffu) | (newsg_hi << 16);
        const uint32_t tB = (newsg_lo >> 16) | (newsg_hi & 0xffff0000u);
        stg_hint(reinterpret_cast<uint4*>(msg_out), make_uint4(nA, nB, tA, tB), pol);
    }
    if (SELF_CHECK)
        return (int)(((syn | zer) & 0x80008000u) != 0);
    return 0;
}

#if !DVBS2_LEGACY_WAVEFRONT
// ---- split step: a conflict layer with the pair mapping kept (code_tables.h) ----------------------
// Phase 1: private links of check nodes p and p+180 in s16x2 (partial minima / sign product).
// Phase 2: the levels of the serial order; a node merges its shared links (scalar, one half of the
//          registers) and updates those bits.  Every thread of the block takes part in the barriers.
// Phase 3: the private links are updated with the final minima; the state word is written.
constexpr int kMaxShared = 12; // = code_tables.h kMaxSharedLinks

// Phase 2 of a split step.  The chain through the levels of a layer is what bounds the step, so per
// level everything that does not depend on the predecessors' writes (operand addresses, old messages,
// the partial minima of this node) is prepared BEFORE the level's barrier, and only
// load -> vqsub -> merge minima -> vqadd -> store sits between the barrier and the hand-over to the next level.
template <bool WIDE>
struct SplitCtx {
    int8_t* L;
    const uint2* sh_edges;      // the layer's shared circulants
    const uint16_t* level_tab;  // [360] level of node j, then [depth + 1] barrier thread counts
    volatile int* progress;
    int p;
    uint32_t pkey;
    int d0, nshared, depth;     // d0 = link index of the first shared link
    int lvA, lvB;
    bool active;
    uint32_t wA, wB, sA, sB;
    unsigned long long* prof; // diagnostics build: [8..12] = barrier wait, chain, hand-over cycles, levels, pre cycles (per warp lane 0)
};

template <int NSH>
struct SplitNode {
    int a[NSH > 0 ? NSH : 1];   // byte addresses of the operands
    int old[NSH > 0 ? NSH : 1]; // old messages
    int k0h, k1h, sxh, hsel;
    uint32_t w, sg;
};

template <int NSH, bool WIDE>
__device__ __forceinline__ void split_operand(const SplitCtx<WIDE>& c, const SplitNode<NSH>& n, int s, int& a, int& old)
{
    const int d = c.d0 + s;
    const uint2 e = c.sh_edges[s];
    const bool lt = e.x > c.pkey;
    // node p is the low byte of the halfword iff (ra ^ (p < a')) == 0, ra = bit 0 of the unpack selector
    a = (int)(e.x & 0xffffu) + 2 * c.p - (lt ? 0 : 360) + (int)(((e.y ^ (lt ? 1u : 0u)) & 1u) ^ (uint32_t)n.hsel);
    const int mc = (d == (int)((n.w >> 12) & 31u)) ? (int)((n.w >> 6) & 63u) : (int)(n.w & 63u);
    old = ((n.sg >> d) & 1u) ? -mc : min(mc, 31);
}

template <int NSH, bool WIDE>
__device__ __forceinline__ void split_pre(const SplitCtx<WIDE>& c, SplitNode<NSH>& n, int hsel, uint32_t k0, uint32_t k1, uint32_t sx)
{
    n.hsel = hsel;
    n.w = hsel ? c.wB : c.wA;
    n.sg = WIDE ? (hsel ? c.sB : c.sA) : (n.w >> 17);
    n.k0h = (int)((k0 >> (16 * hsel)) & 0xffffu);
    n.k1h = (int)((k1 >> (16 * hsel)) & 0xffffu);
    n.sxh = (int)(int16_t)(sx >> (16 * hsel));
#pragma unroll
    for (int s = 0; s < NSH; ++s)
        split_operand<NSH, WIDE>(c, n, s, n.a[s], n.old[s]);
}

// between the barrier of the level and the hand-over: the node's shared operands are read, merged, updated
template <int NSH, bool WIDE>
__device__ __forceinline__ void split_chain(const SplitCtx<WIDE>& c, SplitNode<NSH>& n, uint32_t& shs_lo, uint32_t& shs_hi)
{
    const int sh = 16 * n.hsel;
    if (NSH > 0) {
        int x[NSH > 0 ? NSH : 1], mg[NSH > 0 ? NSH : 1];
#pragma unroll
        for (int s = 0; s < NSH; ++s)
            x[s] = (int)c.L[n.a[s]];
#pragma unroll
        for (int s = 0; s < NSH; ++s) {
            x[s] = min(max(x[s] - n.old[s], -128), 127);
            n.sxh ^= x[s];
            mg[s] = max(min(abs(x[s]), 127) - 1, 0);
            const int key = mg[s] * 32 + c.d0 + s;
            n.k1h = min(n.k1h, max(n.k0h, key));
            n.k0h = min(n.k0h, key);
        }
        const int min0 = n.k0h >> 5, min1 = n.k1h >> 5;
#pragma unroll
        for (int s = 0; s < NSH; ++s) {
            const int m = min0 + min1 - min(mg[s], min1);
            const bool neg = ((n.sxh ^ x[s]) < 0);
            c.L[n.a[s]] = (int8_t)min(max(x[s] + (neg ? -m : m), -128), 127);
            const int d = c.d0 + s;
            const uint32_t bit = (neg ? 1u : 0u) << ((d & 15) + sh);
            if (!WIDE || d < 16)
                shs_lo |= bit;
            else
                shs_hi |= bit;
        }
    } else {
#pragma unroll 1
        for (int s = 0; s < c.nshared; ++s) {
            int a, old;
            split_operand<NSH, WIDE>(c, n, s, a, old);
            const int x = min(max((int)c.L[a] - old, -128), 127);
            n.sxh ^= x;
            const int key = max(min(abs(x), 127) - 1, 0) * 32 + c.d0 + s;
            n.k1h = min(n.k1h, max(n.k0h, key));
            n.k0h = min(n.k0h, key);
        }
        const int min0 = n.k0h >> 5, min1 = n.k1h >> 5;
#pragma unroll 1
        for (int s = 0; s < c.nshared; ++s) {
            int a, old;
            split_operand<NSH, WIDE>(c, n, s, a, old); // the operands are still unmodified: each is written once, below
            const int x = min(max((int)c.L[a] - old, -128), 127);
            const int m = min0 + min1 - min(max(min(abs(x), 127) - 1, 0), min1);
            const bool neg = ((n.sxh ^ x) < 0);
            c.L[a] = (int8_t)min(max(x + (neg ? -m : m), -128), 127);
            const int d = c.d0 + s;
            const uint32_t bit = (neg ? 1u : 0u) << ((d & 15) + sh);
            if (!WIDE || d < 16)
                shs_lo |= bit;
            else
                shs_hi |= bit;
        }
    }
}

__device__ __forceinline__ void split_writeback(int hsel, int k0h, int k1h, int sxh, uint32_t& k0, uint32_t& k1, uint32_t& sx)
{
    const uint32_t keep = hsel ? 0x0000ffffu : 0xffff0000u;
    k0 = (k0 & keep) | ((uint32_t)k0h << (16 * hsel));
    k1 = (k1 & keep) | ((uint32_t)k1h << (16 * hsel));
    sx = (sx & keep) | (((uint32_t)sxh & 0xffffu) << (16 * hsel));
}

// Levels rise with j, so the nodes of a warp sit in two contiguous level ranges (nodes p, nodes p+180).
// A warp only takes part in the levels it has nodes in.  Level l is entered through named barrier
// 1 + l % 15 whose participants are the warps with nodes in level l-1 (their writes must be visible:
// they arrive, or sync if they also have nodes in level l) and in level l (they sync); the host
// precomputed the thread count of each barrier (level_tab[360 + l]).
// Not inlined: one copy per (NSH, WIDE) serves every kernel instantiation, and the register allocation of the
// hot conflict-free pair step is not burdened with this code.
template <int NSH, bool WIDE>
__device__ __forceinline__ void split_levels(const SplitCtx<WIDE>* cp, uint32_t* io)
{
    const SplitCtx<WIDE> c = *cp;
    uint32_t k0 = io[0], k1 = io[1], sx = io[2], shs_lo = 0, shs_hi = 0;
    const int loA = (int)__reduce_min_sync(0xffffffffu, c.active ? (unsigned)c.lvA : 0xffffu);
    const int hiA = (int)__reduce_max_sync(0xffffffffu, c.active ? (unsigned)c.lvA : 0u);
    const int loB = (int)__reduce_min_sync(0xffffffffu, c.active ? (unsigned)c.lvB : 0xffffu);
    const int hiB = (int)__reduce_max_sync(0xffffffffu, c.active ? (unsigned)c.lvB : 0u);
    auto has_nodes = [&](int l) { return (l >= loA && l <= hiA) || (l >= loB && l <= hiB); };
    int cnt_cur = (int)__ldg(c.level_tab + 360 + loA);
#ifdef DVBS2_SKIP_PHASE2 // timing experiment only: results are wrong
    for (int lvl = loA; lvl < loA; ++lvl) {
#else
    for (int lvl = loA; lvl <= hiB; ++lvl) {
#endif
        if (!has_nodes(lvl)) {
            lvl = loB - 1; // the gap between the two ranges
            cnt_cur = (int)__ldg(c.level_tab + 360 + loB);
            continue;
        }
#ifdef DVBS2_PHASE_PROFILE
        const long long tp0 = clock64();
#endif
        const bool next_mine = has_nodes(lvl + 1);
        const int cnt_next = (int)__ldg(c.level_tab + 360 + min(lvl + 1, c.depth));
        const bool doA = c.lvA == lvl, doB = c.lvB == lvl;
        SplitNode<NSH> n;
        if (doA || doB)
            split_pre<NSH, WIDE>(c, n, doA ? 0 : 1, k0, k1, sx);
#ifdef DVBS2_PHASE_PROFILE
        const long long tp1 = clock64();
#endif
        if (lvl > 1) {
            // A warp may get here long before the chain does.  The barrier id is shared with level
            // lvl - 15: wait until that one has completed (progress = highest level known complete).
            if (lvl > 15)
                while (*c.progress < lvl - 16)
                    __nanosleep(64);
            asm volatile("bar.sync %0, %1;" ::"r"(1 + lvl % 15), "r"(cnt_cur) : "memory");
        }
#ifdef DVBS2_PHASE_PROFILE
        const long long tp2 = clock64();
#endif
        if (doA || doB)
            split_chain<NSH, WIDE>(c, n, shs_lo, shs_hi);
        if (doA && doB) { // both nodes of the thread in one level (shallow layers only)
            split_writeback(0, n.k0h, n.k1h, n.sxh, k0, k1, sx);
            split_pre<NSH, WIDE>(c, n, 1, k0, k1, sx);
            split_chain<NSH, WIDE>(c, n, shs_lo, shs_hi);
        }
        // hand over to level lvl + 1: if this warp has nodes there it syncs at the top of the loop
#ifdef DVBS2_PHASE_PROFILE
        const long long tp3 = clock64();
#endif
        if (lvl < c.depth && !next_mine)
            asm volatile("bar.arrive %0, %1;" ::"r"(1 + (lvl + 1) % 15), "r"(cnt_next) : "memory");
        if (lvl > 1 && (threadIdx.x & 31) == 0)
            atomicMax(const_cast<int*>(c.progress), lvl - 1);
        if (doA || doB)
            split_writeback(n.hsel, n.k0h, n.k1h, n.sxh, k0, k1, sx);
        cnt_cur = cnt_next;
#ifdef DVBS2_PHASE_PROFILE
        if (c.prof && (threadIdx.x & 31) == 0) {
            const long long tp4 = clock64();
            atomicAdd(c.prof + 8, (unsigned long long)(tp2 - tp1));
            atomicAdd(c.prof + 9, (unsigned long long)(tp3 - tp2));
            atomicAdd(c.prof + 10, (unsigned long long)(tp4 - tp3));
            atomicAdd(c.prof + 11, 1ull);
            atomicAdd(c.prof + 12, (unsigned long long)(tp1 - tp0));
        }
#endif
    }
    io[0] = k0, io[1] = k1, io[2] = sx, io[3] = shs_lo, io[4] = shs_hi;
}
// Out-of-line copies with compile-time link counts.  A call in the kernel costs every thread a local-memory
// frame (register saves around it), whose traffic competes with the check-node state in L2: codes whose
// conflict layers mostly take the chain form run the kernel variant WITHOUT these calls (one inlined rolled
// loop for the few level-form layers, no stack at all); codes with many level-form layers of three or four
// shared links (2/3 short) are faster with them.  The host chooses (code_tables.cc:choose_level_calls).
template <int NSH, bool WIDE>
__device__ __noinline__ void split_levels_call(const SplitCtx<WIDE>* cp, uint32_t* io)
{
    split_levels<NSH, WIDE>(cp, io);
}

// ---- chain form of phase 2 (code_tables.h) ------------------------------------------------------------
// Node record in the shared-memory scratch, written by phase 1 and rewritten by the walk:
//   x: partial k0 (low half) | partial k1 (high half) -> final k0 | k1
//   y: bit 0 sign of the partial sign product, bits 1..6 old message of shared link 0 (+32), bits 7..12 of link 1
//      -> bit 0 final sign product, bit 1 / 2 sign of the new message on shared link 0 / 1
__device__ __forceinline__ int chain_operand_addr(const uint2 e, int j)
{
    const int ap = (int)(e.x >> 16), ra = (int)(e.y & 1u);
    const int gbase = (int)(e.x & 0xffffu) - 360 + 2 * ap;
    int s = j - ap - kPairs * ra; // (j - shift) mod 360
    s += (s < 0) ? 360 : 0;
    return data_addr(gbase, s);
}

// Lane c < delta walks the chain c, c + delta, ...  The bit that a node updates through its "out" link is the
// bit the next node reads through its "in" link: it travels in a register.  A chain's last node meets, through
// its out link, the bit that the first node of some chain updated through its in link at step 0 -- hence one
// barrier among the walking warps after step 0, and none after that.
__device__ __forceinline__ void split_chain_walk(int8_t* __restrict__ L, uint2* __restrict__ rec, const uint2 e_in, const uint2 e_out,
                                              int d_in, int d_out, int in_is_link1, int delta, int tid)
{
    // called by every lane of the first ceil(delta / 32) warps (the barrier below is warp granular)
    const int nwarps = (delta + 31) >> 5;
    const bool live = tid < delta;
    // what a node needs that does not depend on its predecessor: fetched one node ahead
    struct Pre {
        uint2 r;
        int a_in, a_out, l_out;
    };
    auto fetch = [&](int j) {
        Pre q;
        q.r = rec[j];
        q.a_in = chain_operand_addr(e_in, j);
        q.a_out = chain_operand_addr(e_out, j);
        q.l_out = (int)L[q.a_out];
        return q;
    };
    // the node itself; returns the updated bit of its out link
    auto node = [&](int j, const Pre& q, int l_in) {
        const uint2 r = q.r;
        const int old0 = (int)((r.y >> 1) & 63u) - 32, old1 = (int)((r.y >> 7) & 63u) - 32;
        const int old_in = in_is_link1 ? old1 : old0, old_out = in_is_link1 ? old0 : old1;
        const int x_in = min(max(l_in - old_in, -128), 127), x_out = min(max(q.l_out - old_out, -128), 127);
        const int mg_in = max(min(abs(x_in), 127) - 1, 0), mg_out = max(min(abs(x_out), 127) - 1, 0);
        int k0h = (int)(r.x & 0xffffu), k1h = (int)(r.x >> 16);
        const int key_in = mg_in * 32 + d_in, key_out = mg_out * 32 + d_out;
        // the out link does not depend on the predecessor: merged first, the in link last
        k1h = min(k1h, max(k0h, key_out));
        k0h = min(k0h, key_out);
        k1h = min(k1h, max(k0h, key_in));
        k0h = min(k0h, key_in);
        const int min0 = k0h >> 5, min1 = k1h >> 5;
        const int sgn = (int)(r.y & 1u) ^ (x_in < 0) ^ (x_out < 0); // 1: the product of all signs is negative
        const int m_in = min0 + min1 - min(mg_in, min1), m_out = min0 + min1 - min(mg_out, min1);
        const int neg_in = sgn ^ (x_in < 0), neg_out = sgn ^ (x_out < 0);
        const int nl_in = min(max(x_in + (neg_in ? -m_in : m_in), -128), 127);
        const int nl_out = min(max(x_out + (neg_out ? -m_out : m_out), -128), 127);
        L[q.a_in] = (int8_t)nl_in;
        if (j + delta >= 360)
            L[q.a_out] = (int8_t)nl_out; // end of the chain: nobody takes the bit over
        const int s0 = in_is_link1 ? neg_out : neg_in, s1 = in_is_link1 ? neg_in : neg_out;
        rec[j] = make_uint2((uint32_t)k0h | ((uint32_t)k1h << 16), (uint32_t)sgn | ((uint32_t)s0 << 1) | ((uint32_t)s1 << 2));
        return nl_out;
    };
    int carried = 0;
    if (live) {
        const Pre q = fetch(tid);
        carried = node(tid, q, (int)L[q.a_in]);
    }
    if (nwarps == 1)
        __syncwarp();
    else
        asm volatile("bar.sync 1, %0;" ::"r"(nwarps * 32) : "memory");
    if (live && tid + delta < 360) {
        // (no fetch ahead across the barrier: a chain's second node may already be its last, whose out link
        // meets a bit that a first node has just written)
        int j = tid + delta;
        Pre q = fetch(j);
        for (;;) {
            const int jn = j + delta;
            Pre qn = q;
            if (jn < 360)
                qn = fetch(jn);
            carried = node(j, q, carried);
            if (jn >= 360)
                break;
            j = jn;
            q = qn;
        }
    }
}

template <int CNT_MAX, bool UNIFORM, bool WIDE, bool SELF_CHECK, bool LEVEL_CALLS>
__device__ __forceinline__ int process_split(int8_t* __restrict__ L, const uint2* __restrict__ edges, const LayerView lv, int layer,
                                             int p, bool active, int K, int q, uint32_t wA, uint32_t wB, uint32_t sA, uint32_t sB,
                                             uint32_t* __restrict__ msg_out, uint64_t pol, int depth, const uint16_t* __restrict__ level_tab,
                                             volatile int* progress, unsigned long long* prof, uint2* __restrict__ rec, bool chain, int chain_delta, int chain_out_link)
{
    constexpr int DEG_MAX = CNT_MAX + 2;
    const bool first = (layer == 0 && p == 0);
    const int half = kPairs * q;
    const int npriv = lv.cnt - lv.nshared; // private data links: d - 2 < npriv
    int lvA = 0, lvB = 0;
    if (active && !chain) {
        lvA = (int)__ldg(level_tab + p);
        lvB = (int)__ldg(level_tab + p + kPairs);
    }
    const uint32_t m0 = (wA & 63u) | ((wB & 63u) << 16);
    const uint32_t m1 = ((wA >> 6) & 63u) | (((wB >> 6) & 63u) << 16);
    const uint32_t x01 = m0 ^ m1;
    const uint32_t argA = (wA >> 12) & 31u, argB = (wB >> 12) & 31u;
    uint32_t hot_lo, hot_hi = 0, nsg_lo, nsg_hi = 0;
    if (!WIDE) {
        hot_lo = (1u << argA) | (0x10000u << argB);
        nsg_lo = ~((wA >> 17) | ((wB >> 17) << 16));
    } else {
        hot_lo = ((1u << argA) & 0xffffu) | (((1u << argB) & 0xffffu) << 16);
        hot_hi = ((1u << argA) >> 16) | ((1u << argB) & 0xffff0000u);
        nsg_lo = ~((sA & 0xffffu) | (sB << 16));
        nsg_hi = ~((sA >> 16) | (sB & 0xffff0000u));
    }
    // Only the private links' x values stay live across phase 2; addresses, selectors and magnitudes are
    // recomputed in phase 3 (a few instructions per link), which keeps the step inside the register budget
    // with phase 2 inlined.
    uint32_t v[DEG_MAX];
    const int c = q * p + layer;
    const uint32_t pkey = ((uint32_t)p << 16) | 0xffffu;
    auto link_addr = [&](int d, int& adr, uint32_t& sel) {
        if (d == 0) {
            adr = K + 2 * c;
            sel = 0x9180u | (0x4420u << 16);
        } else if (d == 1) {
            adr = first ? K + 2 * (half - 1) : K + 2 * c - 2;
            sel = first ? (0x8091u | (0x4402u << 16)) : (0x9180u | (0x4420u << 16));
        } else {
            const uint2 e = edges[lv.edge_begin + d - 2];
            const bool lt = e.x > pkey;
            adr = (int)(e.x & 0xffffu) + 2 * p - (lt ? 0 : 360);
            sel = lt ? (e.y ^ 0x00221111u) : e.y;
        }
    };
    const uint32_t n0p = vsub(0u, vmin(m0, h2(31))), n1p = vsub(0u, vmin(m1, h2(31)));
    const uint32_t nx01p = n0p ^ n1p;
    uint32_t k0 = h2(0x7fff), k1 = h2(0x7fff), sx = 0;
    if (active) {
#pragma unroll
        for (int d = 0; d < DEG_MAX; ++d) {
            const bool live = (d < 2) || (d - 2 < npriv);
            if (live) {
                int adr;
                uint32_t sel;
                link_addr(d, adr, sel);
                const uint32_t raw = *reinterpret_cast<const uint16_t*>(L + adr);
                const uint32_t l = prmt(raw, 0, sel);
                const uint32_t hot = (!WIDE || d < 16) ? hot_lo : hot_hi;
                const uint32_t nsg = (!WIDE || d < 16) ? nsg_lo : nsg_hi;
                const uint32_t im = signmask(imad(hot, 1u << (15 - (d & 15)), 0u)); // the shift as IMAD: FMA pipe, not ALU
                const uint32_t nm = signmask(imad(nsg, 1u << (15 - (d & 15)), 0u));
                const uint32_t xa = vadd(l, n0p ^ (nx01p & im));
                const uint32_t xb = vadd(l, m0 ^ (x01 & im));
                uint32_t x = vmin(vmax(bsel(nm, xa, xb), h2(-128)), h2(127));
                if (d == 1 && first)
                    x &= 0xffff0000u;
                v[d] = x;
                sx ^= x;
                const uint32_t mg = vmin(vmax(vmax(vadd(x, h2(-1)), ~x), 0u), h2(126));
                uint32_t key = imad(mg, 32u, h2(d));
                if (d == 1 && first)
                    key |= 0x00007fffu;
                k1 = vmin(k1, vmax(k0, key));
                k0 = vmin(k0, key);
            } else {
                v[d] = 0;
            }
        }
    }
    // ---- phase 2: shared links, level by level ----
    // Levels rise with j, so the nodes of a warp sit in two contiguous level ranges (nodes p, nodes p+180).
    // A warp only takes part in the levels it has nodes in.  Level l is entered through named barrier
    // 1 + l % 15 whose participants are the warps with nodes in level l-1 (their writes must be visible:
    // they arrive, or sync if they also have nodes in level l) and in level l (they sync); the host
    // precomputed the thread count of each barrier (level_tab[360 + l]).
    uint32_t shs_lo = 0, shs_hi = 0; // new sign bits of the shared links, same layout as newsg_lo / newsg_hi
    if (chain) {
        // chain form: hand the nodes' partial results and old shared-link messages to the walking lanes
        const int d0 = 2 + npriv;
#ifdef DVBS2_PHASE_PROFILE
        const long long tq0 = clock64();
#endif
        if (active) {
#pragma unroll
            for (int hsel = 0; hsel < 2; ++hsel) {
                const uint32_t w = hsel ? wB : wA;
                const uint32_t sg = WIDE ? (hsel ? sB : sA) : (w >> 17);
                const int omin0 = (int)(w & 63u), omin1 = (int)((w >> 6) & 63u), oarg = (int)((w >> 12) & 31u);
                uint32_t y = (sx >> (16 * hsel + 15)) & 1u;
#pragma unroll
                for (int sl = 0; sl < 2; ++sl) {
                    const int d = d0 + sl;
                    const int mc = (d == oarg) ? omin1 : omin0;
                    const int old = ((sg >> d) & 1u) ? -mc : min(mc, 31);
                    y |= (uint32_t)(old + 32) << (1 + 6 * sl);
                }
                rec[p + kPairs * hsel] = make_uint2(((k0 >> (16 * hsel)) & 0xffffu) | (((k1 >> (16 * hsel)) & 0xffffu) << 16), y);
            }
        }
        __syncthreads();
#ifdef DVBS2_PHASE_PROFILE
        const long long tq1 = clock64();
#endif
        const int delta = chain_delta, out_link = chain_out_link;
        if ((p >> 5) < ((delta + 31) >> 5)) {
            const uint2 e0 = edges[lv.edge_begin + npriv], e1 = edges[lv.edge_begin + npriv + 1];
            split_chain_walk(L, rec, out_link ? e0 : e1, out_link ? e1 : e0, out_link ? d0 : d0 + 1, out_link ? d0 + 1 : d0,
                             out_link ? 0 : 1, delta, p);
        }
#ifdef DVBS2_PHASE_PROFILE
        const long long tq2 = clock64();
#endif
        __syncthreads();
#ifdef DVBS2_PHASE_PROFILE
        if (prof && p == 0) {
            const long long tq3 = clock64();
            atomicAdd(prof + 8, (unsigned long long)(tq1 - tq0));  // records + barrier
            atomicAdd(prof + 9, (unsigned long long)(tq2 - tq1));  // warp 0's walk
            atomicAdd(prof + 10, (unsigned long long)(tq3 - tq2)); // waiting for the other walking warps
            atomicAdd(prof + 11, (unsigned long long)depth);
        }
#endif
        if (active) {
#pragma unroll
            for (int hsel = 0; hsel < 2; ++hsel) {
                const uint2 r = rec[p + kPairs * hsel];
                const uint32_t keep = hsel ? 0x0000ffffu : 0xffff0000u;
                k0 = (k0 & keep) | ((r.x & 0xffffu) << (16 * hsel));
                k1 = (k1 & keep) | ((r.x >> 16) << (16 * hsel));
                sx = (sx & keep) | ((r.y & 1u) ? (0x8000u << (16 * hsel)) : 0u); // only the sign of sx is used from here on
#pragma unroll
                for (int sl = 0; sl < 2; ++sl) {
                    const int d = d0 + sl;
                    const uint32_t bit = ((r.y >> (1 + sl)) & 1u) << ((d & 15) + 16 * hsel);
                    if (!WIDE || d < 16)
                        shs_lo |= bit;
                    else
                        shs_hi |= bit;
                }
            }
        }
    } else {
        SplitCtx<WIDE> cx = { L, edges + lv.edge_begin + npriv, level_tab, progress, p, pkey, 2 + npriv, lv.nshared, depth, lvA, lvB,
                              active, wA, wB, sA, sB, prof };
        uint32_t io[5] = { k0, k1, sx, 0u, 0u };
        if (!LEVEL_CALLS) {
            split_levels<0, WIDE>(&cx, io); // one rolled variant, inlined: no call in this kernel variant
        } else {
            switch (lv.nshared) { // compile-time link counts for the common cases, a rolled loop for the rest
            case 2: split_levels_call<2, WIDE>(&cx, io); break;
            case 3: split_levels_call<3, WIDE>(&cx, io); break;
            case 4: split_levels_call<4, WIDE>(&cx, io); break;
            default: split_levels_call<0, WIDE>(&cx, io); break;
            }
        }
        k0 = io[0], k1 = io[1], sx = io[2], shs_lo = io[3], shs_hi = io[4];
    }
    if (SELF_CHECK)
        __syncthreads(); // the shared bits are re-read below, final only after the last level
    if (!active)
        return 0;
    // ---- phase 3: private links ----
    const uint32_t min0 = (k0 >> 5) & 0x07ff07ffu;
    const uint32_t min1 = (k1 >> 5) & 0x07ff07ffu;
    const uint32_t s01 = vadd(min0, min1);
    const uint32_t s01p1 = vadd(s01, h2(1)), ns01 = vsub(0u, s01);
    uint32_t newsg_lo = 0, newsg_hi = 0, syn = 0, zer = 0;
#pragma unroll
    for (int dd = 0; dd < DEG_MAX; ++dd) {
        const int d = DEG_MAX - 1 - dd;
        const bool live = (d < 2) || (d - 2 < npriv);
        if (!WIDE || d < 16)
            newsg_lo = vadd(newsg_lo, newsg_lo);
        else
            newsg_hi = vadd(newsg_hi, newsg_hi);
        if (live) {
            int adr;
            uint32_t sel;
            link_addr(d, adr, sel);
            const uint32_t mg = vmin(vmax(vmax(vadd(v[d], h2(-1)), ~v[d]), 0u), h2(126));
            const uint32_t t = vmin(mg, min1);
            const uint32_t np = vadd(vadd(v[d], s01p1), ~t);
            const uint32_t nn = vadd(vadd(v[d], ns01), t);
            const uint32_t ng = signmask(sx ^ v[d]);
            const uint32_t nl = vmin(vmax(bsel(ng, nn, np), h2(-128)), h2(127));
            const uint32_t packed = prmt(nl, 0, sel >> 16);
            if (d == 1 && first)
                L[adr] = (int8_t)(nl >> 16);
            else
                *reinterpret_cast<uint16_t*>(L + adr) = (uint16_t)packed;
            // ng is 0xFFFF = -1 in a half whose new message is negative: subtracting it shifts a 1 in
            if (!WIDE || d < 16)
                newsg_lo = vsub(newsg_lo, ng);
            else
                newsg_hi = vsub(newsg_hi, ng);
            if (SELF_CHECK) {
                uint32_t nlc = nl;
                if (d == 1 && first)
                    nlc = (nl & 0xffff0000u) | 1u;
                syn ^= nlc;
                zer |= vsub(nlc, h2(1)) & ~nlc;
            }
        } else if (SELF_CHECK && (UNIFORM || d - 2 < lv.cnt)) {
            int adr;
            uint32_t sel;
            link_addr(d, adr, sel);
            const uint32_t raw = *reinterpret_cast<const uint16_t*>(L + adr);
            const uint32_t nlc = prmt(raw, 0, sel);
            syn ^= nlc;
            zer |= vsub(nlc, h2(1)) & ~nlc;
        }
    }
    newsg_lo |= shs_lo;
    newsg_hi |= shs_hi;
    const uint32_t c0 = vmin(min0, h2(32)), c1 = vmin(min1, h2(32));
    uint32_t nA = (c0 & 0xffffu) | ((c1 & 0xffffu) << 6) | ((k0 & 31u) << 12);
    uint32_t nB = (c0 >> 16) | ((c1 >> 16) << 6) | (((k0 >> 16) & 31u) << 12);
    if (!WIDE) {
        nA |= (newsg_lo & 0x7fffu) << 17;
        nB |= ((newsg_lo >> 16) & 0x7fffu) << 17;
        stg_hint(reinterpret_cast<uint2*>(msg_out), make_uint2(nA, nB), pol);
    } else {
        const uint32_t tA = (newsg_lo & 0xffffu) | (newsg_hi << 16);
        const uint32_t tB = (newsg_lo >> 16) | (newsg_hi & 0xffff0000u);
        stg_hint(reinterpret_cast<uint4*>(msg_out), make_uint4(nA, nB, tA, tB), pol);
    }
    if (SELF_CHECK)
        return (int)(((syn | zer) & 0x80008000u) != 0);
    return 0;
}

#endif // !DVBS2_LEGACY_WAVEFRONT

// ---- legacy wavefront paths (whole check nodes level by level; -DDVBS2_LEGACY_WAVEFRONT=1) --------
#if DVBS2_LEGACY_WAVEFRONT
// ---- one check node j of a conflict layer (scalar, same arithmetic) -------------------------------
template <int CNT_MAX, bool UNIFORM, bool WIDE, bool SELF_CHECK>
__device__ __forceinline__ int process_cn(int8_t* __restrict__ L, const uint2* __restrict__ edges, const LayerView& lv, int layer,
                                          int j, int K, int q, uint32_t* __restrict__ msg_pair, bool zero_state, uint64_t pol,
                                          bool have_state, uint32_t pw, uint32_t psg, uint32_t* state_out = nullptr)
{
    constexpr int DEG_MAX = CNT_MAX + 2;
    const int hsel = j >= kPairs; // which node of the pair
    const int half = kPairs * q;
    uint32_t w = 0, sg = 0;
    if (have_state) {
        w = pw, sg = psg;
    } else if (!zero_state) {
        w = ldg_hint(msg_pair + hsel, pol);
        sg = WIDE ? ldg_hint(msg_pair + 2 + hsel, pol) : (w >> 17);
    }
    const int old_min0 = (int)(w & 63u), old_min1 = (int)((w >> 6) & 63u), old_arg = (int)((w >> 12) & 31u);
    const int c = q * j + layer;
    const bool has_prev = c > 0;
    int adr[DEG_MAX], v[DEG_MAX];
    adr[0] = parity_addr(K, half, c);
    adr[1] = parity_addr(K, half, has_prev ? c - 1 : 0);
#pragma unroll
    for (int d = 0; d < CNT_MAX; ++d) {
        if (UNIFORM || d < lv.cnt) {
            const uint2 e = edges[lv.edge_begin + d];
            const int ap = (int)(e.x >> 16), ra = (int)(e.y & 1u);
            const int gbase = (int)(e.x & 0xffffu) - 360 + 2 * ap;
            int s = j - ap - kPairs * ra; // (j - shift) mod 360, j < 360, shift < 360
            s += (s < 0) ? 360 : 0;
            adr[d + 2] = data_addr(gbase, s);
        } else {
            adr[d + 2] = 0;
        }
    }
    int min0 = 127, min1 = 127, arg = 0, sx = 0;
#pragma unroll
    for (int d = 0; d < DEG_MAX; ++d) {
        const bool live = (d == 0) || (d == 1 ? has_prev : (UNIFORM || d - 2 < lv.cnt));
        if (live) {
            const int l = (int)L[adr[d]];
            const int mc = (d == old_arg) ? old_min1 : old_min0;
            const int old = ((sg >> d) & 1u) ? -mc : min(mc, 31);
            const int x = min(max(l - old, -128), 127);
            v[d] = x;
            sx ^= x;
            const int mg = max(min(abs(x), 127) - 1, 0);
            if (mg < min0) {
                min1 = min0;
                min0 = mg;
                arg = d;
            } else {
                min1 = min(min1, mg);
            }
        } else {
            v[d] = 0;
        }
    }
    uint32_t new_signs = 0;
    int syn = 0, zer = 0;
#pragma unroll
    for (int d = 0; d < DEG_MAX; ++d) {
        const bool live = (d == 0) || (d == 1 ? has_prev : (UNIFORM || d - 2 < lv.cnt));
        if (live) {
            const int m = (d == arg) ? min1 : min0;
            const bool neg = ((sx ^ v[d]) < 0);
            const int nl = min(max(v[d] + (neg ? -m : m), -128), 127);
            L[adr[d]] = (int8_t)nl;
            new_signs |= (neg ? 1u : 0u) << d;
            if (SELF_CHECK) {
                syn ^= nl;
                zer |= (nl == 0);
            }
        }
    }
    const uint32_t lo = (uint32_t)min(min0, 32) | ((uint32_t)min(min1, 32) << 6) | ((uint32_t)arg << 12);
    if (state_out) {
        *state_out = lo | (new_signs << 17); // one-word state kept in tensor memory by the caller
    } else if (!WIDE) {
        stg_hint(msg_pair + hsel, lo | (new_signs << 17), pol);
    } else {
        stg_hint(msg_pair + hsel, lo, pol);
        stg_hint(msg_pair + 2 + hsel, new_signs, pol);
    }
    if (SELF_CHECK)
        return (syn < 0) | zer;
    return 0;
}

// ---- link-parallel run: narrow wavefront levels of a conflict layer ----------------------------------
// One lane per (check node, link): G = 8/16/32 lanes per check node, minima and sign parity by warp
// REDUX over the group, sign bits by ballot.  The dependent chain through a deep conflict layer (up to
// 180 levels for DVB-S2 3/4) then costs ~70 instructions per level on 1..6 warps instead of a full
// scalar check-node update, and the levels are ordered by a named barrier among the warps that take part.
__device__ __forceinline__ void sub_barrier(int nwarps)
{
    if (nwarps == 1)
        __syncwarp();
    else
        asm volatile("bar.sync 1, %0;" ::"r"(nwarps * 32) : "memory");
}

template <bool WIDE>
__device__ __noinline__ int lp_run(int8_t* __restrict__ L, const uint2* __restrict__ edges, const uint2* __restrict__ steps,
                                   const uint16_t* __restrict__ work, uint32_t* __restrict__ msg, int cnt, uint32_t edge_begin,
                                   int layer, int s0, int run_len, int nwarps, int last_step, int K, int q, bool zero_state,
                                   uint64_t pol, int tid)
{
    constexpr int MW = WIDE ? 2 : 1;
    const int deg = cnt + 2;
    const int gshift = deg <= 8 ? 3 : deg <= 16 ? 4 : 5;
    const int G = 1 << gshift;
    const int link = tid & (G - 1);
    const int lane = tid & 31;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
    const int nthreads = nwarps * 32;
    const int half = kPairs * q;
    const bool data_link = link >= 2 && link - 2 < cnt;
    int ap = 0, ra = 0, gbase = 0;
    if (data_link) {
        const uint2 e = edges[edge_begin + link - 2];
        ap = (int)(e.x >> 16);
        ra = (int)(e.y & 1u);
        gbase = (int)(e.x & 0xffffu) - 360 + 2 * ap;
    }
    int self_bad = 0;
    // The check node of a lane's first item and its state word are fetched one level ahead (the index
    // two levels ahead), so the L2 latency of the state is off the dependent chain through the layer.
    auto load_j = [&](int k) -> int {
        if (k >= run_len)
            return -1;
        const uint2 st = steps[s0 + k];
        return ((tid >> gshift) < (int)(st.x >> 16)) ? (int)work[(st.y & 0x00ffffffu) + (tid >> gshift)] : -1;
    };
    auto load_state = [&](int j, uint32_t& w, uint32_t& sg) {
        w = 0, sg = 0;
        if (j >= 0 && !zero_state) {
            const int hsel = j >= kPairs;
            const uint32_t* mp = msg + ((size_t)layer * kPairs + (j - kPairs * hsel)) * 2 * MW;
            w = ldg_hint(mp + hsel, pol);
            sg = WIDE ? ldg_hint(mp + 2 + hsel, pol) : (w >> 17);
        }
    };
    int j_cur = load_j(0), j_nxt = load_j(1);
    uint32_t w_cur, sg_cur, w_nxt = 0, sg_nxt = 0;
    load_state(j_cur, w_cur, sg_cur);
    for (int k = 0; k < run_len; ++k) {
        const uint2 st = steps[s0 + k];
        const int count = (int)(st.x >> 16);
        const uint32_t work_off = st.y & 0x00ffffffu;
        const bool self_check = (s0 + k == last_step);
        load_state(j_nxt, w_nxt, sg_nxt);
        const int j_nn = load_j(k + 2);
        // every lane of the participating warps runs every pass (warp-wide shuffles / ballots below)
        for (int item0 = 0; item0 < (count << gshift); item0 += nthreads) {
            const int item = item0 + tid;
            const bool grp = (item >> gshift) < count;
            int j = 0;
            uint32_t w = 0, sg = 0;
            if (item0 == 0) {
                j = j_cur < 0 ? 0 : j_cur, w = w_cur, sg = sg_cur;
            } else if (grp) {
                j = (int)work[work_off + (item >> gshift)];
                load_state(j, w, sg);
            }
            const int hsel = j >= kPairs;
            const int pp = j - kPairs * hsel;
            uint32_t* mp = msg + ((size_t)layer * kPairs + pp) * 2 * MW;
            const int c = q * j + layer;
            bool live;
            int adr;
            if (link == 0) {
                live = true;
                adr = parity_addr(K, half, c);
            } else if (link == 1) {
                live = c > 0;
                adr = parity_addr(K, half, live ? c - 1 : 0);
            } else {
                live = data_link;
                int sidx = j - ap - kPairs * ra;
                sidx += (sidx < 0) ? 360 : 0;
                adr = data_link ? data_addr(gbase, sidx) : 0;
            }
            live = live && grp;
            int x = 0, key = 0x7fffffff;
            if (live) {
                const int l = (int)L[adr];
                const int mc = (link == (int)((w >> 12) & 31u)) ? (int)((w >> 6) & 63u) : (int)(w & 63u);
                const int old = ((sg >> link) & 1u) ? -mc : min(mc, 31);
                x = min(max(l - old, -128), 127);
                key = max(min(abs(x), 127) - 1, 0) * 32 + link;
            }
            // Two smallest keys and the sign parity over the G lanes of the check node.  REDUX with a
            // member mask that differs between the groups of a warp falls into a slow divergent path,
            // so only full-warp groups use it; smaller groups use xor-butterfly shuffles and a ballot.
            int k0, k1;
            bool par;
            if (G == 32) {
                k0 = __reduce_min_sync(0xffffffffu, key);
                k1 = __reduce_min_sync(0xffffffffu, key == k0 ? 0x7fffffff : key);
                par = (int)__reduce_xor_sync(0xffffffffu, (unsigned)x) < 0;
            } else {
                int a0 = key, a1 = 0x7fffffff;
#pragma unroll
                for (int m = 1; m < 16; m <<= 1) {
                    if (m < G) {
                        const int b0 = __shfl_xor_sync(0xffffffffu, a0, m), b1 = __shfl_xor_sync(0xffffffffu, a1, m);
                        const int hi = max(a0, b0);
                        a0 = min(a0, b0);
                        a1 = min(hi, min(a1, b1));
                    }
                }
                k0 = a0, k1 = a1;
                par = __popc(__ballot_sync(0xffffffffu, x < 0) & gmask) & 1;
            }
            const int min0 = k0 >> 5, min1 = k1 >> 5, arg = k0 & 31;
            const int m = (link == arg) ? min1 : min0;
            const bool neg = par != (x < 0);
            const int nl = min(max(x + (neg ? -m : m), -128), 127);
            if (live)
                L[adr] = (int8_t)nl;
            const unsigned signs = (__ballot_sync(0xffffffffu, live && neg) & gmask) >> (lane & ~(G - 1));
            if (link == 0 && grp) {
                const uint32_t lo = (uint32_t)min(min0, 32) | ((uint32_t)min(min1, 32) << 6) | ((uint32_t)arg << 12);
                if (!WIDE) {
                    stg_hint(mp + hsel, lo | (signs << 17), pol);
                } else {
                    stg_hint(mp + hsel, lo, pol);
                    stg_hint(mp + 2 + hsel, signs, pol);
                }
            }
            if (self_check) {
                const unsigned negs = __ballot_sync(0xffffffffu, live && nl < 0) & gmask;
                const unsigned zero = __ballot_sync(0xffffffffu, live && nl == 0) & gmask;
                self_bad |= (int)((__popc(negs) & 1) | (zero != 0)) & (int)grp;
            }
        }
        sub_barrier(nwarps); // this level's writes are visible to the next level's lanes
        j_cur = j_nxt, w_cur = w_nxt, sg_cur = sg_nxt;
        j_nxt = j_nn;
    }
    return self_bad;
}

// ---- run of narrow wavefront levels on warp 0, one check node per lane, ordered by __syncwarp() -------
template <int CNT_MAX, bool UNIFORM, bool WIDE, bool TMEM>
__device__ __noinline__ int scalar_run(int8_t* __restrict__ L, const uint2* __restrict__ edges, const uint2* __restrict__ steps,
                                       const uint16_t* __restrict__ work, uint32_t* __restrict__ msg, const LayerView lv, int layer,
                                       int s0, int run_len, int last_step, int K, int q, bool zero_state, uint64_t pol, int tid,
                                       const uint8_t* __restrict__ tcol, uint32_t tmem_base)
{
    constexpr int MW = WIDE ? 2 : 1;
    int self_bad = 0;
    if (TMEM && tcol[s0] != 0xff) {
        // state of this run lives in tensor memory: one column per level, lane = position in the level
        int j_cur = -1, j_nxt = -1;
        {
            const uint2 st = steps[s0];
            j_cur = (tid < (int)(st.x >> 16)) ? (int)work[(st.y & 0x00ffffffu) + tid] : -1;
        }
        for (int k = 0; k < run_len; ++k) {
            if (k + 1 < run_len) {
                const uint2 st = steps[s0 + k + 1];
                j_nxt = (tid < (int)(st.x >> 16)) ? (int)work[(st.y & 0x00ffffffu) + tid] : -1;
            }
            const uint32_t taddr = tmem_base + (uint32_t)tcol[s0 + k];
            uint32_t w = zero_state ? 0u : tmem_ld(taddr);
            if (j_cur >= 0) {
                if (s0 + k == last_step)
                    self_bad |= process_cn<CNT_MAX, UNIFORM, WIDE, true>(L, edges, lv, layer, j_cur, K, q, msg, zero_state, pol, true, w, w >> 17, &w);
                else
                    process_cn<CNT_MAX, UNIFORM, WIDE, false>(L, edges, lv, layer, j_cur, K, q, msg, zero_state, pol, true, w, w >> 17, &w);
            }
            tmem_st(taddr, w);
            __syncwarp();
            j_cur = j_nxt;
            j_nxt = -1;
        }
        return self_bad;
    }
    // this lane's check node and its state word are fetched one level ahead
    auto load_j = [&](int k) -> int {
        if (k >= run_len)
            return -1;
        const uint2 st = steps[s0 + k];
        return (tid < (int)(st.x >> 16)) ? (int)work[(st.y & 0x00ffffffu) + tid] : -1;
    };
    auto load_state = [&](int j, uint32_t& w, uint32_t& sg) {
        w = 0, sg = 0;
        if (j >= 0 && !zero_state) {
            const int hsel = j >= kPairs;
            const uint32_t* mp = msg + ((size_t)layer * kPairs + (j - kPairs * hsel)) * 2 * MW;
            w = ldg_hint(mp + hsel, pol);
            sg = WIDE ? ldg_hint(mp + 2 + hsel, pol) : (w >> 17);
        }
    };
    int j_cur = load_j(0), j_nxt = load_j(1);
    uint32_t w_cur, sg_cur, w_nxt = 0, sg_nxt = 0;
    load_state(j_cur, w_cur, sg_cur);
    for (int k = 0; k < run_len; ++k) {
        load_state(j_nxt, w_nxt, sg_nxt);
        const int j_nn = load_j(k + 2);
        if (j_cur >= 0) {
            const int pp = j_cur >= kPairs ? j_cur - kPairs : j_cur;
            uint32_t* mp = msg + ((size_t)layer * kPairs + pp) * 2 * MW;
            if (s0 + k == last_step)
                self_bad |= process_cn<CNT_MAX, UNIFORM, WIDE, true>(L, edges, lv, layer, j_cur, K, q, mp, zero_state, pol, true, w_cur, sg_cur);
            else
                process_cn<CNT_MAX, UNIFORM, WIDE, false>(L, edges, lv, layer, j_cur, K, q, mp, zero_state, pol, true, w_cur, sg_cur);
        }
        __syncwarp();
        j_cur = j_nxt, w_cur = w_nxt, sg_cur = sg_nxt;
        j_nxt = j_nn;
    }
    return self_bad;
}

#endif // DVBS2_LEGACY_WAVEFRONT

// lib/ldpc_decoder/layered_decoder.hh:32-49 for the pair (p, p+180): unsatisfied if the sign product
// is not +, and a zero LLR counts as unsatisfied (vsign(.,0) = 0, test is "> 0").
template <int CNT_MAX, bool UNIFORM>
__device__ __forceinline__ uint32_t check_pair(const int8_t* __restrict__ L, const uint2* __restrict__ edges, const LayerView& lv,
                                               int layer, int p, int K, int q)
{
    const int c = q * p + layer;
    const bool first = (layer == 0 && p == 0);
    uint32_t raw = *reinterpret_cast<const uint16_t*>(L + K + 2 * c);
    uint32_t s = raw;
    uint32_t z = (raw - 0x0101u) & ~raw;
    if (!first)
        raw = *reinterpret_cast<const uint16_t*>(L + K + 2 * c - 2);
    else
        raw = ((uint32_t)(uint8_t)L[K + 2 * (kPairs * q - 1)] << 8) | 0x01u; // node B's link only
    s ^= raw;
    z |= (raw - 0x0101u) & ~raw;
    const uint32_t pkey = ((uint32_t)p << 16) | 0xffffu;
#pragma unroll
    for (int d = 0; d < CNT_MAX; ++d) {
        if (UNIFORM || d < lv.cnt) {
            const uint2 e = edges[lv.edge_begin + d];
            const bool lt = e.x > pkey;
            raw = *reinterpret_cast<const uint16_t*>(L + (int)(e.x & 0xffffu) + 2 * p - (lt ? 0 : 360));
            z |= (raw - 0x0101u) & ~raw;
            // bring node p's byte to the low position: swap iff (ra ^ lt)
            const uint32_t swap = (e.y & 1u) ^ (lt ? 1u : 0u);
            s ^= swap ? prmt(raw, 0, 0x4401) : raw;
        }
    }
    return (s | z) & 0x8080u;
}

template <int CNT_MAX, bool UNIFORM, bool WIDE, bool TMEM>
__global__ void __launch_bounds__(kLdpcThreads, (DVBS2_LEGACY_WAVEFRONT || CNT_MAX < DVBS2_TWO_CTA_FROM) ? kLdpcCtasPerSm : 2)
    ldpc_decode_kernel(const LdpcLaunch p)
{
    // fourth template flag: tensor-memory state (wavefront build) / out-of-line level calls (split build)
    constexpr bool USE_TMEM = TMEM && DVBS2_LEGACY_WAVEFRONT;
    extern __shared__ __align__(16) uint8_t smem[];
    int8_t* const L = reinterpret_cast<int8_t*>(smem);
    const uint2* layers = reinterpret_cast<const uint2*>(smem + p.smem_tab_off);
    const uint2* edges = reinterpret_cast<const uint2*>(smem + p.smem_tab_off + (size_t)p.q * 8);
    const uint2* steps = reinterpret_cast<const uint2*>(smem + p.smem_tab_off + (size_t)p.q * 8 + (size_t)p.n_circ * 8);
    const uint8_t* tcol = smem + p.smem_tab_off + (size_t)p.q * 8 + (size_t)p.n_circ * 8 + (size_t)p.n_steps * 8;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + p.smem_bar_off);
    uint2* const rec = reinterpret_cast<uint2*>(smem + p.smem_rec_off); // node scratch of chain-form split steps
    (void)rec;
    __shared__ int s_group_bad;
    __shared__ int s_abort;
    __shared__ uint32_t s_tmem_base;
    __shared__ int s_progress; // split steps: highest level of the current layer known to be complete

    const int tid = threadIdx.x;
    const int N = p.N, K = p.K, q = p.q, R = p.R;
    const int half = R / 2;
    constexpr int MW = WIDE ? 2 : 1; // state words per check node
    uint32_t* const msg = p.msg_scratch + (size_t)blockIdx.x * R * MW;
    const uint16_t* __restrict__ work = p.work;
    const uint64_t pol_keep = policy_evict_last(), pol_stream = policy_evict_first();

    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // tensor memory for the state of the order-sensitive layers: warp 0 allocates (and frees at the end)
    if (USE_TMEM && tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "n"(kTmemColsDev)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (USE_TMEM)
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (USE_TMEM)
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = USE_TMEM ? s_tmem_base : 0u;
    const uint32_t tmem_warp = tmem_base + ((uint32_t)((tid >> 5) & 3) << 21); // this warp's 32 USE_TMEM lanes
    (void)tmem_warp;
    (void)tcol;
    // stage the code tables once per CTA (TMA)
    if (tid == 0) {
        mbar_expect_tx(bar, p.tab_bytes);
        tma_load_1d(smem + p.smem_tab_off, p.tab, p.tab_bytes, bar);
    }
    mbar_wait(bar, 0);

#ifdef DVBS2_PHASE_PROFILE
    // diagnostics build: cycles per phase, per CTA (load, syndrome pass, pair steps, narrow runs, wide steps,
    // iteration end, output, total) -> p.prof[blockIdx][8]
    unsigned long long t_phase[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    const long long t_kernel = clock64();
    long long t_mark = t_kernel;
#define LAP(slot)                                                   \
    do {                                                            \
        const long long now__ = clock64();                          \
        t_phase[slot] += (unsigned long long)(now__ - t_mark);      \
        t_mark = now__;                                             \
    } while (0)
#else
#define LAP(slot) ((void)0)
#endif
    if (p.stagger_ns && p.group <= 1) {
        const unsigned int slot = blockIdx.x / (unsigned int)p.sm_count;
        const long long t_end = clock64() + (long long)slot * p.stagger_ns * 2; // ~2 cycles per ns
        while (clock64() < t_end)
            __nanosleep(1000);
    }
    for (int f = blockIdx.x; f < p.frames; f += gridDim.x) {
        // ---- streaming input: wait until the host->device copy of this frame's chunk has landed ----
        if (p.ready) {
            if (tid == 0) {
                const unsigned int need = (unsigned int)(f / p.ready_chunk) + 1u;
                const long long t0 = clock64();
                int gave_up = 0;
                while (*reinterpret_cast<const volatile unsigned int*>(p.ready) < need) {
                    __nanosleep(500);
                    if (clock64() - t0 > (1ll << 35)) { // ~17 s: the copy is not coming (host-side failure)
                        gave_up = 1;
                        break;
                    }
                }
                if (gave_up && p.err)
                    atomicExch(p.err, 1u);
                s_abort = gave_up;
                __threadfence();
            }
            __syncthreads();
            if (s_abort)
                break;
        }
        // ---- soft input: HBM -> shared memory, into the pair-interleaved order --------------------
        {
            const uint32_t* src = reinterpret_cast<const uint32_t*>(p.llr + (size_t)f * N);
            const int data_chunks = K / 8; // 8 output bytes per chunk = 4 halfwords
            for (int u = tid; u < N / 8; u += kLdpcThreads) {
                int ia, ib, out;
                if (u < data_chunks) {
                    const int g = u / 45, w4 = u - g * 45;
                    ia = g * 90 + w4; // word index of bytes g*360 + 4*w4 ..
                    ib = ia + 45;     // + 180 bytes
                    out = g * 360 + 8 * w4;
                } else {
                    const int up = u - data_chunks;
                    ia = K / 4 + up;
                    ib = ia + half / 4;
                    out = K + 8 * up;
                }
                const uint32_t a = ldg_hint(src + ia, pol_stream), b = ldg_hint(src + ib, pol_stream);
                *reinterpret_cast<uint2*>(L + out) = make_uint2(prmt(a, b, 0x5140), prmt(a, b, 0x7362));
            }
        }
        __syncthreads();
        LAP(0);

        // ---- while (bad() && --trials >= 0) update();  layered_decoder.hh:153 ----------------------
        int trials = p.max_trials;
        int iter = 0;
        int bad;
        int proven_bad = 0;
        for (;;) {
            if (!proven_bad) {
                uint32_t flag = 0;
                if (tid < kPairs) {
                    for (int i = 0; i < q; ++i) {
                        const LayerView lv = load_layer(layers, i);
                        flag |= check_pair<CNT_MAX, UNIFORM>(L, edges, lv, i, tid, K, q);
                    }
                }
                bad = __syncthreads_or((int)flag);
                LAP(1);
            } else {
                bad = 1;
            }
            if (p.group > 1) {
                // reference batch semantics: the whole SIMD batch keeps iterating while any of its
                // frames is bad.  One word per (group, iteration): low half counts arrivals, high
                // half counts bad frames.  All CTAs are co-resident (cooperative launch).
                if (tid == 0) {
                    unsigned int* w = p.gsync + (size_t)(f / p.group) * (p.max_trials + 2) + iter;
                    __threadfence();
                    atomicAdd(w, 1u + (bad ? 0x10000u : 0u));
                    unsigned int cur;
                    do {
                        cur = *reinterpret_cast<volatile unsigned int*>(w);
                    } while ((cur & 0xffffu) < (unsigned int)p.group);
                    s_group_bad = (cur >> 16) != 0;
                }
                __syncthreads();
                bad = s_group_bad;
                __syncthreads();
            }
            if (!bad || --trials < 0)
                break;
            const bool zero_state = (iter == 0); // reset(): layered_decoder.hh:27-31, no memset needed
            ++iter;

            // ---- one iteration: walk the step list ---------------------------------------------------
            int self_bad = 0;
            uint32_t wA = 0, wB = 0, sA = 0, sB = 0;
            // state words of a pair step, requested from L2 one step ahead of their use
            auto prefetch = [&](uint2 s) {
                if (((s.x >> 16) == 0 || ((s.y >> 30) & 1u)) && tid < kPairs && !zero_state) {
                    const uint32_t* m = msg + ((size_t)(s.x & 0xffu) * kPairs + tid) * 2 * MW;
                    if (!WIDE) {
                        const uint2 t = ldg_hint(reinterpret_cast<const uint2*>(m), pol_keep);
                        wA = t.x, wB = t.y;
                    } else {
                        const uint4 t = ldg_hint(reinterpret_cast<const uint4*>(m), pol_keep);
                        wA = t.x, wB = t.y, sA = t.z, sB = t.w;
                    }
                }
            };
            uint2 st = steps[0];
            prefetch(st);
            const int last_step = p.n_steps - 1;
            for (int s = 0; s < p.n_steps;) {
                const int layer = (int)(st.x & 0xffu), run_len = (int)((st.x >> 8) & 0xffu), count = (int)(st.x >> 16);
                const uint32_t work_off = st.y & 0x00ffffffu;
                const bool barrier_before = (st.y >> 24) & 1u, is_run = (st.y >> 25) & 1u, link_parallel = (st.y >> 29) & 1u;
                const int sub_warps = (int)((st.y >> 26) & 7u);
                const bool is_split = (st.y >> 30) & 1u, is_chain = (st.y >> 31) & 1u;
                const LayerView lv = load_layer(layers, layer);
                const int next = s + (is_run ? run_len : 1);
                const bool last = (next == p.n_steps);
                const uint32_t cA = wA, cB = wB, csA = sA, csB = sB;
                if (!last) {
                    st = steps[next];
                    prefetch(st); // the next layer's state words travel while this one computes
                }
                // a block barrier only where another thread's writes are read (code_tables.cc); every split
                // step has one (its named barriers and s_progress are reused from one to the next)
                if (barrier_before)
                    __syncthreads();
#if !DVBS2_LEGACY_WAVEFRONT
                if (is_split) {
                    if (tid == 0)
                        s_progress = 0;
                    __syncthreads();
                }
#endif
                if (count == 0) {
                    if (tid < kPairs) {
                        uint32_t* mo = msg + ((size_t)layer * kPairs + tid) * 2 * MW;
                        if (last)
                            self_bad |= process_pair<CNT_MAX, UNIFORM, WIDE, true>(L, edges, lv, layer, tid, K, q, cA, cB, csA, csB, mo, pol_keep);
                        else
                            process_pair<CNT_MAX, UNIFORM, WIDE, false>(L, edges, lv, layer, tid, K, q, cA, cB, csA, csB, mo, pol_keep);
                    }
                }
#if !DVBS2_LEGACY_WAVEFRONT
                else if (is_split) {
                    // conflict layer, pair mapping kept: private links in s16x2, shared links level by level
                    uint32_t* mo = msg + ((size_t)layer * kPairs + (tid < kPairs ? tid : 0)) * 2 * MW;
                    if (last)
                        self_bad |= process_split<CNT_MAX, UNIFORM, WIDE, true, TMEM>(L, edges, lv, layer, tid, tid < kPairs, K, q, cA, cB, csA, csB, mo,
                                                                                pol_keep, count, work + work_off, &s_progress, p.prof ? p.prof + (size_t)blockIdx.x * 16 : nullptr, rec, is_chain, run_len, (int)link_parallel);
                    else
                        process_split<CNT_MAX, UNIFORM, WIDE, false, TMEM>(L, edges, lv, layer, tid, tid < kPairs, K, q, cA, cB, csA, csB, mo, pol_keep,
                                                                     count, work + work_off, &s_progress, p.prof ? p.prof + (size_t)blockIdx.x * 16 : nullptr, rec, is_chain, run_len, (int)link_parallel);
                }
#else
                else if (is_run) {
                    // run of narrow wavefront levels on the first sub_warps warps; the other warps move on
                    if (tid < sub_warps * 32) {
                        if (link_parallel)
                            self_bad |= lp_run<WIDE>(L, edges, steps, work, msg, lv.cnt, lv.edge_begin, layer, s, run_len, sub_warps,
                                                     last_step, K, q, zero_state, pol_keep, tid);
                        else
                            self_bad |= scalar_run<CNT_MAX, UNIFORM, WIDE, TMEM>(L, edges, steps, work, msg, lv, layer, s, run_len, last_step,
                                                                          K, q, zero_state, pol_keep, tid, tcol, tmem_base);
                    }
                } else {
                    // wide wavefront level: single check nodes, one per thread
                    const uint32_t tc = tcol[s];
                    if (USE_TMEM && tc != 0xffu) {
                        // state in tensor memory: a column per 4 warps and pass, lane = thread
                        for (int t0 = 0; t0 < count; t0 += kLdpcThreads) {
                            const int t = t0 + tid;
                            if (t0 + (tid & ~31) < count) { // warp-uniform: USE_TMEM accesses are warp collectives
                                const uint32_t taddr = tmem_warp + tc + 2u * (uint32_t)(t0 / kLdpcThreads) + (uint32_t)(tid >> 7);
                                uint32_t w = zero_state ? 0u : tmem_ld(taddr);
                                if (t < count) {
                                    const int j = (int)work[work_off + t];
                                    if (last)
                                        self_bad |= process_cn<CNT_MAX, UNIFORM, WIDE, true>(L, edges, lv, layer, j, K, q, msg, zero_state, pol_keep, true, w, w >> 17, &w);
                                    else
                                        process_cn<CNT_MAX, UNIFORM, WIDE, false>(L, edges, lv, layer, j, K, q, msg, zero_state, pol_keep, true, w, w >> 17, &w);
                                }
                                tmem_st(taddr, w);
                            }
                        }
                    } else {
                        for (int t = tid; t < count; t += kLdpcThreads) {
                            const int j = (int)work[work_off + t];
                            const int pp = j >= kPairs ? j - kPairs : j;
                            uint32_t* mp = msg + ((size_t)layer * kPairs + pp) * 2 * MW;
                            if (last)
                                self_bad |= process_cn<CNT_MAX, UNIFORM, WIDE, true>(L, edges, lv, layer, j, K, q, mp, zero_state, pol_keep, false, 0, 0);
                            else
                                process_cn<CNT_MAX, UNIFORM, WIDE, false>(L, edges, lv, layer, j, K, q, mp, zero_state, pol_keep, false, 0, 0);
                        }
                    }
                }
#endif
                LAP(count == 0 ? 2 : (is_run ? 3 : 4));
                s = next;
            }
            proven_bad = __syncthreads_or(self_bad);
            LAP(5);
        }

        // ---- outputs -----------------------------------------------------------------------------------
        if (p.trials_left && tid == 0)
            p.trials_left[f] = trials;
        if (p.hard) {
            // llr < 0 -> 1, MSB first: lib/ldpc_decoder_bb_impl.cc:432-442
            uint8_t* dst = p.hard + (size_t)f * p.out_bytes;
            for (int b = tid; b < p.out_bytes; b += kLdpcThreads) {
                const int n0 = 8 * b;
                uint32_t acc = 0;
                if (n0 < K) { // 8 consecutive bits of one 360-bit group (8 | 360)
                    const int g = n0 / 360, m = n0 - g * 360;
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        acc |= (L[data_addr(g * 360, m + k)] < 0 ? 1u : 0u) << (7 - k);
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        acc |= (L[parity_addr(K, half, n0 - K + k)] < 0 ? 1u : 0u) << (7 - k);
                }
                dst[b] = (uint8_t)acc;
            }
        }
        if (p.llr_post) {
            uint32_t* dst = reinterpret_cast<uint32_t*>(p.llr_post + (size_t)f * N);
            const int data_chunks = K / 8;
            for (int u = tid; u < N / 8; u += kLdpcThreads) {
                int ia, ib, in;
                if (u < data_chunks) {
                    const int g = u / 45, w4 = u - g * 45;
                    ia = g * 90 + w4;
                    ib = ia + 45;
                    in = g * 360 + 8 * w4;
                } else {
                    const int up = u - data_chunks;
                    ia = K / 4 + up;
                    ib = ia + half / 4;
                    in = K + 8 * up;
                }
                const uint2 w = *reinterpret_cast<const uint2*>(L + in);
                dst[ia] = prmt(w.x, w.y, 0x6420);
                dst[ib] = prmt(w.x, w.y, 0x7531);
            }
        }
        __syncthreads(); // L is reused by the next frame
        LAP(6);
    }
    if (USE_TMEM) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid < 32)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemColsDev) : "memory");
    }
#ifdef DVBS2_PHASE_PROFILE
    if (p.prof && tid == 0) {
        t_phase[7] = (unsigned long long)(clock64() - t_kernel);
        for (int k = 0; k < 8; ++k)
            p.prof[(size_t)blockIdx.x * 16 + k] = t_phase[k];
    }
#endif
}

template <int CNT_MAX, bool UNIFORM, bool WIDE, bool TMEM>
cudaError_t launch_one(const LdpcLaunch& p, int grid, size_t smem, cudaStream_t stream)
{
    auto kern = ldpc_decode_kernel<CNT_MAX, UNIFORM, WIDE, TMEM>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess)
        return e;
    if (p.group > 1) {
        void* args[] = { (void*)&p };
        return cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(kLdpcThreads), args, smem, stream);
    }
    kern<<<grid, kLdpcThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

template <int CNT_MAX, bool UNIFORM, bool WIDE, bool TMEM>
int occupancy_one(size_t smem)
{
    auto kern = ldpc_decode_kernel<CNT_MAX, UNIFORM, WIDE, TMEM>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return 0;
    if (!(TMEM && DVBS2_LEGACY_WAVEFRONT)) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, kLdpcThreads, smem) != cudaSuccess)
            return 0;
        return n;
    }
    // The occupancy calculator answers 1 for any kernel that contains tcgen05 instructions; the SM does
    // co-schedule CTAs that each allocate a share of the 512 TMEM columns (measured: 3 CTAs x 128 columns
    // run 1.6x faster than 1).  Compute the residency from the kernel's real resources instead.
    cudaFuncAttributes fa;
    int dev = 0, smem_sm = 0, regs_sm = 0, threads_sm = 0;
    if (cudaFuncGetAttributes(&fa, kern) != cudaSuccess || cudaGetDevice(&dev) != cudaSuccess)
        return 0;
    cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev);
    cudaDeviceGetAttribute(&threads_sm, cudaDevAttrMaxThreadsPerMultiProcessor, dev);
    const int regs_per_warp = ((fa.numRegs * 32 + 255) / 256) * 256;
    const int by_regs = regs_sm / (regs_per_warp * (kLdpcThreads / 32));
    const int by_smem = (int)((size_t)smem_sm / (smem + fa.sharedSizeBytes + 1024));
    const int by_threads = threads_sm / kLdpcThreads;
    const int by_tmem = 512 / kTmemColsDev;
    return std::max(0, std::min(std::min(by_regs, by_smem), std::min(by_threads, by_tmem)));
}

} // namespace

#if !DVBS2_LEGACY_WAVEFRONT // shared helpers live in one of the two translation units
size_t ldpc_smem_bytes(int N, uint32_t tab_bytes, bool chain_scratch, LdpcLaunch* p)
{
    size_t off = ((size_t)N + 15) & ~(size_t)15;
    if (p)
        p->smem_tab_off = (uint32_t)off;
    off += tab_bytes;
    off = (off + 15) & ~(size_t)15;
    if (p)
        p->smem_bar_off = (uint32_t)off;
    off += 16;
    if (p)
        p->smem_rec_off = chain_scratch ? (uint32_t)off : 0u;
    if (chain_scratch)
        off += 360 * 8;
    return off;
}

bool ldpc_wide_state(int max_cnt) { return max_cnt > 13; }
#endif

// Kernel instantiations.  Codes whose layers all have the same number of data links per check node
// (every DVB-S2 normal-frame table) get the link count as a compile-time constant: no predication,
// no dead link slots.  The rest take the predicated variant of the next size up.  The narrow state
// word holds 15 sign bits (<= 13 data links), above that the wide (two-word) state is used.
// The fourth template flag is "wavefront state in tensor memory" in the wavefront build and "level-form split
// steps through out-of-line calls" in the split build (narrow-state codes only in both).
#define DVBS2_TM(CALL, C, U) (tmem ? CALL(C, U, false, true) : CALL(C, U, false, false))
#define DVBS2_DISPATCH(CALL)                                   \
    if (uniform) {                                             \
        switch (max_cnt) {                                     \
        case 2: return DVBS2_TM(CALL, 2, true);                \
        case 3: return DVBS2_TM(CALL, 3, true);                \
        case 4: return DVBS2_TM(CALL, 4, true);                \
        case 5: return DVBS2_TM(CALL, 5, true);                \
        case 7: return DVBS2_TM(CALL, 7, true);                \
        case 8: return DVBS2_TM(CALL, 8, true);                \
        case 9: return DVBS2_TM(CALL, 9, true);                \
        case 11: return DVBS2_TM(CALL, 11, true);              \
        case 12: return DVBS2_TM(CALL, 12, true);              \
        case 16: return CALL(16, true, true, false);           \
        case 20: return CALL(20, true, true, false);           \
        case 25: return CALL(25, true, true, false);           \
        case 28: return CALL(28, true, true, false);           \
        default: break;                                        \
        }                                                      \
    }                                                          \
    if (max_cnt <= 5) return DVBS2_TM(CALL, 5, false);         \
    if (max_cnt <= 9) return DVBS2_TM(CALL, 9, false);         \
    if (max_cnt <= 13) return DVBS2_TM(CALL, 13, false);       \
    if (max_cnt <= 18) return CALL(18, false, true, false);    \
    if (max_cnt <= 28) return CALL(28, false, true, false);

cudaError_t LDPC_SYM(ldpc_launch)(const LdpcLaunch& p, int max_cnt, bool uniform, bool tmem, int grid, size_t smem, cudaStream_t stream)
{
#define CALL(C, U, W, T) launch_one<C, U, W, T>(p, grid, smem, stream)
    DVBS2_DISPATCH(CALL)
#undef CALL
    return cudaErrorInvalidValue;
}

int LDPC_SYM(ldpc_ctas_per_sm)(int max_cnt, bool uniform, bool tmem, size_t smem)
{
#define CALL(C, U, W, T) occupancy_one<C, U, W, T>(smem)
    DVBS2_DISPATCH(CALL)
#undef CALL
    return 0;
}

} // namespace dvbs2b200
