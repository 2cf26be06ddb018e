// ldpc_kernel.cu -- layered offset-min-sum LDPC decoder for DVB-S2/S2X/T2 on sm_100a.
//
// Arithmetic contract (bit-exact with the reference CPU path):
//   lib/ldpc_decoder/layered_decoder.hh:27-160  schedule, iteration control
//   lib/ldpc_decoder/algorithms.hh:151-207      OffsetMinSumAlgorithm<int8>, beta = 1,
//                                               stored messages clamped to [-32, 31]
//   lib/ldpc_decoder_bb_impl.cc:432-442         hard decision + MSB-first packing
//
// Design (B200-first, not a translation of the SIMD-across-frames CPU code):
//   * one FECFRAME per CTA of 192 threads, persistent over the batch; three CTAs per SM for normal frames (two
//     for 28 data links per check node), four for short frames;
//   * the frame's N posteriors stay in shared memory for the whole decode as biased bytes, in the
//     pair-interleaved order of code_tables.h: check nodes p and p+180 of a layer read / write ONE aligned 16-bit
//     word per link, so a thread runs TWO check nodes in the halves of a 32-bit register (s16x2);
//   * what a thread does per step is in ldpc_steps.cuh / ldpc_core.cuh (the same code runs on the CPU in
//     tools/ldpc_emul.cc); this file is the frame loop, the barriers between the phases of a step and the
//     traffic: tables staged by one TMA bulk copy per CTA, the compressed check-node state (8 bytes per check-node
//     pair and layer for up to 8 links) in an L2-resident scratch, read one step ahead of its use, the soft
//     input read once, the packed hard decisions (and optionally the posteriors) written once;
//   * layers whose circulants share a 360-bit group are order sensitive in the reference (it visits check nodes
//     serially): they run as split steps (code_tables.h) -- same result, bit for bit;
//   * the syndrome test after an iteration is skipped when the check nodes of the final step, whose posteriors are
//     final by then, already prove the frame is still bad.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "kernels.h"
#include "ldpc_steps.cuh"

namespace dvbs2b200 {

namespace {

using namespace core;

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
__device__ __forceinline__ void stg_hint(uint32_t* p, uint32_t v, uint64_t pol)
{
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void stg_hint(uint2* p, uint2 v, uint64_t pol)
{
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.u32 [%0], {%1,%2}, %3;" ::"l"(p), "r"(v.x), "r"(v.y), "l"(pol) : "memory");
}

// byte address of element s of a 360-bit group / of parity bit c in the pair-interleaved layout
__device__ __forceinline__ int data_addr(int group_base, int s) { return group_base + 2 * (s >= kPairs ? s - kPairs : s) + (s >= kPairs); }
__device__ __forceinline__ int parity_addr(int K, int half, int c) { return K + 2 * (c >= half ? c - half : c) + (c >= half); }

// State scratch of a CTA: plane 0 holds {minima word, first field word} of every (layer, pair) as uint2, plane 1
// the remaining field words of codes with more than 8 links per check node.
template <int NW>
struct StateIo {
    uint32_t* base;
    int q;
    uint64_t pol;
    __device__ __forceinline__ void load(int layer, int p, RawState<NW>& s) const
    {
        const uint2 t = ldg_hint(reinterpret_cast<const uint2*>(base) + (size_t)layer * kPairs + p, pol);
        s.Cw = t.x;
        s.W[0] = t.y;
        const uint32_t* more = base + (size_t)q * kPairs * 2 + ((size_t)layer * kPairs + p) * (NW > 1 ? NW - 1 : 1);
#pragma unroll
        for (int w = 1; w < NW; ++w)
            s.W[w] = ldg_hint(more + w - 1, pol);
    }
    __device__ __forceinline__ void store(int layer, int p, const RawState<NW>& s) const
    {
        stg_hint(reinterpret_cast<uint2*>(base) + (size_t)layer * kPairs + p, make_uint2(s.Cw, s.W[0]), pol);
        uint32_t* more = base + (size_t)q * kPairs * 2 + ((size_t)layer * kPairs + p) * (NW > 1 ? NW - 1 : 1);
#pragma unroll
        for (int w = 1; w < NW; ++w)
            stg_hint(more + w - 1, s.W[w], pol);
    }
};

// ---- level form of a split step: the levels of the serial order -------------------------------------------
// The nodes of a level are independent and a level is a range of j (code_tables.h).  A node takes W = 2, 4, 8 or 16
// neighbouring lanes, one per shared link (ldpc_steps.cuh: level_link_load / level_link_store); what it needs waits in
// shared memory (level_prep), so it does not matter which lanes run it.  One block barrier per level.
__device__ __forceinline__ void level_phase(uint8_t* L, int nshared, int depth, const LevelScratch& ls)
{
    const int logw = nshared <= 2 ? 1 : nshared <= 4 ? 2 : nshared <= 8 ? 3 : 4;
    const int lane = (int)threadIdx.x & 31, s = lane & ((1 << logw) - 1);
    const unsigned gmask = ((1u << (1 << logw)) - 1u) << (lane - s); // the lanes of this node
    const int slots = kLdpcThreads >> logw, slot = (int)threadIdx.x >> logw;
    int j0 = (int)ls.first_node[1], j1 = (int)ls.first_node[2];
    int rot = 0; // the warp that takes the first nodes of a level rotates: the serial phase loads all four schedulers
    for (int lvl = 1; lvl <= depth; ++lvl) {
        const int j2 = (int)ls.first_node[min(lvl + 2, depth + 1)];
        int g = slot - rot;
        g += g < 0 ? slots : 0;
        const int g_warp = g - (lane >> logw); // the warp's first slot: its slots do not wrap (rot is a multiple of them)
        for (int base = j0; base < j1; base += slots) {
            if (base + g_warp >= j1)
                continue; // no node for this warp: straight to the barrier (warp-uniform, the shuffles below stay converged)
            const int j = base + g;
            const bool mine = j < j1 && s < nshared;
            const int hs = j >= kPairs ? 1 : 0, p = j - kPairs * hs;
            LevelLink k;
            k.key = kLevelNoKey;
            k.xb = 255;
            if (mine)
                level_link_load(L, ls, p, hs, s, k);
            // two smallest keys of the node: butterfly over its lanes, (smallest | second << 16) per lane.  (The
            // warp-reduce instructions serialise sub-warp groups; shuffles do not care about groups.)
            uint32_t v = k.key | 0x7fff0000u;
#pragma unroll
            for (int st = 0; st < 4; ++st) {
                if (st < logw) {
                    const uint32_t o = __shfl_xor_sync(0xffffffffu, v, 1 << st);
                    const uint32_t mn = __vmins2(v, o), mx = __vmaxs2(v, o);
                    v = (mn & 0xffffu) | (min(mx & 0xffffu, mn >> 16) << 16);
                }
            }
            const uint32_t k0 = v & 0xffffu, k1 = v >> 16;
            const uint32_t negs = (uint32_t)__popc(__ballot_sync(0xffffffffu, k.xb < 128) & gmask);
            if (mine)
                level_link_store(L, ls, p, hs, s, k, k0, k1, negs);
        }
        __syncthreads();
        j0 = j1;
        j1 = j2;
        rot += 32 >> logw;
        rot = rot == slots ? 0 : rot;
    }
}

// ---- chain form of a split step: lane c < delta walks the chain c, c + delta, ... ---------------------------
// The bit that a node updates through its out link is the bit the next node reads through its in link: it travels
// in a register.  A chain's last node meets, through its out link, the bit that the first node of some chain updated
// through its in link at step 0 -- hence one barrier among the walking warps after step 0, and none after that.
// A node of the walk is ~25 instructions: everything that does not depend on the carried bit was put into the node
// records by phase 1, everything that is not needed to carry it on (the node's own minima, signs, the in-link update)
// is redone in s16x2 by phase 3 with the inputs the walk leaves in lin[].
__device__ __forceinline__ void chain_walk(uint8_t* __restrict__ L, const ChainRec* __restrict__ rec, uint8_t* __restrict__ lin, int delta, int tid)
{
    const int nwarps = (delta + 31) >> 5; // called by every lane of the first nwarps warps
    const bool live = tid < delta;
    int carried = 0;
    if (live) {
        const ChainRec cr = rec[tid];
        const int l_in = (int)L[cr.x & 0xffffu], l_out = (int)L[cr.x >> 16];
        lin[kLinStride * tid] = (uint8_t)l_in;
        carried = chain_node(L, cr, l_in, l_out, true);
    }
    if (nwarps == 1)
        __syncwarp();
    else
        asm volatile("bar.sync 1, %0;" ::"r"(nwarps * 32) : "memory");
    int j = tid + delta;
    if (live && j < 360) {
        ChainRec cr = rec[j];
        int lo = (int)L[cr.x >> 16];
        for (;;) {
            const int jn = j + delta;
            ChainRec crn = cr;
            int lon = lo;
            if (jn < 360) { // what the next node needs that does not depend on this one
                crn = rec[jn];
                lon = (int)L[crn.x >> 16];
            }
            lin[kLinStride * j] = (uint8_t)carried;
            carried = chain_node(L, cr, carried, lo, false);
            if (jn >= 360)
                break;
            j = jn;
            cr = crn;
            lo = lon;
        }
    }
}

constexpr int kShortFrameBits = 16200;

#ifndef DVBS2_THREE_CTA_UPTO
#define DVBS2_THREE_CTA_UPTO 25
#endif
#ifndef DVBS2_FOUR_CTA_UPTO
#define DVBS2_FOUR_CTA_UPTO 12 // wider check nodes spill too much at 80 registers
#endif
#ifndef DVBS2_SHORT_FRAME_REGS
#define DVBS2_SHORT_FRAME_REGS 80
#endif
// resident CTAs per SM the kernels are compiled for: three CTAs are 18 warps = 5 on some scheduler, whose register file
// holds 16384 registers: 96 per thread at most.  Short frames (SMALL: N <= 16200) leave shared memory for more CTAs and
// spend more of their time in split steps (barrier waits), so they are compiled for four CTAs per SM: 80 registers.
constexpr int ldpc_regs(int cnt_max, bool small_frame)
{
    return cnt_max > DVBS2_THREE_CTA_UPTO ? 168 : (small_frame && cnt_max <= DVBS2_FOUR_CTA_UPTO) ? DVBS2_SHORT_FRAME_REGS : 96;
}

template <int CNT_MAX, bool UNIFORM, bool SMALL>
__global__ void __launch_bounds__(kLdpcThreads) __maxnreg__(ldpc_regs(CNT_MAX, SMALL)) ldpc_decode_kernel(const LdpcLaunch p)
{
    constexpr int NW = (CNT_MAX + 2 + 7) / 8;
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* const L = smem;
    const LayerRec* layers = reinterpret_cast<const LayerRec*>(smem + p.smem_tab_off);
    const EdgeRec* edges = reinterpret_cast<const EdgeRec*>(smem + p.smem_tab_off + (size_t)p.q * sizeof(LayerRec));
    const uint2* steps = reinterpret_cast<const uint2*>(smem + p.smem_tab_off + (size_t)p.q * sizeof(LayerRec) + (size_t)p.n_circ * sizeof(EdgeRec));
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + p.smem_bar_off);
    ChainRec* const rec = reinterpret_cast<ChainRec*>(smem + p.smem_rec_off); // node records of chain-form split steps
    uint8_t* const lin = smem + p.smem_rec_off; // byte 0 of a consumed node record (ldpc_steps.cuh)
    __shared__ int s_group_bad;
    __shared__ int s_abort;

    const int tid = threadIdx.x;
    const int N = p.N, K = p.K, q = p.q, R = p.R;
    const int half = R / 2;
    const bool active = tid < kPairs;
    const uint16_t* __restrict__ work = p.work;
    const uint64_t pol_keep = policy_evict_last(), pol_stream = policy_evict_first();
    StateIo<NW> sio = { p.msg_scratch + (size_t)blockIdx.x * (size_t)half * (NW + 1), q, pol_keep };
    ThreadConst tc;
    tc.p = (uint32_t)(active ? tid : 0);
    tc.two = p.two;
    tc.four = p.four;
    tc.c30 = p.c30;
    tc.c16 = p.c16;
    tc.c32 = p.c32;
    tc.neg1 = p.neg1;
    const FrameCtx ctx = { L, layers, edges, K, q };

    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // stage the code tables once per CTA (TMA)
    if (tid == 0) {
        mbar_expect_tx(bar, p.tab_bytes);
        tma_load_1d(smem + p.smem_tab_off, p.tab, p.tab_bytes, bar);
    }
    mbar_wait(bar, 0);

#ifdef DVBS2_PHASE_PROFILE
    // diagnostics build: cycles per phase, per CTA (thread 0) -> p.prof[blockIdx][16]:
    // 0 load, 1 syndrome pass, 2 pair steps, 3 split: phase 1, 4 split: serial phase, 5 split: phase 3, 6 iteration end,
    // 7 output, 8 total, 9 pair steps (count), 10 split steps (count), 11 level-form serial phase as the thread of node
    // 359 sees it (entry to the last level done; written by thread 179), 12 levels walked (count), 13 chain-form serial
    // phase (between its two block barriers), 14 chain nodes per walker (count)
    unsigned long long t_phase[16] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
    const long long t_kernel = clock64();
    long long t_mark = t_kernel;
#define LAP(slot)                                                   \
    do {                                                            \
        const long long now__ = clock64();                          \
        t_phase[slot] += (unsigned long long)(now__ - t_mark);      \
        t_mark = now__;                                             \
    } while (0)
#define COUNT(slot) (t_phase[slot] += 1)
#define COUNTN(slot, n) (t_phase[slot] += (unsigned long long)(n))
#else
#define LAP(slot) ((void)0)
#define COUNT(slot) ((void)0)
#define COUNTN(slot, n) ((void)0)
#endif
    __shared__ int s_next_frame;
    int f = blockIdx.x;
    while (f < p.frames) {
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
        // ---- soft input: HBM -> shared memory, biased, into the pair-interleaved order ----------------
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
                const uint32_t a = ldg_hint(src + ia, pol_stream) ^ 0x80808080u, b = ldg_hint(src + ib, pol_stream) ^ 0x80808080u;
                *reinterpret_cast<uint2*>(L + out) = make_uint2(prmt(a, b, 0x5140u), prmt(a, b, 0x7362u));
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
                // bad(): any unsatisfied check.  A frame that is still bad shows it within a few layers, so the pass
                // goes over the layers in chunks and stops at the first chunk with an unsatisfied check.
#ifndef DVBS2_SYNDROME_CHUNK
#define DVBS2_SYNDROME_CHUNK 4
#endif
                constexpr int kSyndromeChunk = DVBS2_SYNDROME_CHUNK;
                bad = 0;
#pragma unroll 1
                for (int i0 = 0; i0 < q && !bad; i0 += kSyndromeChunk) {
                    int flag = 0;
                    if (active) {
                        const int i1 = min(i0 + kSyndromeChunk, q);
#pragma unroll 1
                        for (int i = i0; i < i1; ++i)
                            flag |= check_pair<CNT_MAX, UNIFORM>(ctx, tc, i);
                    }
                    bad = __syncthreads_or(flag);
                }
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
            RawState<NW> pre; // state words of the next step, requested from L2 one step ahead of their use
            pre.Cw = 0u;
#pragma unroll
            for (int w = 0; w < NW; ++w)
                pre.W[w] = 0u;
            uint2 st = steps[0];
            if (active && !zero_state)
                sio.load((int)(st.x & 0xffu), tid, pre);
            for (int s = 0; s < p.n_steps; ++s) {
                const int layer = (int)(st.x & 0xffu), delta = (int)((st.x >> 8) & 0xffu), depth = (int)(st.x >> 16);
                const uint32_t work_off = st.y & kStepOffMask;
                const bool barrier_before = (st.y & kStepBarrierBefore) != 0, is_split = depth != 0;
                const bool is_chain = (st.y & kStepChain) != 0;
                const int out_link1 = (st.y & kStepChainOutLink1) ? 1 : 0;
                const bool last = (s + 1 == p.n_steps);
                const RawState<NW> cur = pre;
                if (!last) {
                    st = steps[s + 1];
                    if (active && !zero_state)
                        sio.load((int)(st.x & 0xffu), tid, pre); // the next layer's state travels while this one computes
                }
                // a block barrier only where another thread's writes are read (code_tables.cc)
                if (barrier_before)
                    __syncthreads();
                RawState<NW> out;
                if (!is_split) {
                    if (active) {
                        if (last)
                            self_bad |= pair_step<CNT_MAX, UNIFORM, true, NW>(ctx, tc, layer, cur, out);
                        else
                            pair_step<CNT_MAX, UNIFORM, false, NW>(ctx, tc, layer, cur, out);
                        sio.store(layer, tid, out);
                    }
                    LAP(2);
                    COUNT(9);
                } else {
                    SplitRegs<CNT_MAX, NW> r;
                    if (active)
                        split_p1<CNT_MAX, UNIFORM, NW>(ctx, tc, layer, cur, r);
                    if (is_chain) {
                        if (active)
                            chain_p1<CNT_MAX, UNIFORM, NW>(ctx, tc, layer, out_link1, rec, r);
                        __syncthreads();
                        LAP(3);
                        if ((tid >> 5) < ((delta + 31) >> 5))
                            chain_walk(L, rec, lin, delta, tid);
                        __syncthreads();
                        LAP(13);
                        COUNTN(14, depth);
                        if (active)
                            chain_p3_links<CNT_MAX, NW>(ctx, tc, out_link1, delta, lin, r);
                    } else {
                        const int nshared = (int)layers[layer].conflict;
                        const LevelScratch ls = level_scratch(rec, nshared);
                        if (active)
                            level_prep<CNT_MAX, NW>(ctx, tc, layer, ls, r);
                        for (int i = tid; i < depth + 2; i += kLdpcThreads)
                            ls.first_node[i] = __ldg(work + work_off + 360 + i);
                        __syncthreads();
                        LAP(3);
                        level_phase(L, nshared, depth, ls);
                        LAP(4);
                        COUNTN(12, depth);
                        if (active)
                            level_p3_links<CNT_MAX, NW>(tc, nshared, ls, r);
                    }
                    if (active) {
                        Final<NW> fin;
                        split_p3<CNT_MAX, UNIFORM, NW>(ctx, tc, layer, r, fin, out);
                        if (is_chain) {
                            uint32_t syn = 0u, zer = 0u;
                            chain_p3_store<CNT_MAX, NW>(ctx, tc, out_link1, delta, fin, r, syn, zer, false);
                        }
                        sio.store(layer, tid, out);
                    }
                    if (last) {
                        // the layer's posteriors are final once every thread has stored: the syndrome test of this one
                        // layer can prove the frame still bad and save the pass over all layers
                        __syncthreads();
                        if (active)
                            self_bad |= check_pair<CNT_MAX, UNIFORM>(ctx, tc, layer);
                    }
                    LAP(5);
                    COUNT(10);
                }
            }
            proven_bad = __syncthreads_or(self_bad);
            LAP(6);
        }

        // ---- outputs -----------------------------------------------------------------------------------
        if (p.trials_left && tid == 0)
            p.trials_left[f] = trials;
        if (p.hard) {
            // llr < 0 -> 1, MSB first: lib/ldpc_decoder_bb_impl.cc:432-442 (biased: bit 7 clear)
            uint8_t* dst = p.hard + (size_t)f * p.out_bytes;
            for (int b = tid; b < p.out_bytes; b += kLdpcThreads) {
                const int n0 = 8 * b;
                uint32_t acc = 0;
                if (n0 < K) { // 8 consecutive bits of one 360-bit group (8 | 360)
                    const int g = n0 / 360, m = n0 - g * 360;
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        acc |= (uint32_t)((L[data_addr(g * 360, m + k)] >> 7) ^ 1u) << (7 - k);
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        acc |= (uint32_t)((L[parity_addr(K, half, n0 - K + k)] >> 7) ^ 1u) << (7 - k);
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
                dst[ia] = prmt(w.x, w.y, 0x6420u) ^ 0x80808080u;
                dst[ib] = prmt(w.x, w.y, 0x7531u) ^ 0x80808080u;
            }
        }
        __syncthreads(); // L is reused by the next frame
        LAP(7);
        if (p.next_frame) { // next frame: whichever is next in the batch (frames take different numbers of iterations)
            if (tid == 0)
                s_next_frame = (int)gridDim.x + (int)atomicAdd(p.next_frame, 1u);
            __syncthreads();
            f = s_next_frame;
        } else {
            f += gridDim.x;
        }
    }
#ifdef DVBS2_PHASE_PROFILE
    if (p.prof && tid == 0) {
        t_phase[8] = (unsigned long long)(clock64() - t_kernel);
        for (int k = 0; k < 16; ++k)
            if (k != 11)
                p.prof[(size_t)blockIdx.x * 16 + k] = t_phase[k];
    }
    if (p.prof && tid == kPairs - 1)
        p.prof[(size_t)blockIdx.x * 16 + 11] = t_phase[11];
#endif
}

template <int CNT_MAX, bool UNIFORM, bool SMALL>
cudaError_t launch_one(const LdpcLaunch& p, int grid, size_t smem, cudaStream_t stream)
{
    auto kern = ldpc_decode_kernel<CNT_MAX, UNIFORM, SMALL>;
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

template <int CNT_MAX, bool UNIFORM, bool SMALL>
int occupancy_one(size_t smem)
{
    auto kern = ldpc_decode_kernel<CNT_MAX, UNIFORM, SMALL>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return 0;
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, kLdpcThreads, smem) != cudaSuccess)
        return 0;
    return n;
}

} // namespace

size_t ldpc_smem_bytes(int N, uint32_t tab_bytes, uint32_t scratch_bytes, LdpcLaunch* p)
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
        p->smem_rec_off = (uint32_t)off;
    off += scratch_bytes; // split steps: node records of the chain form / shared-link operands of the level form
    return off;
}

// Kernel instantiations.  Codes whose layers all have the same number of data links per check node
// (every DVB-S2 normal-frame table) get the link count as a compile-time constant: no predication,
// no dead link slots.  The rest take the predicated variant of the next size up.
#define DVBS2_DISPATCH(CALL)                                   \
    if (uniform) {                                             \
        switch (max_cnt) {                                     \
        case 2: return CALL(2, true);                          \
        case 3: return CALL(3, true);                          \
        case 4: return CALL(4, true);                          \
        case 5: return CALL(5, true);                          \
        case 7: return CALL(7, true);                          \
        case 8: return CALL(8, true);                          \
        case 9: return CALL(9, true);                          \
        case 11: return CALL(11, true);                        \
        case 12: return CALL(12, true);                        \
        case 16: return CALL(16, true);                        \
        case 20: return CALL(20, true);                        \
        case 25: return CALL(25, true);                        \
        case 28: return CALL(28, true);                        \
        default: break;                                        \
        }                                                      \
    }                                                          \
    if (max_cnt <= 5) return CALL(5, false);                   \
    if (max_cnt <= 9) return CALL(9, false);                   \
    if (max_cnt <= 13) return CALL(13, false);                 \
    if (max_cnt <= 18) return CALL(18, false);                 \
    if (max_cnt <= 28) return CALL(28, false);

cudaError_t ldpc_launch(const LdpcLaunch& p, int max_cnt, bool uniform, int grid, size_t smem, cudaStream_t stream)
{
    const bool small_frame = p.N <= kShortFrameBits;
#define CALL(C, U) (small_frame ? launch_one<C, U, true>(p, grid, smem, stream) : launch_one<C, U, false>(p, grid, smem, stream))
    DVBS2_DISPATCH(CALL)
#undef CALL
    return cudaErrorInvalidValue;
}

int ldpc_ctas_per_sm(int N, int max_cnt, bool uniform, size_t smem)
{
    const bool small_frame = N <= kShortFrameBits;
#define CALL(C, U) (small_frame ? occupancy_one<C, U, true>(smem) : occupancy_one<C, U, false>(smem))
    DVBS2_DISPATCH(CALL)
#undef CALL
    return 0;
}

} // namespace dvbs2b200
