// ldpc_kernel.cu -- layered offset-min-sum LDPC decoder for DVB-S2/S2X/T2 on sm_100a.
//
// Arithmetic contract (bit-exact with the reference CPU path):
//   lib/ldpc_decoder/layered_decoder.hh:27-160  schedule, iteration control
//   lib/ldpc_decoder/algorithms.hh:151-207      OffsetMinSumAlgorithm<int8>, beta = 1,
//                                               stored messages clamped to [-32, 31]
//   lib/ldpc_decoder_bb_impl.cc:432-442         hard decision + MSB-first packing
//
// Design (B200-first, not a translation of the SIMD-across-frames CPU code):
//   * one FECFRAME per CTA, persistent CTAs (one per SM) striding over the frames of a launch;
//   * the frame's N int8 posteriors live in shared memory in codeword order for the whole
//     decode; the soft input arrives with ONE bulk async copy (TMA, cp.async.bulk) per frame,
//     HBM is touched again only to write the packed hard decisions (and, optionally, the
//     posterior LLRs);
//   * check->variable messages are never stored per edge: a check node's `deg` int8 messages
//     are a function of {min0, min1, argmin, one sign bit per edge}, which is kept as one
//     32-bit (deg <= 15) or 64-bit word per check node in shared memory -- lossless w.r.t.
//     the reference's clamp-on-store, and what makes 64800 + 4*32400 bytes fit one SM;
//   * the code's circulant table (layer records + edge words) is staged into shared memory
//     with one TMA bulk copy per CTA;
//   * one thread per check node of a layer (360 of the 384 threads); layers whose circulants
//     share a 360-bit group are order sensitive in the reference (serial j), so they run as
//     precomputed wavefront steps (code_tables.cc:build_schedule) -- same result, bit for bit.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace dvbs2b200 {

namespace {

constexpr int kM = 360;

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
// TMA bulk copy shared -> global
__device__ __forceinline__ void tma_store_1d(void* dst_gmem, const void* src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ int clamp8(int x) { return min(max(x, -128), 127); }

struct LayerView {
    uint32_t edge_begin;
    int cnt;
    int n_steps;
    uint32_t step_begin;
    uint32_t order_begin;
};

__device__ __forceinline__ LayerView load_layer(const uint4* layers, int i)
{
    uint4 r = layers[i];
    LayerView v;
    v.edge_begin = r.x;
    v.cnt = (int)(r.y & 0xffffu);
    v.n_steps = (int)(r.y >> 16);
    v.step_begin = r.z;
    v.order_begin = r.w;
    return v;
}

// data-bit index read by check node j through edge word e (code_tables.h:pack_edge)
__device__ __forceinline__ int edge_index(uint32_t e, int j)
{
    int a = (int)(e >> 17);
    int idx = (int)(e & 0x1ffffu) + j;
    return (j >= a) ? idx - kM : idx;
}

// Compressed check-node state.
//   bits  0..5   min(min0, 32)     bits 6..11  min(min1, 32)     bits 12..16 argmin link
//   MSG_WORDS == 1: sign bit of link d at bit 17 + d       (deg <= 15)
//   MSG_WORDS == 2: sign bits in the second word            (deg <= 32)
template <int MSG_WORDS>
struct CnState {
    uint32_t lo, hi;
    __device__ __forceinline__ uint32_t signs() const { return MSG_WORDS == 1 ? (lo >> 17) : hi; }
};

template <int MSG_WORDS>
__device__ __forceinline__ CnState<MSG_WORDS> load_state(const uint32_t* msg, int cn)
{
    CnState<MSG_WORDS> s;
    if (MSG_WORDS == 1) {
        s.lo = msg[cn];
        s.hi = 0;
    } else {
        uint2 w = reinterpret_cast<const uint2*>(msg)[cn];
        s.lo = w.x;
        s.hi = w.y;
    }
    return s;
}

// One check-node update: lib/ldpc_decoder/layered_decoder.hh:57-76 + algorithms.hh:170-206.
// Link order inside a check node is irrelevant to the result (min0/min1/sign-xor are
// symmetric; a tie on the minimum gives min1 == min0), so parity links come first here.
template <int CNT_MAX, int MSG_WORDS>
__device__ __forceinline__ void process_cn(int8_t* __restrict__ L, uint32_t* __restrict__ msg, const uint32_t* __restrict__ edges,
                                           const LayerView& lv, int layer, int j, int K, int q)
{
    constexpr int DEG_MAX = CNT_MAX + 2;
    const int cn = kM * layer + j;
    const int c = q * j + layer; // parity bit of this check in codeword order
    const CnState<MSG_WORDS> st = load_state<MSG_WORDS>(msg, cn);
    const int old_min0 = (int)(st.lo & 63u);
    const int old_min1 = (int)((st.lo >> 6) & 63u);
    const int old_arg = (int)((st.lo >> 12) & 31u);
    const uint32_t old_signs = st.signs();

    int adr[DEG_MAX];
    int v[DEG_MAX];
    // link 0: own parity bit; link 1: previous parity bit of the zig-zag (absent for check 0)
    adr[0] = K + c;
    adr[1] = K + c - 1;
    const bool has_prev = (c > 0);
#pragma unroll
    for (int d = 0; d < CNT_MAX; ++d)
        adr[d + 2] = (d < lv.cnt) ? edge_index(edges[lv.edge_begin + d], j) : 0;

    int min0 = 127, min1 = 127, arg = 0;
    int sx = 0;
#pragma unroll
    for (int d = 0; d < DEG_MAX; ++d) {
        const bool live = (d == 0) || (d == 1 ? has_prev : (d - 2 < lv.cnt));
        if (live) {
            const int l = (int)L[adr[d]];
            // stored message = clamp(+-m, -32, 31) rebuilt from the compressed state
            const int mc = (d == old_arg) ? old_min1 : old_min0;
            const int old = ((old_signs >> d) & 1u) ? -mc : min(mc, 31);
            const int x = clamp8(l - old);          // vqsub
            v[d] = x;
            sx ^= x;
            const int mag = max(min(abs(x), 127) - 1, 0); // vqabs, then unsigned vqsub beta=1
            if (mag < min0) {
                min1 = min0;
                min0 = mag;
                arg = d;
            } else {
                min1 = min(min1, mag);
            }
        } else {
            v[d] = 0;
        }
    }
    uint32_t new_signs = 0;
#pragma unroll
    for (int d = 0; d < DEG_MAX; ++d) {
        const bool live = (d == 0) || (d == 1 ? has_prev : (d - 2 < lv.cnt));
        if (live) {
            const int m = (d == arg) ? min1 : min0;
            const bool neg = ((sx ^ v[d]) < 0); // product of the OTHER signs, zero counts as +
            const int out = neg ? -m : m;
            L[adr[d]] = (int8_t)clamp8(v[d] + out); // vqadd with the unclamped message
            new_signs |= (neg ? 1u : 0u) << d;
        }
    }
    const uint32_t lo = (uint32_t)min(min0, 32) | ((uint32_t)min(min1, 32) << 6) | ((uint32_t)arg << 12);
    if (MSG_WORDS == 1)
        msg[cn] = lo | (new_signs << 17);
    else
        reinterpret_cast<uint2*>(msg)[cn] = make_uint2(lo, new_signs);
}

// lib/ldpc_decoder/layered_decoder.hh:32-49 for one check node: unsatisfied if the sign product
// is not +, and a zero LLR counts as unsatisfied (vsign(.,0) = 0, test is "> 0").
template <int CNT_MAX>
__device__ __forceinline__ int check_cn(const int8_t* __restrict__ L, const uint32_t* __restrict__ edges, const LayerView& lv,
                                        int layer, int j, int K, int q)
{
    const int c = q * j + layer;
    int x = (int)L[K + c];
    int s = x;
    int z = (x == 0);
    if (c > 0) {
        x = (int)L[K + c - 1];
        s ^= x;
        z |= (x == 0);
    }
#pragma unroll
    for (int d = 0; d < CNT_MAX; ++d) {
        if (d < lv.cnt) {
            x = (int)L[edge_index(edges[lv.edge_begin + d], j)];
            s ^= x;
            z |= (x == 0);
        }
    }
    return (s < 0) | z;
}

template <int CNT_MAX, int MSG_WORDS>
__global__ void __launch_bounds__(kLdpcThreads, 1) ldpc_decode_kernel(const LdpcLaunch p)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t* msg = reinterpret_cast<uint32_t*>(smem + p.smem_msg_off);
    const uint4* layers = reinterpret_cast<const uint4*>(smem + p.smem_tab_off);
    const uint32_t* edges = reinterpret_cast<const uint32_t*>(smem + p.smem_tab_off + (size_t)p.q * 16);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + p.smem_bar_off);

    const int tid = threadIdx.x;
    const int N = p.N, K = p.K, q = p.q;
    const StepRecDev* __restrict__ steps = p.steps;
    const uint16_t* __restrict__ order = p.order;

    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t phase = 0;
    // stage the code's circulant table once per CTA (TMA)
    if (tid == 0) {
        mbar_expect_tx(bar, p.tab_bytes);
        tma_load_1d(smem + p.smem_tab_off, p.tab, p.tab_bytes, bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;

    for (int f = blockIdx.x; f < p.frames; f += gridDim.x) {
        // ---- soft input: HBM -> shared memory ----------------------------------------------
        const int8_t* src = p.llr + (size_t)f * N;
        // bulk copies need 16-byte aligned addresses and sizes on both sides: place the frame in
        // shared memory with the same misalignment as its global address (16 spare bytes are
        // reserved) and peel the unaligned head / tail with plain loads.
        const uint32_t mis = (uint32_t)((uintptr_t)src & 15u);
        int8_t* const L = reinterpret_cast<int8_t*>(smem) + mis;
        {
            const uint32_t head = mis ? (16u - mis) : 0u;
            const uint32_t body = ((uint32_t)N - head) & ~15u;
            const uint32_t tail = (uint32_t)N - head - body;
            if (tid == 0) {
                mbar_expect_tx(bar, body);
                tma_load_1d(L + head, src + head, body, bar);
            }
            if (tid < (int)head)
                L[tid] = src[tid];
            if (tid >= 32 && tid < 32 + (int)tail)
                L[head + body + (tid - 32)] = src[head + body + (tid - 32)];
        }
        for (int i = tid; i < p.R * MSG_WORDS; i += kLdpcThreads) // reset(): layered_decoder.hh:27-31
            msg[i] = 0;
        mbar_wait(bar, phase);
        phase ^= 1;
        __syncthreads();

        // ---- while (bad() && --trials >= 0) update();  layered_decoder.hh:153 ----------------
        int trials = p.max_trials;
        int iter = 0;
        for (;;) {
            int flag = 0;
            if (tid < kM) {
                for (int i = 0; i < q; ++i) {
                    const LayerView lv = load_layer(layers, i);
                    flag |= check_cn<CNT_MAX>(L, edges, lv, i, tid, K, q);
                }
            }
            int bad = __syncthreads_or(flag);
            if (p.group > 1) {
                // reference batch semantics: the whole SIMD batch keeps iterating while any of
                // its frames is bad.  One word per (group, iteration): low half counts
                // arrivals, high half counts bad frames.  All CTAs of a group are co-resident
                // (cooperative launch, grid a multiple of the group size).
                __shared__ int s_group_bad;
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
            ++iter;
            for (int i = 0; i < q; ++i) {
                const LayerView lv = load_layer(layers, i);
                if (lv.n_steps == 1) {
                    if (tid < kM)
                        process_cn<CNT_MAX, MSG_WORDS>(L, msg, edges, lv, i, tid, K, q);
                    __syncthreads();
                } else {
                    for (int s = 0; s < lv.n_steps; ++s) {
                        const StepRecDev st = steps[lv.step_begin + s];
                        if (tid < (int)st.count) {
                            const int j = (int)order[lv.order_begin + st.begin + tid];
                            process_cn<CNT_MAX, MSG_WORDS>(L, msg, edges, lv, i, j, K, q);
                        }
                        __syncthreads();
                    }
                }
            }
        }

        // ---- outputs ---------------------------------------------------------------------------
        // posteriors were written through the generic proxy; the TMA store below and the next
        // frame's TMA load go through the async proxy
        fence_proxy_async();
        __syncthreads();
        if (p.trials_left && tid == 0)
            p.trials_left[f] = trials;
        if (p.hard) {
            // llr < 0 -> 1, MSB first: lib/ldpc_decoder_bb_impl.cc:432-442
            uint8_t* dst = p.hard + (size_t)f * p.out_bytes;
            if ((mis & 7u) == 0) {
                const uint2* L8 = reinterpret_cast<const uint2*>(L);
                for (int b = tid; b < p.out_bytes; b += kLdpcThreads) {
                    const uint2 w = L8[b];
                    const uint32_t hi4 = ((((w.x >> 7) & 0x01010101u) * 0x08040201u) >> 24) & 0xFu;
                    const uint32_t lo4 = ((((w.y >> 7) & 0x01010101u) * 0x08040201u) >> 24) & 0xFu;
                    dst[b] = (uint8_t)((hi4 << 4) | lo4);
                }
            } else {
                for (int b = tid; b < p.out_bytes; b += kLdpcThreads) {
                    uint32_t v = 0;
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        v |= (L[8 * b + k] < 0 ? 1u : 0u) << (7 - k);
                    dst[b] = (uint8_t)v;
                }
            }
        }
        if (p.llr_post) {
            int8_t* dst = p.llr_post + (size_t)f * N;
            if ((((uintptr_t)dst) & 15u) == 0 && (N & 15) == 0 && mis == 0) {
                if (tid == 0) {
                    tma_store_1d(dst, L, (uint32_t)N);
                    tma_store_commit();
                    tma_store_wait_read();
                }
            } else {
                for (int i = tid; i < N; i += kLdpcThreads)
                    dst[i] = L[i];
            }
        }
        __syncthreads(); // L is reused by the next frame
    }
}

template <int CNT_MAX, int MSG_WORDS>
cudaError_t launch_one(const LdpcLaunch& p, int grid, size_t smem, cudaStream_t stream)
{
    auto kern = ldpc_decode_kernel<CNT_MAX, MSG_WORDS>;
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

} // namespace

size_t ldpc_smem_bytes(int N, int R, int msg_words, uint32_t tab_bytes, LdpcLaunch* p)
{
    size_t off = (((size_t)N + 15) & ~(size_t)15) + 16; // +16: frames keep their global misalignment
    if (p)
        p->smem_msg_off = (uint32_t)off;
    off += (size_t)R * msg_words * 4;
    off = (off + 15) & ~(size_t)15;
    if (p)
        p->smem_tab_off = (uint32_t)off;
    off += tab_bytes;
    off = (off + 15) & ~(size_t)15;
    if (p)
        p->smem_bar_off = (uint32_t)off;
    off += 16;
    return off;
}

cudaError_t ldpc_launch(const LdpcLaunch& p, int max_cnt, int msg_words, int grid, size_t smem, cudaStream_t stream)
{
#define DVBS2_CASE(C, W)                   \
    if (max_cnt <= C && msg_words == W)    \
        return launch_one<C, W>(p, grid, smem, stream);
    DVBS2_CASE(4, 1)
    DVBS2_CASE(6, 1)
    DVBS2_CASE(8, 1)
    DVBS2_CASE(10, 1)
    DVBS2_CASE(13, 1)
    DVBS2_CASE(16, 2)
    DVBS2_CASE(20, 2)
    DVBS2_CASE(28, 2)
#undef DVBS2_CASE
    return cudaErrorInvalidValue;
}

} // namespace dvbs2b200
