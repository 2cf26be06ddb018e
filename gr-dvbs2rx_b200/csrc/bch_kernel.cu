// bch_kernel.cu -- binary BCH decoder over GF(2^m) for the DVB-S2 outer code on sm_100a.
//
// Behavioural contract (bit-exact outputs and return codes):
//   lib/bch.cc:467-487  decode(u8): copy systematic bytes, syndrome, Berlekamp, roots, flips
//   lib/bch.cc:216-222,175-189  syndromes S_i = r(alpha^i), i = 1..2t (all-zero -> return 0)
//   lib/bch.cc:224-304  simplified Berlekamp, Lin & Costello table form (rows mu = -1/2, 0..t)
//   lib/bch.cc:306-385 + lib/gf.cc:289-404  roots over exponents s+1 .. n+s
//   lib/bch.cc:428-452  flips of message bits only, network bit order
//
// Design: persistent CTAs (one per SM, 32 warps), one warp per codeword at a time.  The antilog table of the
// field (2^m uint16: 128 KB for GF(2^16)) is copied into shared memory once per CTA -- every hot loop of the
// decoder is a stream of random look-ups into it.  Syndromes are evaluated directly as r(alpha^i) (the reference
// first reduces r mod g with a serial byte LUT; same field element), odd i only (S_2i = S_i^2 in a binary code),
// a BYTE at a time: B_i[b] = b(alpha^i) is tabulated in the log domain for the 256 byte values (built in shared
// memory at kernel start), so a byte costs one table look-up + one antilog look-up per syndrome instead of one per
// set bit.  Berlekamp's table keeps one polynomial coefficient per lane (degree <= 2t-1 < 32) so the discrepancy is
// a warp XOR reduction and the polynomial update a shuffle.  The Chien search strides the n exponents over the
// lanes in the log domain.
//
// Where the reference would throw out of general_work (closed-form degree-1/2 roots landing
// outside the shortened code, lib/bch.cc:441 / lib/gf.h:110 -- needs > t errors imitating a
// 1- or 2-error syndrome), this kernel reports -1 and flips nothing.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "kernels.h"

namespace dvbs2b200 {

namespace {

constexpr int kMaxT = 12;
constexpr unsigned kFull = 0xffffffffu;

struct Gf {
    const uint16_t* antilog;            // shared memory
    const uint16_t* __restrict__ log;   // global, read-only path (Berlekamp only)
    uint32_t nz; // 2^m - 1
    int m;
    __device__ __forceinline__ uint32_t fold(uint32_t x) const
    { // x mod (2^m - 1) for x < 2^(2m)
        x = (x & nz) + (x >> m);
        x = (x & nz) + (x >> m);
        return x >= nz ? x - nz : x;
    }
    __device__ __forceinline__ uint32_t alpha(uint32_t e) const { return antilog[e]; } // e <= nz
    __device__ __forceinline__ uint32_t mul(uint32_t a, uint32_t b) const
    {
        if (!a || !b)
            return 0;
        uint32_t e = (uint32_t)__ldg(log + a) + (uint32_t)__ldg(log + b);
        return alpha(e >= nz ? e - nz : e);
    }
    __device__ __forceinline__ uint32_t div(uint32_t a, uint32_t b) const
    { // b != 0
        if (!a)
            return 0;
        uint32_t e = (uint32_t)__ldg(log + a) + nz - (uint32_t)__ldg(log + b);
        return alpha(e >= nz ? e - nz : e);
    }
};

__device__ __forceinline__ uint32_t warp_xor(uint32_t v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v ^= __shfl_xor_sync(kFull, v, o);
    return v;
}

constexpr int kBchWarps = 32; // warps per (persistent) CTA

// shared-memory carve-up (dynamic): antilog [2^m] | logB [kMaxT][256] | sig [warps][kMaxT + 2][32] | S [warps][2 kMaxT]
// | roots [warps][kMaxT + 1] (uint32) | nroots [warps] (int)
__host__ __device__ inline size_t bch_smem_bytes(int m)
{
    return ((size_t)2 << m) + (size_t)kMaxT * 256 * 2 + (size_t)kBchWarps * (kMaxT + 2) * 32 * 2 + (size_t)kBchWarps * 2 * kMaxT * 2 +
           (size_t)kBchWarps * (kMaxT + 1) * 4 + (size_t)kBchWarps * 4;
}

// T = error-correction capability (8, 10 or 12 for DVB-S2): compile time, so that the per-syndrome code is straight line
template <int T>
__global__ void __launch_bounds__(kBchWarps * 32, 1) bch_decode_kernel(const BchLaunch p)
{
    extern __shared__ __align__(16) uint8_t bch_smem[];
    uint16_t* const s_antilog = reinterpret_cast<uint16_t*>(bch_smem);
    uint16_t* const s_logB = s_antilog + ((size_t)1 << p.m);
    uint16_t(*const s_sig)[kMaxT + 2][32] = reinterpret_cast<uint16_t(*)[kMaxT + 2][32]>(s_logB + kMaxT * 256);
    uint16_t(*const s_S)[2 * kMaxT] = reinterpret_cast<uint16_t(*)[2 * kMaxT]>(&s_sig[kBchWarps][0][0]);
    uint32_t(*const s_roots)[kMaxT + 1] = reinterpret_cast<uint32_t(*)[kMaxT + 1]>(&s_S[kBchWarps][0]);
    int* const s_nroots = reinterpret_cast<int*>(&s_roots[kBchWarps][0]);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Gf gf;
    gf.antilog = s_antilog;
    gf.log = p.log;
    gf.m = p.m;
    gf.nz = (1u << p.m) - 1u;
    const int n = p.n, k = p.k;
    constexpr int t = T;
    const int n_bytes = n >> 3, k_bytes = k >> 3;

    // ---- tables: antilog into shared memory, then B_a[b] = sum_kbit bit_kbit(b) alpha^(a (7 - kbit)) in the log domain
    {
        const uint4* src = reinterpret_cast<const uint4*>(p.antilog);
        uint4* dst = reinterpret_cast<uint4*>(s_antilog);
        for (int i = threadIdx.x; i < (int)(((size_t)2 << p.m) / 16); i += blockDim.x)
            dst[i] = __ldg(src + i);
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < t * 256; idx += blockDim.x) {
        const uint32_t a = (uint32_t)(2 * (idx >> 8) + 1), b = (uint32_t)(idx & 255);
        uint32_t v = 0;
#pragma unroll
        for (int kb = 0; kb < 8; ++kb)
            if (b & (0x80u >> kb))
                v ^= gf.alpha(gf.fold(a * (uint32_t)(7 - kb)));
        s_logB[idx] = v ? __ldg(p.log + v) : (uint16_t)0; // b = 0: masked by the caller
    }
    __syncthreads();

    // frames go round the CTAs first: every SM gets its share even when there are fewer frames than warps
    for (int frame = warp * gridDim.x + blockIdx.x; frame < p.frames; frame += gridDim.x * kBchWarps) {
        const uint8_t* __restrict__ cw = p.cw + (size_t)frame * p.cw_stride;
        uint8_t* __restrict__ msg = p.msg + (size_t)frame * p.msg_stride;

        // ---- systematic copy (lib/bch.cc:471) + odd syndromes -----------------------------------
        // byte y holds the coefficients of x^(n - 1 - 8y) .. x^(n - 8 - 8y): term = B_a[byte] * alpha^(a (n - 8 - 8y))
        uint32_t S_odd[kMaxT], ex[kMaxT], st[kMaxT];
#pragma unroll
        for (int i = 0; i < kMaxT; ++i) {
            S_odd[i] = 0;
            const uint32_t a = (uint32_t)(2 * i + 1);
            ex[i] = gf.fold(a * (uint32_t)(n - 8 - 8 * lane)); // lane < n_bytes always (n_bytes >= 32)
            st[i] = gf.fold(a * 256u);                         // 32 bytes further on
        }
        const uint32_t nz = gf.nz;
        for (int y = lane; y < n_bytes; y += 32) {
            const uint32_t byte = __ldg(cw + y);
            if (y < k_bytes)
                msg[y] = (uint8_t)byte;
            const uint32_t live = byte ? 0xffffu : 0u; // B_a[0] = 0 has no logarithm: its table entry is 0, the term is masked
#pragma unroll
            for (int i = 0; i < kMaxT; ++i) {
                if (i < t) {
                    const uint32_t e = (uint32_t)s_logB[i * 256 + byte] + ex[i];
                    S_odd[i] ^= (uint32_t)s_antilog[min(e, e - nz)] & live; // unsigned: e - nz wraps when e < nz
                    const uint32_t d = ex[i] - st[i];
                    ex[i] = min(d, d + nz); // (ex - st) mod nz
                }
            }
        }
        uint32_t any = 0;
#pragma unroll
        for (int i = 0; i < kMaxT; ++i) {
            S_odd[i] = warp_xor(S_odd[i]);
            any |= S_odd[i];
        }
        __syncwarp();
        if (!any) { // lib/bch.cc:179-180: zero remainder <=> all syndromes zero -> no errors
            if (lane == 0 && p.corrections)
                p.corrections[frame] = 0;
            continue;
        }
        // S[0..2t) = S_1..S_2t; S_2j = S_j^2
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < kMaxT; ++i)
                if (i < t)
                    s_S[warp][2 * i] = (uint16_t)S_odd[i];
            for (int j = 1; j <= t; ++j) { // S_{2j} from S_j (ascending j: S_j already known)
                const uint32_t sj = s_S[warp][j - 1];
                s_S[warp][2 * j - 1] = (uint16_t)gf.mul(sj, sj);
            }
        }
        __syncwarp();
        const uint16_t* S = s_S[warp];

        // ---- Berlekamp, table form (lib/bch.cc:224-304); lane = coefficient index -----------------
        int deg[kMaxT + 3];
        int two_mu[kMaxT + 3];
        uint32_t dis[kMaxT + 3];
        two_mu[0] = -1;
#pragma unroll
        for (int i = 0; i < kMaxT + 2; ++i)
            two_mu[i + 1] = 2 * i;
        s_sig[warp][0][lane] = (lane == 0);
        s_sig[warp][1][lane] = (lane == 0);
        s_sig[warp][2][lane] = (lane == 0) ? 1 : (lane == 1 ? S[0] : 0);
        deg[0] = 0;
        deg[1] = 0;
        deg[2] = S[0] ? 1 : 0;
        dis[0] = 1;
        dis[1] = S[0];
        __syncwarp();
        int row = 2;
        for (; row <= t; ++row) {
            const int tm = two_mu[row];
            const uint32_t cur = s_sig[warp][row][lane];
            uint32_t term = 0;
            if (lane >= 1 && lane <= deg[row] && lane <= tm && cur)
                term = gf.mul(cur, S[tm - lane]);
            const uint32_t d = (uint32_t)S[tm] ^ warp_xor(term);
            dis[row] = d;
            uint32_t nxt = cur;
            if (d != 0) {
                int row_rho = 0, max_diff = -2;
                for (int j = row - 1; j >= 0; --j) // latest row wins ties (strict >)
                    if (dis[j] != 0) {
                        const int diff = two_mu[j] - deg[j];
                        if (diff > max_diff) {
                            max_diff = diff;
                            row_rho = j;
                        }
                    }
                const uint32_t coef = gf.div(d, dis[row_rho]);
                const int shift = tm - two_mu[row_rho];
                if (lane >= shift)
                    nxt ^= gf.mul(coef, s_sig[warp][row_rho][lane - shift]);
            }
            s_sig[warp][row + 1][lane] = (uint16_t)nxt;
            const unsigned nzmask = __ballot_sync(kFull, nxt != 0);
            deg[row + 1] = nzmask ? 31 - __clz((int)nzmask) : -1;
            __syncwarp();
        }
        const int L = deg[row];
        const uint32_t my_sigma = s_sig[warp][row][lane];

        // ---- roots (lib/bch.cc:306-385) --------------------------------------------------------------
        if (lane == 0)
            s_nroots[warp] = 0;
        __syncwarp();
        int found = 0;
        if (L >= 1 && L <= t) {
            // sigma(alpha^e) = XOR_j alpha^(log sigma_j + e*j); lanes take e = s+1+lane, +32, ...
            uint32_t acc[kMaxT + 1], step[kMaxT + 1];
            bool on[kMaxT + 1];
            const uint32_t e0 = p.shorten + 1u + (uint32_t)lane;
#pragma unroll
            for (int j = 0; j <= kMaxT; ++j) {
                const uint32_t sj = __shfl_sync(kFull, my_sigma, j);
                on[j] = (j <= L) && (sj != 0);
                const uint32_t lg = on[j] ? (uint32_t)__ldg(gf.log + sj) : 0u;
                acc[j] = gf.fold(lg + gf.fold(e0 * (uint32_t)j));
                step[j] = gf.fold(32u * (uint32_t)j);
            }
            const uint32_t e_end = (uint32_t)n + p.shorten; // inclusive
            bool dense = (L == T);
#pragma unroll
            for (int j = 0; j <= T; ++j)
                dense = dense && on[j];
            if (dense) {
                // the usual case of an uncorrectable word (and of t errors): all t + 1 coefficients present -- no predicates
                uint32_t stm[T + 1];
#pragma unroll
                for (int j = 0; j <= T; ++j)
                    stm[j] = step[j] - nz;
                for (uint32_t e = e0; e <= e_end; e += 32) {
                    uint32_t res = 0;
#pragma unroll
                    for (int j = 0; j <= T; ++j) {
                        res ^= (uint32_t)s_antilog[acc[j]];
                        acc[j] = min(acc[j] + step[j], acc[j] + stm[j]); // (acc + step) mod nz, unsigned
                    }
                    if (res == 0) {
                        const int slot = atomicAdd(&s_nroots[warp], 1);
                        if (slot <= kMaxT)
                            s_roots[warp][slot] = e;
                    }
                }
            } else {
                for (uint32_t e = e0; e <= e_end; e += 32) {
                    uint32_t res = 0;
#pragma unroll
                    for (int j = 0; j <= kMaxT; ++j) {
                        if (on[j]) {
                            res ^= (uint32_t)s_antilog[acc[j]];
                            const uint32_t a = acc[j] + step[j];
                            acc[j] = min(a, a - nz);
                        }
                    }
                    if (res == 0) {
                        const int slot = atomicAdd(&s_nroots[warp], 1);
                        if (slot <= kMaxT)
                            s_roots[warp][slot] = e;
                    }
                }
            }
            __syncwarp();
            found = s_nroots[warp];
        }

        // ---- flips (lib/bch.cc:428-452) and return code (:476-483) -----------------------------------
        const bool closed_form_failure = (L <= 2) && (found != L); // reference throws / returns {} here
        __syncwarp(); // the systematic copy above (all lanes) before lane 0 flips bits in it
        if (lane == 0) {
            if (!closed_form_failure) {
                for (int r = 0; r < found && r <= kMaxT; ++r) {
                    const uint32_t bit_idx = gf.nz - s_roots[warp][r]; // locator exponent, in [0, n)
                    if (bit_idx < (uint32_t)(n - k))
                        continue; // parity bit: the message is all that is emitted
                    const uint32_t net = (uint32_t)n - 1u - bit_idx;
                    msg[net >> 3] ^= (uint8_t)(1u << (7u - (net & 7u)));
                }
            }
            if (p.corrections)
                p.corrections[frame] = (found == L && L >= 1) ? found : -1;
        }
        __syncwarp();
    }
}

} // namespace

cudaError_t bch_launch(const BchLaunch& p, cudaStream_t stream)
{
    if (p.frames <= 0)
        return cudaSuccess;
    if (p.t > kMaxT || p.m < 8 || p.m > 16 || (p.n >> 3) < 32)
        return cudaErrorInvalidValue;
    const size_t smem = bch_smem_bytes(p.m);
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = std::max(1, std::min(sms, p.frames)); // frames go round the CTAs: every SM takes part
    auto go = [&](auto kern) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess)
            return e;
        kern<<<blocks, kBchWarps * 32, smem, stream>>>(p);
        return cudaGetLastError();
    };
    switch (p.t) {
    case 8: return go(bch_decode_kernel<8>);
    case 10: return go(bch_decode_kernel<10>);
    case 12: return go(bch_decode_kernel<12>);
    default: return cudaErrorInvalidValue; // lib/fec_params.cc knows no other t
    }
}

// Forces the module that holds these kernels to be loaded now (CUDA loads lazily at the first launch, and that
// load can wait for the device to go idle -- which never happens while the persistent LDPC kernel of the
// streaming path is resident and waiting for input that the blocked host thread has yet to send).
cudaError_t bch_preload()
{
    cudaFuncAttributes a;
    cudaError_t e;
    if ((e = cudaFuncGetAttributes(&a, bch_decode_kernel<8>)) != cudaSuccess || (e = cudaFuncGetAttributes(&a, bch_decode_kernel<10>)) != cudaSuccess ||
        (e = cudaFuncGetAttributes(&a, bch_decode_kernel<12>)) != cudaSuccess)
        return e;
    return cudaSuccess;
}

} // namespace dvbs2b200
