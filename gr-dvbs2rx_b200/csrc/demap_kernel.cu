// demap_kernel.cu -- XFECFRAME soft demapper (QPSK, 8PSK + column deinterleave) on sm_100a.
//
// Arithmetic contract:
//   QPSK  lib/qpsk.h:208-214: llr = convert_8i(x * (float)(2*sqrt(2)/N0)) per real component
//         (VOLK volk_32f_s32f_convert_8i: saturate to [-128,127], round to nearest even)
//   8PSK  lib/psk.hh:143-150 (rotate by e^{-j pi/8}, three quantised metrics, precision 4/N0)
//         lib/xfecframe_demapper_cb_impl.cc:48-69,162-176 (3-column deinterleave by rate)
// All float operations use explicit round-to-nearest intrinsics so nvcc cannot contract
// them into FMAs: results equal the reference built with -ffp-contract=off.
//
// HBM-bound streaming kernels: 8 bytes in, 1 byte out per LLR pair element; 128-bit loads,
// 32-bit packed stores, one grid row per frame so N0 is uniform per CTA.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace dvbs2b200 {

namespace {

// volk_32f_s32f_convert_8i_generic: clamp to [-128, 127] first, then rintf (round half to even) ...
// ... which is one instruction: round to nearest even, then saturate to int8 (identical for every input: anything
// above 127 or below -128 ends at the rail either way, NaN gives 0 both ways); the low byte of the result is the int8
__device__ __forceinline__ uint32_t convert_8i_sat(float r)
{
    int v;
    asm("cvt.rni.sat.s8.f32 %0, %1;" : "=r"(v) : "f"(r));
    return (uint32_t)v & 0xffu;
}

constexpr int kQpskLoads = 4;

__global__ void __launch_bounds__(256) demap_qpsk_kernel(const DemapLaunch p)
{
    const int frame = blockIdx.y;
    const int n_f4 = p.n_syms / 2; // float4 = 2 symbols = 4 LLRs (n_syms is even for every frame size)
    const float scalar = (float)(2.0 * 1.41421356237309504880 / (double)p.n0[frame]);
    const float4* __restrict__ in = reinterpret_cast<const float4*>(p.iq + (size_t)frame * p.n_syms * 2);
    uint32_t* __restrict__ out = reinterpret_cast<uint32_t*>(p.llr + (size_t)frame * p.n_syms * 2);
    // kQpskLoads 128-bit loads in flight per thread, streamed once
    for (int i0 = blockIdx.x * (kQpskLoads * blockDim.x) + threadIdx.x; i0 < n_f4; i0 += gridDim.x * kQpskLoads * blockDim.x) {
        float4 v[kQpskLoads];
#pragma unroll
        for (int h = 0; h < kQpskLoads; ++h)
            v[h] = i0 + h * (int)blockDim.x < n_f4 ? __ldcs(in + i0 + h * (int)blockDim.x) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int h = 0; h < kQpskLoads; ++h) {
            const int i = i0 + h * (int)blockDim.x;
            if (i < n_f4) {
                const uint32_t b0 = convert_8i_sat(__fmul_rn(v[h].x, scalar));
                const uint32_t b1 = convert_8i_sat(__fmul_rn(v[h].y, scalar));
                const uint32_t b2 = convert_8i_sat(__fmul_rn(v[h].z, scalar));
                const uint32_t b3 = convert_8i_sat(__fmul_rn(v[h].w, scalar));
                __stcs(out + i, b0 | (b1 << 8) | (b2 << 16) | (b3 << 24));
            }
        }
    }
}

// lib/psk.hh:123-131 quantize: value *= DIST*precision; nearbyint; clamp; cast
__device__ __forceinline__ uint32_t quantize8(float scale, float value)
{
    float r = rintf(__fmul_rn(value, scale));
    r = fminf(fmaxf(r, -128.0f), 127.0f);
    return (uint32_t)((int)r) & 0xffu;
}

constexpr int kPskGroups = 2;

__global__ void __launch_bounds__(256) demap_8psk_kernel(const DemapLaunch p)
{
    const int frame = blockIdx.y;
    const float rcp_sqrt_2 = 0.70710678118654752440f;
    const float dist = 2 * 0.38268343236508977173f;           // 2 sin(pi/8)
    const float rot_re = (float)0.92387953251128675613;       // cos(-pi/8)
    const float rot_im = (float)-0.38268343236508977173;      // sin(-pi/8)
    const float precision = (float)(4.0 / (double)p.n0[frame]);
    const float scale = __fmul_rn(dist, precision);
    const int n4 = p.n_syms / 4; // 4 symbols per thread -> one packed 32-bit store per column
    const float4* __restrict__ in = reinterpret_cast<const float4*>(p.iq + (size_t)frame * p.n_syms * 2);
    int8_t* __restrict__ out = p.llr + (size_t)frame * p.n_syms * 3;
    uint32_t* __restrict__ c0 = reinterpret_cast<uint32_t*>(out + p.row0);
    uint32_t* __restrict__ c1 = reinterpret_cast<uint32_t*>(out + p.row1);
    uint32_t* __restrict__ c2 = reinterpret_cast<uint32_t*>(out + p.row2);
    // kPskGroups groups of 4 symbols per thread: all their 128-bit loads are issued before the arithmetic starts
    for (int i0 = blockIdx.x * (kPskGroups * blockDim.x) + threadIdx.x; i0 < n4; i0 += gridDim.x * kPskGroups * blockDim.x) {
        float4 v[kPskGroups][2];
#pragma unroll
        for (int g = 0; g < kPskGroups; ++g) {
            const int i = i0 + g * (int)blockDim.x;
#pragma unroll
            for (int h = 0; h < 2; ++h)
                v[g][h] = i < n4 ? __ldcs(in + 2 * i + h) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int g = 0; g < kPskGroups; ++g) {
            const int i = i0 + g * (int)blockDim.x;
            if (i >= n4)
                break;
            uint32_t w0 = 0, w1 = 0, w2 = 0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float a[2] = { v[g][h].x, v[g][h].z }, b[2] = { v[g][h].y, v[g][h].w };
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    // std::complex<float> c *= rot: (a*c - b*d) + (a*d + b*c) i, unfused
                    const float re = __fsub_rn(__fmul_rn(a[s], rot_re), __fmul_rn(b[s], rot_im));
                    const float im = __fadd_rn(__fmul_rn(a[s], rot_im), __fmul_rn(b[s], rot_re));
                    const float m0 = __fmul_rn(rcp_sqrt_2, __fsub_rn(fabsf(re), fabsf(im)));
                    const int sh = 8 * (2 * h + s);
                    w0 |= quantize8(scale, m0) << sh;
                    w1 |= quantize8(scale, re) << sh;
                    w2 |= quantize8(scale, im) << sh;
                }
            }
            __stcs(c0 + i, w0);
            __stcs(c1 + i, w1);
            __stcs(c2 + i, w2);
        }
    }
}

// ---- SNR estimate of one XFECFRAME per CTA (lib/xfecframe_demapper_cb_impl.cc:128-142,267-302; lib/qpsk.h:41-65,
//      240-281): Es/N0 = sum |ref|^2 / sum |x - ref|^2 with the reference points taken from the sliced symbols
//      (llr == null) or from the signs of the posterior LLRs.  Floating point sums: tolerance-checked, the
//      reference's own summation order (VOLK) is unspecified.  HBM bound: 8 bytes per symbol (+ bits per symbol).
// One symbol's contribution.  CONSTELLATION and HAS_LLR are compile-time: no branch per symbol.
template <int CONSTELLATION, bool HAS_LLR>
__device__ __forceinline__ void snr_symbol(const SnrLaunch& p, const int8_t* __restrict__ llr, int j, float xr, float xi, float& sp, float& np)
{
    const float a = 0.70710678118654752440f;
    const float rot_re = (float)0.92387953251128675613, rot_im = (float)-0.38268343236508977173;
    float sr, si;
    if (CONSTELLATION == 0) {
        const bool pr = HAS_LLR ? llr[2 * j] >= 0 : xr >= 0.f, pi = HAS_LLR ? llr[2 * j + 1] >= 0 : xi >= 0.f;
        sr = pr ? a : -a;
        si = pi ? a : -a;
    } else {
        bool b0, b1, b2; // true = the bit's LLR is negative
        if (HAS_LLR) {
            b0 = llr[p.row0 + j] < 0, b1 = llr[p.row1 + j] < 0, b2 = llr[p.row2 + j] < 0;
        } else { // lib/psk.hh:135-141
            const float re = __fsub_rn(__fmul_rn(xr, rot_re), __fmul_rn(xi, rot_im));
            const float im = __fadd_rn(__fmul_rn(xr, rot_im), __fmul_rn(xi, rot_re));
            b1 = re < 0.f, b2 = im < 0.f, b0 = fabsf(re) < fabsf(im);
        }
        // lib/psk.hh:114-121,152-157: point 4 b0 + 2 b1 + b2 of { (a,a), (1,0), (-1,0), (-a,-a), (0,1), (a,-a), (-a,a), (0,-1) },
        // by selects (an indexed table would live in local memory)
        sr = b0 ? (b1 ? (b2 ? 0.f : -a) : (b2 ? a : 0.f)) : (b1 ? (b2 ? -a : -1.f) : (b2 ? 1.f : a));
        si = b0 ? (b1 ? (b2 ? -1.f : a) : (b2 ? -a : 1.f)) : (b1 ? (b2 ? -a : 0.f) : (b2 ? 0.f : a));
    }
    const float er = xr - sr, ei = xi - si;
    sp += sr * sr + si * si;
    np += er * er + ei * ei;
}

// Two symbols per 128-bit load, kSnrLoads loads in flight per thread (a frame is 64.8 - 259 KB of symbols: with one
// 8-byte load per thread and iteration the CTA had too little under way to cover the DRAM latency).
constexpr int kSnrLoads = 4;

template <int CONSTELLATION, bool HAS_LLR>
__device__ __forceinline__ void snr_frame(const SnrLaunch& p, int frame, int tid, float& sp, float& np)
{
    const float4* __restrict__ in = reinterpret_cast<const float4*>(p.iq + (size_t)frame * p.n_syms * 2);
    const int8_t* __restrict__ llr = HAS_LLR ? p.llr + (size_t)frame * p.n_syms * (CONSTELLATION == 0 ? 2 : 3) : nullptr;
    const int pairs = p.n_syms >> 1; // n_syms is even for every frame size (checked by the host)
    for (int j0 = tid; j0 < pairs; j0 += 256 * kSnrLoads) {
        float4 x[kSnrLoads];
#pragma unroll
        for (int h = 0; h < kSnrLoads; ++h)
            x[h] = j0 + 256 * h < pairs ? __ldcs(in + j0 + 256 * h) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int h = 0; h < kSnrLoads; ++h) {
            const int j = 2 * (j0 + 256 * h);
            if (j0 + 256 * h < pairs) {
                snr_symbol<CONSTELLATION, HAS_LLR>(p, llr, j, x[h].x, x[h].y, sp, np);
                snr_symbol<CONSTELLATION, HAS_LLR>(p, llr, j + 1, x[h].z, x[h].w, sp, np);
            }
        }
    }
}

__global__ void __launch_bounds__(256) snr_kernel(const SnrLaunch p)
{
    const int frame = blockIdx.x, tid = threadIdx.x;
    float sp = 0.f, np = 0.f;
    if (p.constellation == 0) {
        if (p.llr)
            snr_frame<0, true>(p, frame, tid, sp, np);
        else
            snr_frame<0, false>(p, frame, tid, sp, np);
    } else {
        if (p.llr)
            snr_frame<4, true>(p, frame, tid, sp, np);
        else
            snr_frame<4, false>(p, frame, tid, sp, np);
    }
    __shared__ float s_sp[8], s_np[8];
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        sp += __shfl_xor_sync(0xffffffffu, sp, m);
        np += __shfl_xor_sync(0xffffffffu, np, m);
    }
    if ((tid & 31) == 0)
        s_sp[tid >> 5] = sp, s_np[tid >> 5] = np;
    __syncthreads();
    if (tid == 0) {
        float tsp = 0.f, tnp = 0.f;
        for (int w = 0; w < 8; ++w)
            tsp += s_sp[w], tnp += s_np[w];
        if (!(tnp > 0.f))
            tnp = 1e-12f;
        p.snr_lin[frame] = tsp / tnp;
    }
}

// ---- table-driven max-log soft demapper (16APSK, 32APSK, any constellation of up to 32 points) -------------
// The reference has no demapper beyond QPSK / 8PSK (lib/xfecframe_demapper_cb_impl.cc:70-72 throws), so there
// is no parity target; this is SURVEY 8f rank 3, defined here as the max-log LLR
//     llr_k = ( min_{s: bit k of s = 1} |y - s|^2  -  min_{s: bit k of s = 0} |y - s|^2 ) / N0
// (positive = bit 0, the convention of lib/qpsk.h:208-214; for QPSK it reduces to the reference's 2 sqrt(2) x / N0),
// rounded to nearest even and saturated to int8 like the other two.  Symbol index s = the symbol's bits, first
// bit = MSB.  Bit k of symbol j goes to llr[row_off[k] + j]: the DVB-S2 block interleaver read back, any column
// order.
// |y|^2 is common to every distance and cancels in the difference: the metric of point s is |s|^2 - 2 Re(y conj s),
// two FMAs.  The 2 x BITS minima come from a halving tree: at the level of label bit b the metrics are split by
// that bit (two minima) and folded pairwise over it for the remaining bits -- 82 min operations instead of 160
// for 32 points, and the compiler fuses pairs of them into three-input FMNMX3.  A thread demaps two consecutive
// symbols (one 128-bit load) and stores one 16-bit word per bit row.  About 170 instructions per 32APSK symbol, 105 of
// them on the ALU pipe (FMNMX): that pipe, not HBM, is the bound.
#ifndef DVBS2_TABLE_LOADS
#define DVBS2_TABLE_LOADS 4
#endif
constexpr int kTableLoads = DVBS2_TABLE_LOADS;

template <int BITS>
__global__ void __launch_bounds__(256, 3) demap_table_kernel(const TableDemapLaunch p, const TableDemapConst t)
{
    // the constellation rides in the kernel parameters (constant bank): -2 Re s, -2 Im s, |s|^2 are operands of the
    // FMAs straight from there -- no shared-memory loads, no registers held for them
    constexpr int NP = 1 << BITS;
    const int frame = blockIdx.y;
    const float inv_n0 = 1.0f / p.n0[frame];
    const float4* __restrict__ in = reinterpret_cast<const float4*>(p.iq + (size_t)frame * p.n_syms * 2);
    int8_t* __restrict__ out = p.llr + (size_t)frame * p.n_syms * BITS;
    // two consecutive symbols per thread (one 128-bit load), one 16-bit store per bit row: n_syms and the row offsets are
    // even for every DVB-S2 APSK frame (4050, 16200, 3240, 12960 symbols) -- checked by the host
    const int pairs = p.n_syms >> 1;
    // kTableLoads 128-bit loads in flight per thread before the arithmetic starts: with one, a warp has 512 bytes under
    // way for ~400 instructions of work and the SM cannot cover the DRAM latency (long-scoreboard stalls, 23 - 31 % of
    // the HBM rate; four loads: 26 - 37 %, then the ALU pipe binds).  The symbol pairs are taken one after the other
    // (unrolling them spills the 32 metrics of a symbol)
    for (int pr0 = blockIdx.x * (kTableLoads * blockDim.x) + threadIdx.x; pr0 < pairs; pr0 += gridDim.x * kTableLoads * blockDim.x) {
        float4 y[kTableLoads];
#pragma unroll
        for (int h = 0; h < kTableLoads; ++h) {
            const int pr = pr0 + h * (int)blockDim.x;
            y[h] = pr < pairs ? __ldcs(in + pr) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll 1
        for (int h = 0; h < kTableLoads; ++h) {
            const int pr = pr0 + h * (int)blockDim.x;
            if (pr >= pairs)
                break;
            float4 y01 = y[0];
#pragma unroll
            for (int g = 1; g < kTableLoads; ++g)
                y01 = h == g ? y[g] : y01;
            uint32_t packed[BITS];
#pragma unroll
            for (int k = 0; k < BITS; ++k)
                packed[k] = 0u;
#pragma unroll 1 // one symbol at a time: 2^BITS live metrics
            for (int u = 0; u < 2; ++u) {
                const float yx = u == 0 ? y01.x : y01.z, yy = u == 0 ? y01.y : y01.w;
                float v[NP];
#pragma unroll
                for (int s = 0; s < NP; ++s)
                    v[s] = fmaf(yx, t.a[s], fmaf(yy, t.b[s], t.c[s]));
                // level lvl: label bit k = BITS - 1 - lvl is the LOW bit of the current index
#pragma unroll
                for (int lvl = 0; lvl < BITS; ++lvl) {
                    const int n = NP >> lvl, k = BITS - 1 - lvl;
                    float d0 = v[0], d1 = v[1];
#pragma unroll
                    for (int i2 = 2; i2 < n; i2 += 2) {
                        d0 = fminf(d0, v[i2]);
                        d1 = fminf(d1, v[i2 + 1]);
                    }
#pragma unroll
                    for (int i2 = 0; i2 < n; i2 += 2)
                        v[i2 >> 1] = fminf(v[i2], v[i2 + 1]);
                    packed[k] |= convert_8i_sat(__fmul_rn(__fsub_rn(d1, d0), inv_n0)) << (8 * u);
                }
            }
#pragma unroll
            for (int k = 0; k < BITS; ++k)
                *reinterpret_cast<uint16_t*>(out + t.row[k] + 2 * pr) = (uint16_t)packed[k];
        }
    }
}

} // namespace

cudaError_t demap_table_launch(const TableDemapLaunch& p, const TableDemapConst& t, cudaStream_t stream)
{
    if (p.frames <= 0)
        return cudaSuccess;
    if (p.n_syms % 2)
        return cudaErrorInvalidValue;
    dim3 grid((p.n_syms / 2 + 256 * kTableLoads - 1) / (256 * kTableLoads), p.frames);
    switch (p.bits) {
    case 1: demap_table_kernel<1><<<grid, 256, 0, stream>>>(p, t); break;
    case 2: demap_table_kernel<2><<<grid, 256, 0, stream>>>(p, t); break;
    case 3: demap_table_kernel<3><<<grid, 256, 0, stream>>>(p, t); break;
    case 4: demap_table_kernel<4><<<grid, 256, 0, stream>>>(p, t); break;
    case 5: demap_table_kernel<5><<<grid, 256, 0, stream>>>(p, t); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t snr_launch(const SnrLaunch& p, cudaStream_t stream)
{
    if (p.frames <= 0)
        return cudaSuccess;
    if (p.n_syms % 2)
        return cudaErrorInvalidValue;
    snr_kernel<<<p.frames, 256, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t demap_launch(const DemapLaunch& p, cudaStream_t stream)
{
    if (p.frames <= 0)
        return cudaSuccess;
    const int threads = 256;
    if (p.constellation == 0) {
        const int work = p.n_syms / 2;
        dim3 grid((work + threads * kQpskLoads - 1) / (threads * kQpskLoads), p.frames);
        demap_qpsk_kernel<<<grid, threads, 0, stream>>>(p);
    } else if (p.constellation == 4) {
        const int work = p.n_syms / 4;
        dim3 grid((work + threads * kPskGroups - 1) / (threads * kPskGroups), p.frames);
        demap_8psk_kernel<<<grid, threads, 0, stream>>>(p);
    } else {
        return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// Forces the module that holds these kernels to be loaded now (CUDA loads lazily at the first launch, and that
// load can wait for the device to go idle -- which never happens while the persistent LDPC kernel of the
// streaming path is resident and waiting for input that the blocked host thread has yet to send).
cudaError_t demap_preload()
{
    cudaFuncAttributes a;
    cudaError_t e;
    if ((e = cudaFuncGetAttributes(&a, demap_table_kernel<4>)) != cudaSuccess || (e = cudaFuncGetAttributes(&a, demap_table_kernel<5>)) != cudaSuccess)
        return e;
    if ((e = cudaFuncGetAttributes(&a, snr_kernel)) != cudaSuccess)
        return e;
    if ((e = cudaFuncGetAttributes(&a, demap_qpsk_kernel)) != cudaSuccess)
        return e;
    if ((e = cudaFuncGetAttributes(&a, demap_8psk_kernel)) != cudaSuccess)
        return e;
    return cudaSuccess;
}

} // namespace dvbs2b200
