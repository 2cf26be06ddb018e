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

// volk_32f_s32f_convert_8i_generic: clamp first, then rintf (round half to even)
__device__ __forceinline__ int convert_8i(float r)
{
    if (r > 127.0f)
        return 127;
    if (r < -128.0f)
        return -128;
    return __float2int_rn(r);
}

__global__ void __launch_bounds__(256) demap_qpsk_kernel(const DemapLaunch p)
{
    const int frame = blockIdx.y;
    const int n_f4 = p.n_syms / 2; // float4 = 2 symbols = 4 LLRs (n_syms is even for every frame size)
    const float scalar = (float)(2.0 * 1.41421356237309504880 / (double)p.n0[frame]);
    const float4* __restrict__ in = reinterpret_cast<const float4*>(p.iq + (size_t)frame * p.n_syms * 2);
    uint32_t* __restrict__ out = reinterpret_cast<uint32_t*>(p.llr + (size_t)frame * p.n_syms * 2);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_f4; i += gridDim.x * blockDim.x) {
        const float4 v = __ldcs(in + i); // streamed once
        const uint32_t b0 = (uint32_t)convert_8i(__fmul_rn(v.x, scalar)) & 0xffu;
        const uint32_t b1 = (uint32_t)convert_8i(__fmul_rn(v.y, scalar)) & 0xffu;
        const uint32_t b2 = (uint32_t)convert_8i(__fmul_rn(v.z, scalar)) & 0xffu;
        const uint32_t b3 = (uint32_t)convert_8i(__fmul_rn(v.w, scalar)) & 0xffu;
        out[i] = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
    }
}

// lib/psk.hh:123-131 quantize: value *= DIST*precision; nearbyint; clamp; cast
__device__ __forceinline__ uint32_t quantize8(float scale, float value)
{
    float r = rintf(__fmul_rn(value, scale));
    r = fminf(fmaxf(r, -128.0f), 127.0f);
    return (uint32_t)((int)r) & 0xffu;
}

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
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
        uint32_t w0 = 0, w1 = 0, w2 = 0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const float4 v = __ldcs(in + 2 * i + h);
            const float a[2] = { v.x, v.z }, b[2] = { v.y, v.w };
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
        c0[i] = w0;
        c1[i] = w1;
        c2[i] = w2;
    }
}

} // namespace

cudaError_t demap_launch(const DemapLaunch& p, cudaStream_t stream)
{
    if (p.frames <= 0)
        return cudaSuccess;
    const int threads = 256;
    if (p.constellation == 0) {
        const int work = p.n_syms / 2;
        dim3 grid((work + threads * 2 - 1) / (threads * 2), p.frames);
        demap_qpsk_kernel<<<grid, threads, 0, stream>>>(p);
    } else if (p.constellation == 4) {
        const int work = p.n_syms / 4;
        dim3 grid((work + threads - 1) / threads, p.frames);
        demap_8psk_kernel<<<grid, threads, 0, stream>>>(p);
    } else {
        return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

} // namespace dvbs2b200
