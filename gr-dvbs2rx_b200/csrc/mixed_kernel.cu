// mixed_kernel.cu -- gather / scatter of variable-size frames for mixed-MODCOD batches.
//
// A VCM/ACM stream carries FECFRAMEs of different codes back to back (BASELINE config 5).  The decoder
// kernels want the frames of ONE code contiguous, so a mixed batch is bucketed by code: the frames of a
// code are gathered into a contiguous staging area, decoded by that code's kernels on that code's
// stream, and the results are scattered back to the frames' positions in the batch.  Pure byte moves,
// HBM bound: n_ldpc bytes in + kbch/8 bytes out per frame on top of the decoder's own traffic.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace dvbs2b200 {

namespace {

// dst[k][0..bytes) = src[off[k] ..], 16-byte vectors when everything is aligned
__global__ void gather_kernel(const uint8_t* __restrict__ src, const unsigned long long* __restrict__ off, uint8_t* __restrict__ dst,
                              int bytes)
{
    const uint8_t* s = src + off[blockIdx.x];
    uint8_t* d = dst + (size_t)blockIdx.x * bytes;
    if ((((uintptr_t)s | (uintptr_t)d | (uintptr_t)bytes) & 15) == 0) {
        const uint4* s4 = reinterpret_cast<const uint4*>(s);
        uint4* d4 = reinterpret_cast<uint4*>(d);
        for (int i = threadIdx.x; i < bytes / 16; i += blockDim.x)
            d4[i] = __ldcs(s4 + i);
    } else {
        for (int i = threadIdx.x; i < bytes; i += blockDim.x)
            d[i] = s[i];
    }
}

// dst[off[k] ..] = src[k][0..bytes); idx32[pos[k]] = val[k] for the two per-frame status words
__global__ void scatter_kernel(const uint8_t* __restrict__ src, const unsigned long long* __restrict__ off, uint8_t* __restrict__ dst,
                               int bytes, const int32_t* __restrict__ v0, const int32_t* __restrict__ v1,
                               const int32_t* __restrict__ pos, int32_t* __restrict__ o0, int32_t* __restrict__ o1)
{
    const uint8_t* s = src + (size_t)blockIdx.x * bytes;
    uint8_t* d = dst + off[blockIdx.x];
    for (int i = threadIdx.x; i < bytes; i += blockDim.x)
        d[i] = s[i];
    if (threadIdx.x == 0) {
        const int f = pos[blockIdx.x];
        if (o0)
            o0[f] = v0[blockIdx.x];
        if (o1)
            o1[f] = v1[blockIdx.x];
    }
}

} // namespace

cudaError_t gather_launch(const uint8_t* src, const unsigned long long* off, uint8_t* dst, int bytes, int frames, cudaStream_t stream)
{
    if (frames <= 0)
        return cudaSuccess;
    gather_kernel<<<frames, 256, 0, stream>>>(src, off, dst, bytes);
    return cudaGetLastError();
}

cudaError_t scatter_launch(const uint8_t* src, const unsigned long long* off, uint8_t* dst, int bytes, int frames, const int32_t* v0,
                           const int32_t* v1, const int32_t* pos, int32_t* o0, int32_t* o1, cudaStream_t stream)
{
    if (frames <= 0)
        return cudaSuccess;
    scatter_kernel<<<frames, 128, 0, stream>>>(src, off, dst, bytes, v0, v1, pos, o0, o1);
    return cudaGetLastError();
}

// Forces the module that holds these kernels to be loaded now (CUDA loads lazily at the first launch, and that
// load can wait for the device to go idle -- which never happens while the persistent LDPC kernel of the
// streaming path is resident and waiting for input that the blocked host thread has yet to send).
cudaError_t mixed_preload()
{
    cudaFuncAttributes a;
    cudaError_t e;
    if ((e = cudaFuncGetAttributes(&a, gather_kernel)) != cudaSuccess)
        return e;
    if ((e = cudaFuncGetAttributes(&a, scatter_kernel)) != cudaSuccess)
        return e;
    return cudaSuccess;
}

} // namespace dvbs2b200
