// pl_kernel.cu -- PL descrambler + pilot-segment de-rotation on sm_100a.
//
// What it replaces: the output stage of plsync_cc_impl::handle_payload (lib/plsync_cc_impl.cc:639-802) with
// pl_descrambler::descramble (lib/pl_descrambler.cc:100-104) in front of it: every payload symbol is multiplied by
// the conjugate of the Gold-code scrambling factor exp(j R_n pi / 2) (pilot blocks count in n), the 36-symbol pilot
// blocks after every 16 slots are dropped, and the data symbols are de-rotated by a phase that starts at the
// PLHEADER phase estimate, advances by 2 pi fine_foffset per symbol when the frame is coarse corrected, and restarts
// from the estimate of the preceding pilot block at every 16-slot segment.  The result is the XFECFRAME the
// demapper consumes.
//
// The reference does the de-rotation with a serial recurrence (VOLK rotator: phase *= increment per sample); here
// every symbol evaluates its phase in closed form (segment base - k * increment, one sincosf), which is what makes the
// stage data parallel.  Agreement with the reference is therefore to float tolerance, not bit-exact (stated in the
// header and the tests).  Memory bound: 8 bytes in, 8 bytes out per data symbol, two symbols per thread (128-bit
// accesses), the 2-bit scrambling codes as one byte per symbol (33 KB per Gold code, L1/L2 resident).
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace dvbs2b200 {

namespace {

constexpr int kSegment = 16 * 90; // data symbols between pilot blocks
constexpr int kPilotBlock = 36;

__global__ void __launch_bounds__(256) pl_derotate_kernel(const PlLaunch p)
{
    const int frame = blockIdx.y;
    const PlFrameInfo info = p.info[frame];
    const float phase_inc = info.coarse_corrected ? 6.283185307179586f * info.fine_foffset : 0.0f;
    const int n_out = p.n_slots * 90;
    const float4* __restrict__ in = reinterpret_cast<const float4*>(p.payload + (size_t)frame * p.payload_len * 2);
    float4* __restrict__ out = reinterpret_cast<float4*>(p.out + (size_t)frame * n_out * 2);
    for (int pair = blockIdx.x * blockDim.x + threadIdx.x; pair < n_out / 2; pair += gridDim.x * blockDim.x) {
        const int m = 2 * pair;                   // output symbols m, m + 1: same segment (1440 is even)
        const int seg = p.has_pilots ? m / kSegment : 0;
        const int k = p.has_pilots ? m - seg * kSegment : m;
        const int i = m + kPilotBlock * seg;      // index in the payload (pilot blocks included): even
        const float base = (seg > 0 && info.coarse_corrected) ? info.pilot_phase[seg - 1] : info.plheader_phase;
        // without the restart (not coarse corrected) the increment is zero: the PLHEADER phase holds for the frame
        const float4 y = __ldcs(in + (i >> 1));
        const uint32_t r2 = *reinterpret_cast<const uint16_t*>(p.rn + i);
        float sr[2], si[2];
        sincosf(-(base + (float)k * phase_inc), &si[0], &sr[0]);
        sincosf(-(base + (float)(k + 1) * phase_inc), &si[1], &sr[1]);
        const float yr[2] = { y.x, y.z }, yi[2] = { y.y, y.w };
        float o[4];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const uint32_t r = (r2 >> (8 * u)) & 3u;
            // times {1, -j, -1, +j}[r]
            const float dr = (r & 1u) ? ((r & 2u) ? -yi[u] : yi[u]) : ((r & 2u) ? -yr[u] : yr[u]);
            const float di = (r & 1u) ? ((r & 2u) ? yr[u] : -yr[u]) : ((r & 2u) ? -yi[u] : yi[u]);
            o[2 * u] = dr * sr[u] - di * si[u];
            o[2 * u + 1] = dr * si[u] + di * sr[u];
        }
        __stcs(out + pair, make_float4(o[0], o[1], o[2], o[3]));
    }
}

} // namespace

cudaError_t pl_launch(const PlLaunch& p, cudaStream_t stream)
{
    if (p.frames <= 0)
        return cudaSuccess;
    const int pairs = p.n_slots * 45;
    for (int f0 = 0; f0 < p.frames; f0 += 32768) { // gridDim.y limit
        PlLaunch q = p;
        q.frames = p.frames - f0 < 32768 ? p.frames - f0 : 32768;
        q.payload = p.payload + (size_t)f0 * p.payload_len * 2;
        q.out = p.out + (size_t)f0 * p.n_slots * 90 * 2;
        q.info = p.info + f0;
        dim3 grid((pairs + 255) / 256, q.frames);
        pl_derotate_kernel<<<grid, 256, 0, stream>>>(q);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess)
            return e;
    }
    return cudaSuccess;
}

cudaError_t pl_preload()
{
    cudaFuncAttributes a;
    return cudaFuncGetAttributes(&a, pl_derotate_kernel);
}

} // namespace dvbs2b200
