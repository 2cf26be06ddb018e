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
// header and the tests).  Memory bound: 8 bytes in, 8 bytes out per data symbol, two symbols per 128-bit access, four
// loads in flight per thread, one sincosf per symbol pair; the 2-bit scrambling codes as one byte per symbol (33 KB per
// Gold code, L1/L2 resident).  Runs at the HBM copy rate.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace dvbs2b200 {

namespace {

constexpr int kSegment = 16 * 90; // data symbols between pilot blocks
constexpr int kPilotBlock = 36;

constexpr int kPlLoads = 4;

__global__ void __launch_bounds__(256) pl_derotate_kernel(const PlLaunch p)
{
    const int frame = blockIdx.y;
    // scalars in registers; the pilot-block phases are indexed by segment and stay where they are (read-only path)
    const PlFrameInfo* __restrict__ fi = p.info + frame;
    const float plheader_phase = fi->plheader_phase;
    const bool coarse_corrected = fi->coarse_corrected != 0;
    const float phase_inc = coarse_corrected ? 6.283185307179586f * fi->fine_foffset : 0.0f;
    const int n_out = p.n_slots * 90;
    const float4* __restrict__ in = reinterpret_cast<const float4*>(p.payload + (size_t)frame * p.payload_len * 2);
    float4* __restrict__ out = reinterpret_cast<float4*>(p.out + (size_t)frame * n_out * 2);
    // the rotation of the second symbol of a pair is the first one's advanced by one increment (angle addition: one
    // sincosf per pair instead of two)
    float inc_s, inc_c;
    sincosf(-phase_inc, &inc_s, &inc_c);
    const int pairs = n_out / 2;
    // kPlLoads 128-bit loads in flight per thread before the arithmetic (one load per ~60 instructions does not cover
    // the DRAM latency)
    for (int pair0 = blockIdx.x * (kPlLoads * blockDim.x) + threadIdx.x; pair0 < pairs; pair0 += gridDim.x * kPlLoads * blockDim.x) {
        float4 y[kPlLoads];
        uint32_t rr[kPlLoads];
#pragma unroll
        for (int h = 0; h < kPlLoads; ++h) {
            const int pair = pair0 + h * (int)blockDim.x;
            const int m = 2 * pair;
            const int seg = p.has_pilots ? m / kSegment : 0;
            const int i = m + kPilotBlock * seg; // index in the payload (pilot blocks included): even
            const bool ok = pair < pairs;
            y[h] = ok ? __ldcs(in + (i >> 1)) : make_float4(0.f, 0.f, 0.f, 0.f);
            rr[h] = ok ? (uint32_t) * reinterpret_cast<const uint16_t*>(p.rn + i) : 0u;
        }
#pragma unroll
        for (int h = 0; h < kPlLoads; ++h) {
            const int pair = pair0 + h * (int)blockDim.x;
            if (pair >= pairs)
                break;
            const int m = 2 * pair;                   // output symbols m, m + 1: same segment (1440 is even)
            const int seg = p.has_pilots ? m / kSegment : 0;
            const int k = p.has_pilots ? m - seg * kSegment : m;
            // without the restart (not coarse corrected) the increment is zero: the PLHEADER phase holds for the frame
            const float base = (seg > 0 && coarse_corrected) ? __ldg(&fi->pilot_phase[seg - 1]) : plheader_phase;
            float sr[2], si[2];
            sincosf(-(base + (float)k * phase_inc), &si[0], &sr[0]);
            sr[1] = sr[0] * inc_c - si[0] * inc_s;
            si[1] = si[0] * inc_c + sr[0] * inc_s;
            const float yr[2] = { y[h].x, y[h].z }, yi[2] = { y[h].y, y[h].w };
            float o[4];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const uint32_t r = (rr[h] >> (8 * u)) & 3u;
                // times {1, -j, -1, +j}[r]
                const float dr = (r & 1u) ? ((r & 2u) ? -yi[u] : yi[u]) : ((r & 2u) ? -yr[u] : yr[u]);
                const float di = (r & 1u) ? ((r & 2u) ? yr[u] : -yr[u]) : ((r & 2u) ? -yi[u] : yi[u]);
                o[2 * u] = dr * sr[u] - di * si[u];
                o[2 * u + 1] = dr * si[u] + di * sr[u];
            }
            __stcs(out + pair, make_float4(o[0], o[1], o[2], o[3]));
        }
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
        dim3 grid((pairs + 256 * kPlLoads - 1) / (256 * kPlLoads), q.frames);
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
