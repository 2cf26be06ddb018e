// bb_kernel.cu -- BBFRAME descrambler and deheader on sm_100a: BCH output in, MPEG-TS bytes out.
//
// What it replaces:
//   lib/bbdescrambler_bb_impl.cc:51-82   XOR of every BBFRAME with the fixed PRBS (1 + x^14 + x^15)
//   lib/bbdeheader_bb_impl.cc:76-136     BBHEADER CRC-8 check, field extraction, validation
//   lib/bbdeheader_bb_impl.cc:144-261    TS packet extraction: re-synchronisation on SYNCD, partial packets
//                                        carried from one BBFRAME to the next, per-packet CRC-8, sync byte
//                                        restored, transport-error indicator set on CRC failure
//
// The reference walks the BBFRAMEs one after the other because the deheader carries state (synchronised?,
// bytes of a partial TS packet) from frame to frame.  That state is two small integers, so the work splits:
//   bb_header_kernel  one thread per BBFRAME: descramble the 10 header bytes, CRC-8, parse, validate;
//   bb_scan_kernel    one thread walks the per-frame records in order (a few instructions each) and
//                     writes a plan per frame: where its packets start in the output, how many, how
//                     many bytes of a carried partial packet go in front, which frame they come from;
//   bb_ts_kernel      one CTA per BBFRAME: descramble the DATAFIELD into shared memory, CRC-8 of each
//                     188-byte unit (one thread per packet, table in shared memory), coalesced stores.
// Everything is bytes and table lookups: HBM bound, kbch/8 bytes read and about as many written per frame.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "kernels.h"

namespace dvbs2b200 {

namespace {

constexpr int kTs = 188;        // TS_PACKET_LENGTH (lib/bbdeheader_bb_impl.h:20)
constexpr int kHdr = 10;        // BB_HEADER_LENGTH_BYTES
constexpr int kTsThreads = 128; // threads of bb_ts_kernel

// CRC-8, g(x) = x^8 + x^7 + x^6 + x^4 + x^2 + 1 (lib/bbdeheader_bb_impl.cc:55): T[i] = (i * x^8) mod g
__device__ __forceinline__ uint32_t crc8_entry(uint32_t i)
{
    uint32_t r = i << 8;
#pragma unroll
    for (int b = 15; b >= 8; --b)
        if (r & (1u << b))
            r ^= 0x1D5u << (b - 8);
    return r & 0xffu;
}

__global__ void bb_descramble_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const uint8_t* __restrict__ prbs,
                                     int frames, int kbytes)
{
    const size_t total = (size_t)frames * kbytes;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        out[i] = in[i] ^ __ldg(prbs + (i % kbytes));
}

// lib/bbdeheader_bb_impl.cc:76-136
__global__ void bb_header_kernel(const uint8_t* __restrict__ bb, const uint8_t* __restrict__ prbs, int scrambled, int frames,
                                 int kbytes, uint32_t* __restrict__ rec)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= frames)
        return;
    const uint8_t* p = bb + (size_t)f * kbytes;
    uint8_t h[kHdr];
    uint32_t crc = 0;
#pragma unroll
    for (int i = 0; i < kHdr; ++i) {
        h[i] = p[i] ^ (scrambled ? __ldg(prbs + i) : (uint8_t)0);
        crc = crc8_entry(crc ^ h[i]);
    }
    const uint32_t upl = ((uint32_t)h[2] << 8) | h[3];
    const uint32_t dfl = ((uint32_t)h[4] << 8) | h[5];
    const uint32_t syncd = ((uint32_t)h[7] << 8) | h[8];
    const uint32_t max_dfl = (uint32_t)kbytes * 8 - 80;
    const bool valid = crc == 0 && dfl <= max_dfl && (dfl % 8) == 0 && syncd <= dfl && upl == kTs * 8 && (syncd % 8) == 0;
    rec[f] = valid ? (0x80000000u | ((dfl / 8) << 16) | (syncd / 8)) : 0u;
}

// lib/bbdeheader_bb_impl.cc:144-261, the frame-to-frame state only.
__global__ void bb_scan_kernel(const uint32_t* __restrict__ rec, int frames, BbState* __restrict__ st, BbPlan* __restrict__ plan,
                               unsigned long long ts_cap_packets)
{
    __shared__ uint32_t s_rec[1024];
    int synched = st->synched;
    unsigned partial = st->partial;
    int src_frame = -1;   // where the bytes of the pending partial packet live: -1 = the carry buffer of the state
    unsigned src_off = 0; //   (offset into that frame's DATAFIELD)
    bool new_partial = false;
    unsigned long long packets = 0, drops = 0, gaps = 0;
    for (int base = 0; base < frames; base += 1024) {
        const int n = min(1024, frames - base);
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x)
            s_rec[i] = rec[base + i];
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int i = 0; i < n; ++i) {
                const uint32_t r = s_rec[i];
                BbPlan pl;
                pl.out_pkt = (uint32_t)packets;
                pl.n_pkts = 0;
                pl.p_in = 0;
                pl.skip = 0;
                pl.src_frame = -2;
                pl.src_off = 0;
                if (!(r >> 31)) {
                    synched = 0;
                    ++drops;
                } else {
                    unsigned df = (r >> 16) & 0x7fffu;
                    const unsigned syncd = r & 0xffffu;
                    if (partial > 0 && syncd != (unsigned)kTs - 1 - partial) {
                        synched = 0;
                        ++gaps;
                    }
                    if (!synched) {
                        // the reference's unsigned df_remaining wraps when syncd/8 + 1 > dfl/8 and it reads past
                        // the frame (undefined); here such a frame yields nothing
                        const unsigned skip = min(syncd + 1, df);
                        pl.skip = skip;
                        df -= skip;
                        synched = 1;
                        partial = 0;
                    }
                    if (df >= (unsigned)kTs) {
                        unsigned used = 0;
                        if (partial > 0) {
                            pl.p_in = partial;
                            pl.src_frame = src_frame;
                            pl.src_off = src_off;
                            used = kTs - partial;
                            partial = 0;
                            pl.n_pkts = 1;
                        }
                        pl.n_pkts += (df - used) / kTs;
                        df = (df - used) % kTs;
                        // df bytes remain at DATAFIELD offset skip + used + (n_pkts - (p_in ? 1 : 0)) * 188
                        if (df > 0)
                            src_off = pl.skip + used + (pl.n_pkts - (pl.p_in ? 1u : 0u)) * kTs;
                    } else if (df > 0) {
                        src_off = pl.skip;
                    }
                    if (df > 0) {
                        partial = df;
                        src_frame = base + i;
                        new_partial = true;
                    }
                    if (packets + pl.n_pkts > ts_cap_packets) // output buffer full: drop what does not fit
                        pl.n_pkts = (uint32_t)(ts_cap_packets > packets ? ts_cap_packets - packets : 0);
                    packets += pl.n_pkts;
                }
                plan[base + i] = pl;
            }
        }
    }
    if (threadIdx.x == 0) {
        st->synched = synched;
        st->partial = partial;
        st->read_idx = st->carry_idx;
        st->carry_src_frame = -1;
        if (new_partial && partial > 0) { // the CTA of that frame stores the bytes into the other carry buffer
            st->carry_idx ^= 1;
            st->carry_src_frame = src_frame;
            st->carry_src_off = src_off;
            st->carry_len = partial;
        }
        st->packet_cnt += packets;
        st->bbframe_cnt += (unsigned long long)frames;
        st->bbframe_drop_cnt += drops;
        st->bbframe_gap_cnt += gaps;
        st->produced = packets * kTs;
    }
}

__global__ void __launch_bounds__(kTsThreads) bb_ts_kernel(const uint8_t* __restrict__ bb, const uint8_t* __restrict__ prbs, int scrambled,
                                                            int kbytes, const BbPlan* __restrict__ plan, BbState* __restrict__ st,
                                                            uint8_t* __restrict__ ts)
{
    extern __shared__ uint8_t s_buf[]; // [p_in + datafield bytes of this frame's packets], then 64 flags
    __shared__ uint8_t s_crc[256];
    __shared__ uint32_t s_bad[kTsThreads / 32];
    const int f = blockIdx.x, tid = threadIdx.x;
    const BbPlan pl = plan[f];
    const uint8_t* frame = bb + (size_t)f * kbytes;
    // the partial packet this call leaves behind comes from this frame: copy it into the carry buffer
    if (st->carry_src_frame == f) {
        uint8_t* dst = st->carry[st->carry_idx];
        const unsigned off = kHdr + st->carry_src_off;
        for (unsigned i = tid; i < st->carry_len; i += kTsThreads)
            dst[i] = frame[off + i] ^ (scrambled ? __ldg(prbs + off + i) : (uint8_t)0);
    }
    if (pl.n_pkts == 0)
        return;
    for (int i = tid; i < 256; i += kTsThreads)
        s_crc[i] = (uint8_t)crc8_entry((uint32_t)i);
    // front: the partial packet carried over (from an earlier frame of this call, or from the previous call)
    if (pl.p_in) {
        if (pl.src_frame >= 0) {
            const uint8_t* src = bb + (size_t)pl.src_frame * kbytes;
            const unsigned off = kHdr + pl.src_off;
            for (unsigned i = tid; i < pl.p_in; i += kTsThreads)
                s_buf[i] = src[off + i] ^ (scrambled ? __ldg(prbs + off + i) : (uint8_t)0);
        } else {
            const uint8_t* src = st->carry[st->read_idx];
            for (unsigned i = tid; i < pl.p_in; i += kTsThreads)
                s_buf[i] = src[i];
        }
    }
    const unsigned total = pl.n_pkts * kTs, need = total - pl.p_in, off = kHdr + pl.skip;
    for (unsigned i = tid; i < need; i += kTsThreads)
        s_buf[pl.p_in + i] = frame[off + i] ^ (scrambled ? __ldg(prbs + off + i) : (uint8_t)0);
    __syncthreads();
    // CRC-8 over each 188-byte unit (187 payload bytes + the CRC that sits in the next sync position)
    uint8_t* s_tei = s_buf + ((total + 3) & ~3u);
    uint32_t nbad = 0;
    for (unsigned k = tid; k < pl.n_pkts; k += kTsThreads) {
        uint32_t c = 0;
        const uint8_t* pk = s_buf + k * kTs;
        for (int i = 0; i < kTs; ++i)
            c = s_crc[c ^ pk[i]];
        s_tei[k] = c ? 0x80 : 0;
        nbad += c != 0;
    }
    nbad = __reduce_add_sync(0xffffffffu, nbad);
    if ((tid & 31) == 0)
        s_bad[tid >> 5] = nbad;
    __syncthreads();
    if (tid == 0) {
        uint32_t b = 0;
        for (int w = 0; w < kTsThreads / 32; ++w)
            b += s_bad[w];
        if (b)
            atomicAdd(&st->error_cnt, (unsigned long long)b);
    }
    // out[0] = 0x47, out[1..187] = unit[0..186], TEI on a CRC failure (lib/bbdeheader_bb_impl.cc:232-240)
    uint8_t* out = ts + (size_t)pl.out_pkt * kTs;
    for (unsigned i = tid; i < total; i += kTsThreads) {
        const unsigned k = i / kTs, r = i - k * kTs;
        uint8_t v = r ? s_buf[k * kTs + r - 1] : (uint8_t)0x47;
        if (r == 1)
            v |= s_tei[k];
        out[i] = v;
    }
}

} // namespace

cudaError_t bb_descramble_launch(const uint8_t* in, uint8_t* out, const uint8_t* prbs, int frames, int kbytes, cudaStream_t stream)
{
    const size_t total = (size_t)frames * kbytes;
    const int grid = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
    bb_descramble_kernel<<<grid, 256, 0, stream>>>(in, out, prbs, frames, kbytes);
    return cudaGetLastError();
}

cudaError_t bb_deheader_launch(const BbLaunch& p, cudaStream_t stream)
{
    bb_header_kernel<<<(p.frames + 127) / 128, 128, 0, stream>>>(p.bb, p.prbs, p.scrambled, p.frames, p.kbytes, p.rec);
    bb_scan_kernel<<<1, 256, 0, stream>>>(p.rec, p.frames, p.state, p.plan, p.ts_cap / kTs);
    // shared memory: a frame contributes at most 187 carried bytes + its whole DATAFIELD, plus one flag per packet
    const size_t smem = (size_t)p.kbytes + 192 + 4 + (size_t)p.kbytes / kTs + 8;
    cudaError_t e = cudaFuncSetAttribute(bb_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess)
        return e;
    bb_ts_kernel<<<p.frames, kTsThreads, smem, stream>>>(p.bb, p.prbs, p.scrambled, p.kbytes, p.plan, p.state, p.ts);
    return cudaGetLastError();
}

} // namespace dvbs2b200
