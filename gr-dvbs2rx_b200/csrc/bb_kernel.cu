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
//   bb_scan_kernel    the frame-to-frame recurrence as a parallel scan (its transition functions have at
//                     most three distinct results and compose), then a plan per frame: where its packets
//                     start in the output, how many, how many bytes of a carried partial packet go in
//                     front, which frame they come from;
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

// ---- the frame-to-frame state of lib/bbdeheader_bb_impl.cc:144-261 as a parallel scan ----------------------
// State between BBFRAMEs: (synched, bytes of a pending partial packet), here sy << 8 | p.  What a frame does
// depends on the incoming state only through three classes: "synchronised, nothing pending" (p = 0),
// "synchronised and the pending bytes are exactly what SYNCD announces" (p = 187 - syncd/8), and everything
// else, which re-synchronises and forgets the past.  So the transition of a run of frames is a function
// with at most three distinct results, closed under composition -- a scan operator:
//   kind 0: only invalid headers so far: (sy, p) -> (0, p)
//   kind 1: constant r
//   kind 2: sy && p == 0 -> a0;  sy && p == x1 -> a1;  else r     (x1, the classes of the run's first frame)
struct BbFn {
    int kind, x1, a0, a1, r;
};
__device__ __forceinline__ int bb_apply(const BbFn& f, int st)
{
    if (f.kind == 0)
        return st & 0xff;
    if (f.kind == 1)
        return f.r;
    const int sy = st >> 8, p = st & 0xff;
    if (sy && p == 0)
        return f.a0;
    if (sy && f.x1 > 0 && p == f.x1)
        return f.a1;
    return f.r;
}
__device__ __forceinline__ BbFn bb_compose(const BbFn& f, const BbFn& g) // f first, then g
{
    BbFn h;
    if (f.kind == 0) {
        if (g.kind == 0)
            return f;
        h.kind = 1, h.x1 = 0, h.a0 = h.a1 = 0, h.r = g.r; // g sees sy = 0: its re-synchronisation result
        return h;
    }
    h.kind = f.kind, h.x1 = f.x1;
    h.r = bb_apply(g, f.r);
    h.a0 = f.kind == 2 ? bb_apply(g, f.a0) : 0;
    h.a1 = f.kind == 2 ? bb_apply(g, f.a1) : 0;
    return h;
}

// One BBFRAME given the incoming state: lib/bbdeheader_bb_impl.cc:163-253 without the byte moves.
struct BbStep {
    int st_out;          // state after the frame
    uint32_t n_pkts, p_in, skip;
    uint32_t left, left_off; // bytes of a new partial packet stored by this frame, their DATAFIELD offset
    int drop, gap;
};
__device__ __forceinline__ BbStep bb_step(uint32_t rec, int st)
{
    BbStep o;
    o.n_pkts = o.p_in = o.skip = o.left = o.left_off = 0;
    o.drop = o.gap = 0;
    int sy = st >> 8;
    unsigned p = (unsigned)(st & 0xff);
    if (!(rec >> 31)) {
        o.st_out = (int)p; // synched = false, the partial count stays
        o.drop = 1;
        return o;
    }
    unsigned df = (rec >> 16) & 0x7fffu;
    const unsigned syncd = rec & 0xffffu;
    if (p > 0 && syncd != (unsigned)kTs - 1 - p) {
        sy = 0;
        o.gap = 1;
    }
    if (!sy) {
        // the reference's unsigned df_remaining wraps when syncd/8 + 1 > dfl/8 and it reads past the frame
        // (undefined); here such a frame yields nothing
        o.skip = min(syncd + 1, df);
        df -= o.skip;
        p = 0;
    }
    if (df >= (unsigned)kTs) {
        o.p_in = p;
        o.n_pkts = (df + p) / kTs;
        o.left = (df + p) % kTs;
        o.left_off = o.skip + df - o.left;
        p = o.left;
    } else if (df > 0) {
        o.left = df;
        o.left_off = o.skip;
        p = df;
    }
    o.st_out = 0x100 | (int)p;
    return o;
}
// the transition of one frame as a function of the incoming state
__device__ __forceinline__ BbFn bb_fn_of(uint32_t rec)
{
    BbFn f;
    f.x1 = f.a0 = f.a1 = f.r = 0;
    if (!(rec >> 31)) {
        f.kind = 0;
        return f;
    }
    const int syncd = (int)(rec & 0xffffu);
    f.kind = 2;
    f.x1 = (syncd <= kTs - 2) ? kTs - 1 - syncd : 0; // the one pending count that continues the stream
    f.a0 = bb_step(rec, 0x100).st_out;
    f.a1 = f.x1 ? bb_step(rec, 0x100 | f.x1).st_out : 0;
    f.r = bb_step(rec, 0).st_out;
    return f;
}

constexpr int kScanThreads = 1024;

__global__ void __launch_bounds__(kScanThreads) bb_scan_kernel(const uint32_t* __restrict__ rec, int frames, BbState* __restrict__ st,
                                                                BbPlan* __restrict__ plan, unsigned long long ts_cap_packets)
{
    __shared__ BbFn s_fn[kScanThreads];
    __shared__ unsigned long long s_pk[kScanThreads], s_wr[kScanThreads]; // packets; last writer (frame + 1) << 32 | offset
    __shared__ unsigned int s_drop[kScanThreads / 32], s_gap[kScanThreads / 32];
    const int tid = threadIdx.x;
    const int chunk = (frames + kScanThreads - 1) / kScanThreads;
    const int lo = min(frames, tid * chunk), hi = min(frames, lo + chunk);
    // pass 1: the transition function of this thread's run of frames; inclusive scan over the threads
    BbFn fn;
    fn.kind = 0, fn.x1 = fn.a0 = fn.a1 = fn.r = 0;
    bool first = true;
    for (int i = lo; i < hi; ++i) {
        const BbFn g = bb_fn_of(rec[i]);
        fn = first ? g : bb_compose(fn, g);
        first = false;
    }
    // an empty run is the identity; kind 0 only clears `synched`, so mark emptiness separately
    const bool empty = lo >= hi;
    s_fn[tid] = fn;
    s_pk[tid] = empty ? 1ull : 0ull; // reused as the "empty" flag during the function scan
    __syncthreads();
    for (int off = 1; off < kScanThreads; off <<= 1) {
        BbFn left;
        bool take = false, left_empty = true;
        if (tid >= off) {
            left = s_fn[tid - off];
            left_empty = s_pk[tid - off] != 0;
            take = true;
        }
        __syncthreads();
        if (take && !left_empty) {
            const bool me_empty = s_pk[tid] != 0;
            s_fn[tid] = me_empty ? left : bb_compose(left, s_fn[tid]);
            s_pk[tid] = 0;
        }
        __syncthreads();
    }
    // incoming state of this thread's run = prefix of the runs before it applied to the call's initial state
    const int st0 = (st->synched ? 0x100 : 0) | (int)(st->partial & 0xff);
    int state = st0;
    if (tid > 0 && s_pk[tid - 1] == 0)
        state = bb_apply(s_fn[tid - 1], st0);
    const int st_final = (s_pk[kScanThreads - 1] == 0) ? bb_apply(s_fn[kScanThreads - 1], st0) : st0;
    __syncthreads();
    // pass 2: walk the run with the known state: packets, last frame that stored a partial packet, counters
    unsigned long long pk = 0, wr = 0;
    unsigned int drops = 0, gaps = 0;
    {
        int s = state;
        for (int i = lo; i < hi; ++i) {
            const BbStep o = bb_step(rec[i], s);
            s = o.st_out;
            pk += o.n_pkts;
            if (o.left)
                wr = ((unsigned long long)(i + 1) << 32) | o.left_off;
            drops += o.drop;
            gaps += o.gap;
        }
    }
    s_pk[tid] = pk;
    s_wr[tid] = wr;
    drops = __reduce_add_sync(0xffffffffu, drops);
    gaps = __reduce_add_sync(0xffffffffu, gaps);
    if ((tid & 31) == 0)
        s_drop[tid >> 5] = drops, s_gap[tid >> 5] = gaps;
    __syncthreads();
    for (int off = 1; off < kScanThreads; off <<= 1) {
        unsigned long long a = 0, w = 0;
        if (tid >= off)
            a = s_pk[tid - off], w = s_wr[tid - off];
        __syncthreads();
        s_pk[tid] += a;
        s_wr[tid] = max(s_wr[tid], w); // frame indices rise: the later writer wins
        __syncthreads();
    }
    // pass 3: the plans
    {
        unsigned long long base = tid ? s_pk[tid - 1] : 0ull, writer = tid ? s_wr[tid - 1] : 0ull;
        int s = state;
        for (int i = lo; i < hi; ++i) {
            const BbStep o = bb_step(rec[i], s);
            s = o.st_out;
            BbPlan pl;
            pl.out_pkt = (uint32_t)min(base, ts_cap_packets);
            pl.n_pkts = (uint32_t)min((unsigned long long)o.n_pkts, ts_cap_packets > base ? ts_cap_packets - base : 0ull);
            pl.p_in = o.p_in;
            pl.skip = o.skip;
            pl.src_frame = o.p_in ? (writer ? (int32_t)(writer >> 32) - 1 : -1) : -2;
            pl.src_off = (uint32_t)(writer & 0xffffffffu);
            plan[i] = pl;
            base += o.n_pkts;
            if (o.left)
                writer = ((unsigned long long)(i + 1) << 32) | o.left_off;
        }
    }
    if (tid == 0) {
        const unsigned long long total = min(s_pk[kScanThreads - 1], ts_cap_packets), writer = s_wr[kScanThreads - 1];
        unsigned long long d = 0, g = 0;
        for (int w = 0; w < kScanThreads / 32; ++w)
            d += s_drop[w], g += s_gap[w];
        const unsigned partial = (unsigned)(st_final & 0xff);
        st->synched = st_final >> 8;
        st->partial = partial;
        st->read_idx = st->carry_idx;
        st->carry_src_frame = -1;
        if (writer && partial > 0) { // the CTA of that frame stores the bytes into the other carry buffer
            st->carry_idx ^= 1;
            st->carry_src_frame = (int)(writer >> 32) - 1;
            st->carry_src_off = (unsigned)(writer & 0xffffffffu);
            st->carry_len = partial;
        }
        st->packet_cnt += total;
        st->bbframe_cnt += (unsigned long long)frames;
        st->bbframe_drop_cnt += d;
        st->bbframe_gap_cnt += g;
        st->produced = total * kTs;
    }
}

// four bytes from an arbitrarily aligned address, as two aligned words (the neighbour's word is an L1 hit)
__device__ __forceinline__ uint32_t ld32_unaligned(const uint8_t* p)
{
    const uint32_t a = (uint32_t)((uintptr_t)p & 3);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(p - a);
    const uint32_t lo = __ldg(w), hi = a ? __ldg(w + 1) : 0u;
    return __funnelshift_r(lo, hi, 8 * a);
}

constexpr int kFront = 192; // shared-memory offset of the frame's own DATAFIELD bytes; carried bytes sit right before it

__global__ void __launch_bounds__(kTsThreads) bb_ts_kernel(const uint8_t* __restrict__ bb, const uint8_t* __restrict__ prbs, int scrambled,
                                                            int kbytes, const BbPlan* __restrict__ plan, BbState* __restrict__ st,
                                                            uint8_t* __restrict__ ts)
{
    extern __shared__ __align__(16) uint8_t s_buf[]; // [kFront - p_in, kFront): carried bytes; [kFront, ..): this frame's bytes; flags
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
    uint8_t* const pkts = s_buf + kFront - pl.p_in; // first packet of this frame
    if (pl.p_in) {
        if (pl.src_frame >= 0) {
            const uint8_t* src = bb + (size_t)pl.src_frame * kbytes;
            const unsigned off = kHdr + pl.src_off;
            for (unsigned i = tid; i < pl.p_in; i += kTsThreads)
                pkts[i] = src[off + i] ^ (scrambled ? __ldg(prbs + off + i) : (uint8_t)0);
        } else {
            const uint8_t* src = st->carry[st->read_idx];
            for (unsigned i = tid; i < pl.p_in; i += kTsThreads)
                pkts[i] = src[i];
        }
    }
    // this frame's bytes: 32-bit words, descrambled on the way in
    const unsigned total = pl.n_pkts * kTs, need = total - pl.p_in, off = kHdr + pl.skip;
    // (the very last frame keeps its final word for the byte loop: no read past the end of the caller's buffer)
    const unsigned nvec = (f == (int)gridDim.x - 1 && need >= 4) ? need / 4 - 1 : need / 4;
    for (unsigned w = tid; w < nvec; w += kTsThreads) {
        uint32_t v = ld32_unaligned(frame + off + 4 * w);
        if (scrambled)
            v ^= ld32_unaligned(prbs + off + 4 * w);
        *reinterpret_cast<uint32_t*>(s_buf + kFront + 4 * w) = v;
    }
    for (unsigned i = 4 * nvec + tid; i < need; i += kTsThreads)
        s_buf[kFront + i] = frame[off + i] ^ (scrambled ? __ldg(prbs + off + i) : (uint8_t)0);
    __syncthreads();
    // CRC-8 over each 188-byte unit (187 payload bytes + the CRC that sits in the next sync position)
    uint8_t* s_tei = s_buf + kFront + ((need + 3) & ~3u);
    uint32_t nbad = 0;
    for (unsigned k = tid; k < pl.n_pkts; k += kTsThreads) {
        uint32_t c = 0;
        const uint8_t* pk = pkts + k * kTs;
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
    // out[0] = 0x47, out[1..187] = unit[0..186], TEI on a CRC failure (lib/bbdeheader_bb_impl.cc:232-240);
    // 188 = 4 * 47: a 32-bit word of the output never straddles two packets
    uint32_t* out = reinterpret_cast<uint32_t*>(ts + (size_t)pl.out_pkt * kTs);
    for (unsigned j = tid; j < total / 4; j += kTsThreads) {
        const unsigned k = j / 47, r0 = (j - k * 47) * 4;
        const uint8_t* pk = pkts + k * kTs;
        uint32_t v;
        if (r0 == 0)
            v = 0x47u | ((uint32_t)(pk[0] | s_tei[k]) << 8) | ((uint32_t)pk[1] << 16) | ((uint32_t)pk[2] << 24);
        else
            v = (uint32_t)pk[r0 - 1] | ((uint32_t)pk[r0] << 8) | ((uint32_t)pk[r0 + 1] << 16) | ((uint32_t)pk[r0 + 2] << 24);
        out[j] = v;
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
    bb_scan_kernel<<<1, kScanThreads, 0, stream>>>(p.rec, p.frames, p.state, p.plan, p.ts_cap / kTs);
    // shared memory: a frame contributes at most 187 carried bytes + its whole DATAFIELD, plus one flag per packet
    const size_t smem = (size_t)kFront + p.kbytes + 8 + (size_t)p.kbytes / kTs + 8;
    cudaError_t e = cudaFuncSetAttribute(bb_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess)
        return e;
    bb_ts_kernel<<<p.frames, kTsThreads, smem, stream>>>(p.bb, p.prbs, p.scrambled, p.kbytes, p.plan, p.state, p.ts);
    return cudaGetLastError();
}

// Forces the module that holds these kernels to be loaded now (CUDA loads lazily at the first launch, and that
// load can wait for the device to go idle -- which never happens while the persistent LDPC kernel of the
// streaming path is resident and waiting for input that the blocked host thread has yet to send).
cudaError_t bb_preload()
{
    cudaFuncAttributes a;
    cudaError_t e;
    if ((e = cudaFuncGetAttributes(&a, bb_descramble_kernel)) != cudaSuccess)
        return e;
    if ((e = cudaFuncGetAttributes(&a, bb_header_kernel)) != cudaSuccess)
        return e;
    if ((e = cudaFuncGetAttributes(&a, bb_scan_kernel)) != cudaSuccess)
        return e;
    if ((e = cudaFuncGetAttributes(&a, bb_ts_kernel)) != cudaSuccess)
        return e;
    return cudaSuccess;
}

} // namespace dvbs2b200
