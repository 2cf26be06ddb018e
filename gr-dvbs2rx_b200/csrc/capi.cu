// capi.cu -- the C ABI of libdvbs2_b200.so (include/dvbs2_b200.h): handles, staging, launches.
// No CPU fallback: without a CUDA device every compute entry point fails with DVBS2B200_ECUDA.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/dvbs2_b200.h"
#include "code_tables.h"
#include "kernels.h"

using namespace dvbs2b200;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg)
{
    g_err = msg;
    return code;
}
int cuda_fail(cudaError_t e, const char* what)
{
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return DVBS2B200_ECUDA;
}
#define CU(call)                             \
    do {                                     \
        cudaError_t e__ = (call);            \
        if (e__ != cudaSuccess)              \
            return cuda_fail(e__, #call);    \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap)
            return DVBS2B200_OK;
        if (p)
            cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            g_err = std::string("cudaMalloc: ") + cudaGetErrorString(e);
            return DVBS2B200_ENOMEM;
        }
        cap = want;
        return DVBS2B200_OK;
    }
    void release()
    {
        if (p)
            cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct HostBuf { // pinned host memory
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap)
            return DVBS2B200_OK;
        if (p)
            cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocDefault);
        if (e != cudaSuccess) {
            g_err = std::string("cudaHostAlloc: ") + cudaGetErrorString(e);
            return DVBS2B200_ENOMEM;
        }
        cap = bytes;
        return DVBS2B200_OK;
    }
    void release()
    {
        if (p)
            cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

// true if the driver can DMA straight from / to this host pointer (pinned or registered memory)
bool host_ptr_is_pinned(const void* ptr)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

} // namespace

struct dvbs2b200_code {
    int device = 0;
    int sm_count = 0;
    int smem_optin = 0;
    cudaStream_t stream = nullptr;                    // compute (and default staging) stream
    cudaStream_t s_in = nullptr, s_out = nullptr;     // copy streams of the pipelined host path
    std::vector<uint8_t> blob;
    BlobHeader hdr;
    uint8_t* d_blob = nullptr;
    size_t ldpc_smem = 0;
    int ldpc_ctas = 0;      // resident LDPC CTAs per SM
    uint64_t launches = 0;
    // staging for the host-pointer entry points
    DevBuf d_in, d_mid, d_out, d_post, d_i32a, d_i32b, d_n0, d_llr, d_sync, d_scratch, d_flag, d_prof, d_next;
    // BB layer: descrambling sequence, deheader stream state, per-call scratch, TS output staging
    DevBuf d_prbs, d_bbstate, d_bbrec, d_bbplan, d_ts;
    DevBuf d_points; // table-driven demapper: constellation [32][2] floats + row offsets [5] ints
    std::vector<uint8_t> points_cache; // what d_points holds (the table is re-sent only when it changes)
    bool bb_ready = false;
    // The handle's device scratch (check-node state, intermediates, BB state) is shared by all its calls: work
    // issued on one stream must not overlap work issued on another.  Every call records ev_last on the stream
    // it used; a call on a different stream waits for it first.
    cudaStream_t last_stream = nullptr;
    cudaEvent_t ev_last = nullptr;
    bool ev_valid = false;
    // pinned staging for pageable host buffers: a ring of input slots, one output area, the arrival counters
    HostBuf h_ring, h_out, h_cnt;
    cudaEvent_t ev_slot[3] = { nullptr, nullptr, nullptr };
    DevBuf d_err; // error word of the streaming LDPC launch (input never arrived)
};

// A set of codes on one device for mixed-MODCOD (VCM/ACM) batches.
struct dvbs2b200_mixed {
    int device = 0;
    std::vector<dvbs2b200_code*> codes;
    cudaStream_t stream = nullptr; // batch-wide copies
    DevBuf d_in, d_out, d_tr, d_co;
    struct PerCode {
        DevBuf d_in_off, d_out_off, d_pos, d_stage_in, d_stage_out, d_tr, d_co;
    };
    std::vector<PerCode> per;
};

// One code on several devices of this process: a batch is split into contiguous frame ranges, one per device.
struct dvbs2b200_multi {
    std::vector<dvbs2b200_code*> codes; // one handle (tables, streams, scratch) per device
};

namespace {

void info_from_header(const BlobHeader& h, dvbs2b200_code_info* info)
{
    info->table = h.table;
    info->n_ldpc = h.N;
    info->k_ldpc = h.K;
    info->q = h.q;
    info->n_circ = h.n_circ;
    info->links_total = h.links_total;
    info->max_cn_deg = h.max_cn_deg;
    info->kbch = h.kbch;
    info->nbch = h.nbch;
    info->t = h.t;
    info->gf_m = h.gf_m;
}

int create_from_blob(dvbs2b200_code** out, int device, std::vector<uint8_t>&& blob, bool validated = false)
{
    std::string err;
    if (!validated && !validate_blob(blob.data(), blob.size(), err))
        return fail(DVBS2B200_EINVAL, err);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0)
        return fail(DVBS2B200_ECUDA, std::string("no CUDA device available (") + cudaGetErrorString(e) +
                                         "); libdvbs2_b200 has no CPU fallback");
    if (device < 0 || device >= ndev)
        return fail(DVBS2B200_EINVAL, "device index out of range");
    dvbs2b200_code* h = new (std::nothrow) dvbs2b200_code();
    if (!h)
        return fail(DVBS2B200_ENOMEM, "out of host memory");
    h->device = device;
    h->blob = std::move(blob);
    memcpy(&h->hdr, h->blob.data(), sizeof(BlobHeader));
    auto bail = [&](int code) {
        dvbs2b200_code_destroy(h);
        return code;
    };
    if (cudaSetDevice(device) != cudaSuccess)
        return bail(fail(DVBS2B200_ECUDA, "cudaSetDevice failed"));
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&h->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking)) != cudaSuccess)
        return bail(cuda_fail(e, "cudaStreamCreate"));
    if ((e = cudaMalloc((void**)&h->d_blob, h->blob.size())) != cudaSuccess)
        return bail(cuda_fail(e, "cudaMalloc(tables)"));
    if ((e = cudaMemcpy(h->d_blob, h->blob.data(), h->blob.size(), cudaMemcpyHostToDevice)) != cudaSuccess)
        return bail(cuda_fail(e, "cudaMemcpy(tables)"));
    // a pageable-source copy may return while its DMA is still in flight on the legacy stream; the kernels
    // run on non-blocking streams that do not order against it
    if ((e = cudaDeviceSynchronize()) != cudaSuccess)
        return bail(cuda_fail(e, "cudaDeviceSynchronize(tables)"));
    if ((e = bch_preload()) != cudaSuccess || (e = bb_preload()) != cudaSuccess || (e = demap_preload()) != cudaSuccess ||
        (e = mixed_preload()) != cudaSuccess)
        return bail(cuda_fail(e, "kernel preload"));
    h->ldpc_smem = ldpc_smem_bytes(h->hdr.N, h->hdr.smem_bytes, h->hdr.chain_scratch, nullptr);
    if (h->hdr.max_cnt <= 28 && h->ldpc_smem <= (size_t)h->smem_optin)
        h->ldpc_ctas = ldpc_ctas_per_sm(h->hdr.N, h->hdr.max_cnt, h->hdr.uniform_cnt != 0, h->ldpc_smem);
    if (getenv("DVBS2B200_DEBUG"))
        fprintf(stderr, "[dvbs2b200] table %d: ldpc smem %zu B, %d CTAs/SM, %d state words per check-node pair\n", h->hdr.table,
                h->ldpc_smem, h->ldpc_ctas, h->hdr.msg_words);
    if (const char* cap = getenv("DVBS2B200_LDPC_CTAS_PER_SM")) { // tuning knob: cap the resident CTAs per SM
        int c = atoi(cap);
        if (c > 0 && c < h->ldpc_ctas)
            h->ldpc_ctas = c;
    }
    *out = h;
    return DVBS2B200_OK;
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev)
    {
        cudaGetDevice(&prev);
        if (prev != dev)
            cudaSetDevice(dev);
        else
            prev = -1;
    }
    ~DeviceGuard()
    {
        if (prev >= 0)
            cudaSetDevice(prev);
    }
};

// Stream ordering of a handle's calls (see dvbs2b200_code::ev_last).
struct StreamOrder {
    dvbs2b200_code* h;
    cudaStream_t s;
    StreamOrder(dvbs2b200_code* h_, cudaStream_t s_) : h(h_), s(s_)
    {
        if (h->ev_valid && h->last_stream != s)
            cudaStreamWaitEvent(s, h->ev_last, 0);
    }
    ~StreamOrder()
    {
        if (!h->ev_last && cudaEventCreateWithFlags(&h->ev_last, cudaEventDisableTiming) != cudaSuccess) {
            h->ev_last = nullptr;
            cudaStreamSynchronize(s); // no event to order by: fall back to completing the work
            h->ev_valid = false;
            return;
        }
        cudaEventRecord(h->ev_last, s);
        h->last_stream = s;
        h->ev_valid = true;
    }
};

int ldpc_out_bytes(const BlobHeader& h, int output_mode) { return (output_mode ? h.kldpc_out : h.N) / 8; }

// grid size: persistent CTAs, as many as are resident at once; in group mode a multiple of the group
int ldpc_grid(const dvbs2b200_code* h, int frames, int group)
{
    const int resident = h->sm_count * h->ldpc_ctas;
    int grid = std::min(frames, resident);
    if (group > 1)
        grid = std::min(frames, (resident / group) * group);
    return grid;
}

int ldpc_dev(dvbs2b200_code* h, const int8_t* d_llr, int frames, int max_trials, int term_group, int output_mode,
             uint8_t* d_hard, int8_t* d_llr_post, int32_t* d_trials_left, cudaStream_t stream,
             const unsigned int* d_ready = nullptr, int ready_chunk = 0, unsigned int* d_err = nullptr)
{
    const BlobHeader& hd = h->hdr;
    if (frames < 0 || max_trials < 0)
        return fail(DVBS2B200_EINVAL, "negative frames/max_trials");
    if (frames == 0)
        return DVBS2B200_OK;
    if (!d_llr)
        return fail(DVBS2B200_EINVAL, "llr is null");
    if (term_group != 0 && term_group != 1 && term_group != 16 && term_group != 32)
        return fail(DVBS2B200_EINVAL, "term_group must be 0, 16 or 32");
    if (term_group > 1 && frames % term_group)
        return fail(DVBS2B200_EINVAL, "frames must be a multiple of term_group");
    if (hd.max_cnt > 28)
        return fail(DVBS2B200_EUNSUPPORTED, "more than 28 data links per check node");
    if (h->ldpc_ctas <= 0)
        return fail(DVBS2B200_ECUDA, "LDPC kernel cannot be resident on this device (shared memory / registers)");
    if (term_group > 1 && h->sm_count * h->ldpc_ctas < term_group)
        return fail(DVBS2B200_EUNSUPPORTED, "device cannot hold term_group frames at once");
    if (((uintptr_t)d_llr & 3) || ((uintptr_t)d_llr_post & 3))
        return fail(DVBS2B200_EINVAL, "llr buffers must be 4-byte aligned");
    if (max_trials == 0)
        max_trials = 25; // lib/ldpc_decoder_bb_impl.cc:391,402
    LdpcLaunch p;
    memset(&p, 0, sizeof(p));
    p.N = hd.N;
    p.K = hd.K;
    p.R = hd.R;
    p.q = hd.q;
    p.n_circ = hd.n_circ;
    p.n_steps = hd.n_steps_total;
    p.tab = h->d_blob + hd.smem_off;
    p.tab_bytes = hd.smem_bytes;
    p.work = reinterpret_cast<const uint16_t*>(h->d_blob + hd.order_off);
    size_t smem = ldpc_smem_bytes(hd.N, hd.smem_bytes, hd.chain_scratch, &p);
    const int group = term_group > 1 ? term_group : 0;
    const int grid = ldpc_grid(h, frames, group);
    {
        int rc = h->d_scratch.ensure((size_t)grid * (hd.R / 2) * hd.msg_words * sizeof(uint32_t));
        if (rc)
            return rc;
        p.msg_scratch = (uint32_t*)h->d_scratch.p;
    }
    p.llr = d_llr;
    p.ready = d_ready;
    p.ready_chunk = ready_chunk;
    p.err = d_err;
    p.frames = frames;
    p.max_trials = max_trials;
    p.group = group;
    p.hard = d_hard;
    p.out_bytes = ldpc_out_bytes(hd, output_mode);
    p.llr_post = d_llr_post;
    p.trials_left = d_trials_left;
    p.sm_count = h->sm_count;
    p.two = 2u;
    p.four = 4u;
    p.c30 = 1u << 30;
    p.c16 = 1u << 16;
    p.c32 = 32u;
    p.neg1 = 0xffffffffu;
    if (p.group) {
        size_t words = (size_t)(frames / p.group) * (max_trials + 2);
        int rc = h->d_sync.ensure(words * sizeof(unsigned));
        if (rc)
            return rc;
        CU(cudaMemsetAsync(h->d_sync.p, 0, words * sizeof(unsigned), stream));
        p.gsync = (unsigned*)h->d_sync.p;
    }
    else {
        // frames converge after different numbers of iterations: hand them out dynamically
        int rc = h->d_next.ensure(16);
        if (rc)
            return rc;
        CU(cudaMemsetAsync(h->d_next.p, 0, 4, stream));
        p.next_frame = (unsigned int*)h->d_next.p;
    }
#ifdef DVBS2_PHASE_PROFILE
    const char* prof_path = getenv("DVBS2B200_PHASE_PROFILE"); // diagnostics build: per-CTA cycles per phase
#else
    const char* prof_path = nullptr;
#endif
    if (prof_path) {
        int rc = h->d_prof.ensure((size_t)grid * 16 * sizeof(unsigned long long));
        if (rc)
            return rc;
        p.prof = (unsigned long long*)h->d_prof.p;
        CU(cudaMemsetAsync(p.prof, 0, (size_t)grid * 16 * sizeof(unsigned long long), stream));
    }
    cudaError_t e = ldpc_launch(p, hd.max_cnt, hd.uniform_cnt != 0, grid, smem, stream);
    if (e != cudaSuccess)
        return cuda_fail(e, "ldpc_launch");
    if (prof_path) {
        std::vector<unsigned long long> host((size_t)grid * 16);
        CU(cudaStreamSynchronize(stream));
        CU(cudaMemcpy(host.data(), h->d_prof.p, host.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        if (FILE* f = fopen(prof_path, "w")) {
            static const char* names[15] = { "load", "syndrome_pass", "pair_steps", "split_phase1", "level_serial_phase_thread0", "split_phase3",
                                             "iteration_end", "output", "total", "n_pair_steps", "n_split_steps", "level_serial_phase_last_node",
                                             "n_levels", "chain_serial_phase", "n_chain_nodes_per_walker" };
            for (int k = 0; k < 15; ++k) {
                double sum = 0;
                for (int b = 0; b < grid; ++b)
                    sum += (double)host[(size_t)b * 16 + k];
                fprintf(f, "%s %.0f\n", names[k], sum / grid);
            }
            fprintf(f, "ctas %d frames %d\n", grid, frames);
            fclose(f);
        }
    }
    h->launches += 1;
    return DVBS2B200_OK;
}

int bch_dev(dvbs2b200_code* h, const uint8_t* d_cw, int cw_stride, int frames, uint8_t* d_msg, int32_t* d_corr,
            cudaStream_t stream)
{
    const BlobHeader& hd = h->hdr;
    if (frames < 0)
        return fail(DVBS2B200_EINVAL, "negative frames");
    if (frames == 0)
        return DVBS2B200_OK;
    if (!d_cw || !d_msg)
        return fail(DVBS2B200_EINVAL, "null buffer");
    if (hd.kbch <= 0 || hd.kbch % 8 || hd.nbch % 8)
        return fail(DVBS2B200_EUNSUPPORTED, "BCH k and n must be multiples of 8 (lib/bch.cc:19-24)");
    BchLaunch p;
    memset(&p, 0, sizeof(p));
    p.cw = d_cw;
    p.msg = d_msg;
    p.corrections = d_corr;
    p.frames = frames;
    p.n = hd.nbch;
    p.k = hd.kbch;
    p.t = hd.t;
    p.m = hd.gf_m;
    p.shorten = hd.bch_shorten;
    p.antilog = reinterpret_cast<const uint16_t*>(h->d_blob + hd.antilog_off);
    p.log = reinterpret_cast<const uint16_t*>(h->d_blob + hd.log_off);
    p.cw_stride = cw_stride;
    p.msg_stride = hd.kbch / 8;
    cudaError_t e = bch_launch(p, stream);
    if (e != cudaSuccess)
        return cuda_fail(e, "bch_launch");
    h->launches += 1;
    return DVBS2B200_OK;
}

int bits_per_symbol(int constellation)
{
    if (constellation == 0)
        return 2; // MOD_QPSK
    if (constellation == 4)
        return 3; // MOD_8PSK
    return 0;
}

// 8PSK deinterleaver row offsets by code rate: lib/xfecframe_demapper_cb_impl.cc:48-69
void rows_8psk(int rate, int rows, int& r0, int& r1, int& r2)
{
    if (rate == 4) { // C3_5
        r0 = rows * 2, r1 = rows, r2 = 0;
    } else if (rate == 26 || rate == 28 || rate == 38 || rate == 39 || rate == 19) {
        r0 = rows, r1 = 0, r2 = rows * 2; // C25_36 C13_18 C7_15 C8_15 C26_45
    } else {
        r0 = 0, r1 = rows, r2 = rows * 2;
    }
}

int snr_dev(dvbs2b200_code* h, int constellation, const float* d_iq, const int8_t* d_llr, int frames, float* d_snr,
            cudaStream_t stream)
{
    const BlobHeader& hd = h->hdr;
    const int bits = bits_per_symbol(constellation);
    if (!bits)
        return fail(DVBS2B200_EUNSUPPORTED, "Unsupported constellation");
    if (frames < 0)
        return fail(DVBS2B200_EINVAL, "negative frames");
    if (frames == 0)
        return DVBS2B200_OK;
    if (!d_iq || !d_snr)
        return fail(DVBS2B200_EINVAL, "null buffer");
    SnrLaunch p;
    memset(&p, 0, sizeof(p));
    p.iq = d_iq;
    p.llr = d_llr;
    p.snr_lin = d_snr;
    p.frames = frames;
    p.n_syms = hd.N / bits;
    p.constellation = constellation;
    if (constellation == 4)
        rows_8psk(hd.rate, p.n_syms, p.row0, p.row1, p.row2);
    cudaError_t e = snr_launch(p, stream);
    if (e != cudaSuccess)
        return cuda_fail(e, "snr_launch");
    h->launches += 1;
    return DVBS2B200_OK;
}

int demap_dev(dvbs2b200_code* h, int constellation, const float* d_iq, int frames, const float* d_n0, int8_t* d_llr,
              cudaStream_t stream)
{
    const BlobHeader& hd = h->hdr;
    const int bits = bits_per_symbol(constellation);
    if (!bits) // lib/xfecframe_demapper_cb_impl.cc:70-72
        return fail(DVBS2B200_EUNSUPPORTED, "Unsupported constellation");
    if (frames < 0)
        return fail(DVBS2B200_EINVAL, "negative frames");
    if (frames == 0)
        return DVBS2B200_OK;
    if (!d_iq || !d_n0 || !d_llr)
        return fail(DVBS2B200_EINVAL, "null buffer");
    if (((uintptr_t)d_iq & 15) || ((uintptr_t)d_llr & 3)) // 128-bit symbol loads, 32-bit packed LLR stores
        return fail(DVBS2B200_EINVAL, "symbol buffer must be 16-byte aligned and the LLR buffer 4-byte aligned");
    DemapLaunch p;
    memset(&p, 0, sizeof(p));
    p.n_syms = hd.N / bits;
    p.constellation = constellation;
    if (constellation == 4)
        rows_8psk(hd.rate, p.n_syms, p.row0, p.row1, p.row2);
    for (int f0 = 0; f0 < frames; f0 += 32768) { // gridDim.y limit
        p.frames = std::min(32768, frames - f0);
        p.iq = d_iq + (size_t)f0 * p.n_syms * 2;
        p.n0 = d_n0 + f0;
        p.llr = d_llr + (size_t)f0 * hd.N;
        cudaError_t e = demap_launch(p, stream);
        if (e != cudaSuccess)
            return cuda_fail(e, "demap_launch");
        h->launches += 1;
    }
    return DVBS2B200_OK;
}

// ---- BB layer ---------------------------------------------------------------------------------
// lib/bbdescrambler_bb_impl.cc:51-65: PRBS 1 + x^14 + x^15, register loaded with 100101010000000
void bb_prbs(std::vector<uint8_t>& seq, int nbytes)
{
    seq.assign((size_t)nbytes, 0);
    int sr = 0x4A80;
    for (int i = 0; i < nbytes * 8; i++) {
        const int b = ((sr) ^ (sr >> 1)) & 1;
        seq[i / 8] |= (uint8_t)(b << (7 - (i % 8)));
        sr >>= 1;
        if (b)
            sr |= 0x4000;
    }
}

int bb_ensure(dvbs2b200_code* h)
{
    if (h->bb_ready)
        return DVBS2B200_OK;
    const BlobHeader& hd = h->hdr;
    if (hd.kbch <= 80 || hd.kbch % 8)
        return fail(DVBS2B200_EUNSUPPORTED, "BBFRAME length must be a multiple of 8 bits");
    int rc;
    if ((rc = h->d_prbs.ensure((size_t)hd.kbch / 8)) || (rc = h->d_bbstate.ensure(sizeof(BbState))))
        return rc;
    std::vector<uint8_t> seq;
    bb_prbs(seq, hd.kbch / 8);
    // on the handle's stream: a plain cudaMemset runs on the legacy default stream, asynchronously with
    // respect to the host, and would race with the first kernels on the (non-blocking) handle stream
    CU(cudaMemcpyAsync(h->d_prbs.p, seq.data(), seq.size(), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemsetAsync(h->d_bbstate.p, 0, sizeof(BbState), h->stream));
    CU(cudaStreamSynchronize(h->stream)); // seq is a local
    h->bb_ready = true;
    return DVBS2B200_OK;
}

size_t bb_ts_capacity(const BlobHeader& hd, int frames)
{
    // every BBFRAME yields at most (187 carried + its DATAFIELD) / 188 packets
    return (size_t)frames * (((size_t)hd.kbch / 8 - 10 + 187) / 188) * 188;
}

int bb_descramble_dev(dvbs2b200_code* h, const uint8_t* d_in, int frames, uint8_t* d_out, cudaStream_t stream)
{
    if (frames < 0)
        return fail(DVBS2B200_EINVAL, "negative frames");
    if (frames == 0)
        return DVBS2B200_OK;
    if (!d_in || !d_out)
        return fail(DVBS2B200_EINVAL, "null buffer");
    int rc = bb_ensure(h);
    if (rc)
        return rc;
    cudaError_t e = bb_descramble_launch(d_in, d_out, (const uint8_t*)h->d_prbs.p, frames, h->hdr.kbch / 8, stream);
    if (e != cudaSuccess)
        return cuda_fail(e, "bb_descramble_launch");
    h->launches += 1;
    return DVBS2B200_OK;
}

int bb_deheader_dev(dvbs2b200_code* h, const uint8_t* d_bb, int frames, int scrambled, uint8_t* d_ts, size_t ts_cap,
                    cudaStream_t stream)
{
    if (frames < 0)
        return fail(DVBS2B200_EINVAL, "negative frames");
    int rc = bb_ensure(h);
    if (rc)
        return rc;
    if (frames == 0) {
        CU(cudaMemsetAsync(&((BbState*)h->d_bbstate.p)->produced, 0, sizeof(unsigned long long), stream));
        return DVBS2B200_OK;
    }
    if (!d_bb || !d_ts)
        return fail(DVBS2B200_EINVAL, "null buffer");
    if ((uintptr_t)d_ts & 3)
        return fail(DVBS2B200_EINVAL, "ts buffer must be 4-byte aligned");
    // a smaller buffer would lose packets while the deheader state advances past them: refuse before anything moves
    if (ts_cap < bb_ts_capacity(h->hdr, frames))
        return fail(DVBS2B200_EINVAL, "ts buffer smaller than dvbs2b200_bb_ts_capacity(frames)");
    if ((rc = h->d_bbrec.ensure((size_t)frames * sizeof(uint32_t))) || (rc = h->d_bbplan.ensure((size_t)frames * sizeof(BbPlan))))
        return rc;
    BbLaunch p;
    memset(&p, 0, sizeof(p));
    p.bb = d_bb;
    p.prbs = (const uint8_t*)h->d_prbs.p;
    p.scrambled = scrambled ? 1 : 0;
    p.frames = frames;
    p.kbytes = h->hdr.kbch / 8;
    p.rec = (uint32_t*)h->d_bbrec.p;
    p.plan = (BbPlan*)h->d_bbplan.p;
    p.state = (BbState*)h->d_bbstate.p;
    p.ts = d_ts;
    p.ts_cap = ts_cap;
    cudaError_t e = bb_deheader_launch(p, stream);
    if (e != cudaSuccess)
        return cuda_fail(e, "bb_deheader_launch");
    h->launches += 3;
    return DVBS2B200_OK;
}

// after the kernels of a host-pointer call: read the byte count, copy the packets out
int bb_fetch_ts(dvbs2b200_code* h, const uint8_t* d_ts, uint8_t* ts, size_t ts_cap, size_t* ts_bytes, cudaStream_t s)
{
    unsigned long long produced = 0;
    CU(cudaMemcpyAsync(&produced, &((BbState*)h->d_bbstate.p)->produced, sizeof(produced), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (produced > ts_cap)
        produced = ts_cap / 188 * 188;
    if (produced)
        CU(cudaMemcpyAsync(ts, d_ts, (size_t)produced, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (ts_bytes)
        *ts_bytes = (size_t)produced;
    return DVBS2B200_OK;
}

} // namespace

extern "C" {

int dvbs2b200_version(void) { return DVBS2B200_VERSION; }
const char* dvbs2b200_last_error(void) { return g_err.c_str(); }

int dvbs2b200_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver)
        return 0;
    if (e != cudaSuccess)
        return cuda_fail(e, "cudaGetDeviceCount");
    return n;
}

int dvbs2b200_host_register(void* ptr, size_t bytes)
{
    if (!ptr || bytes == 0)
        return fail(DVBS2B200_EINVAL, "null buffer");
    CU(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    return DVBS2B200_OK;
}

int dvbs2b200_host_unregister(void* ptr)
{
    if (!ptr)
        return fail(DVBS2B200_EINVAL, "null buffer");
    CU(cudaHostUnregister(ptr));
    return DVBS2B200_OK;
}

int dvbs2b200_num_tables(void) { return num_tables(); }
const char* dvbs2b200_table_name(int table)
{
    const LdpcTableDef* d = table_def(table);
    return d ? d->name : nullptr;
}

int dvbs2b200_lookup(int standard, int framesize, int rate, dvbs2b200_code_info* info)
{
    const ModcodDef* mc = find_modcod(standard, framesize, rate);
    if (!mc)
        return fail(DVBS2B200_EUNSUPPORTED, "no LDPC table for this (standard, framesize, rate)");
    if (info) {
        const LdpcTableDef* d = table_def(mc->table);
        Schedule s;
        build_schedule(*d, s);
        info->table = mc->table;
        info->n_ldpc = d->N;
        info->k_ldpc = d->K;
        info->q = d->q;
        info->n_circ = d->n_circ;
        info->links_total = d->links_total;
        info->max_cn_deg = s.max_cnt + 2;
        info->kbch = mc->kbch;
        info->nbch = mc->nbch;
        info->t = mc->t;
        info->gf_m = framesize == 1 ? 16 : framesize == 0 ? 14 : 15;
    }
    return DVBS2B200_OK;
}

int dvbs2b200_table_circulants(int table, uint32_t* out, int cap)
{
    const LdpcTableDef* d = table_def(table);
    if (!d)
        return fail(DVBS2B200_EINVAL, "bad table index");
    for (int i = 0; i < d->n_circ && i < cap; ++i)
        out[i] = d->circ[i];
    return d->n_circ;
}

int dvbs2b200_bch_genpoly(int framesize, int t, uint8_t* g, int cap)
{
    if (t < 1 || t > 12)
        return fail(DVBS2B200_EINVAL, "t out of range");
    std::vector<uint8_t> gp = bch_genpoly(bch_prim_poly(framesize), t);
    for (size_t i = 0; i < gp.size() && (int)i < cap; ++i)
        g[i] = gp[i];
    return (int)gp.size() - 1;
}

int dvbs2b200_schedule_stats(int table, int* steps_per_iter, int* max_depth, int* conflict_layers)
{
    const LdpcTableDef* d = table_def(table);
    if (!d)
        return fail(DVBS2B200_EINVAL, "bad table index");
    Schedule s;
    build_schedule(*d, s);
    if (steps_per_iter)
        *steps_per_iter = s.steps_per_iter;
    if (max_depth)
        *max_depth = s.max_depth;
    if (conflict_layers)
        *conflict_layers = s.conflict_layers;
    return DVBS2B200_OK;
}

int dvbs2b200_tables_build(int standard, int framesize, int rate, void* buf, size_t cap, size_t* size)
{
    std::vector<uint8_t> blob;
    std::string err;
    if (!build_blob(standard, framesize, rate, blob, err))
        return fail(DVBS2B200_EUNSUPPORTED, err);
    if (size)
        *size = blob.size();
    if (buf) {
        if (cap < blob.size())
            return fail(DVBS2B200_EINVAL, "buffer too small");
        memcpy(buf, blob.data(), blob.size());
    }
    return DVBS2B200_OK;
}

int dvbs2b200_code_create(dvbs2b200_code** h, int device, int standard, int framesize, int rate)
{
    if (!h)
        return fail(DVBS2B200_EINVAL, "null handle pointer");
    *h = nullptr;
    std::vector<uint8_t> blob;
    std::string err;
    if (!build_blob(standard, framesize, rate, blob, err))
        return fail(DVBS2B200_EUNSUPPORTED, err);
    return create_from_blob(h, device, std::move(blob));
}

int dvbs2b200_code_create_from_tables(dvbs2b200_code** h, int device, const void* blob, size_t size)
{
    if (!h || !blob)
        return fail(DVBS2B200_EINVAL, "null argument");
    *h = nullptr;
    std::vector<uint8_t> copy((const uint8_t*)blob, (const uint8_t*)blob + size);
    return create_from_blob(h, device, std::move(copy));
}

int dvbs2b200_code_export_tables(const dvbs2b200_code* h, void* buf, size_t cap, size_t* size)
{
    if (!h)
        return fail(DVBS2B200_EINVAL, "null handle");
    if (size)
        *size = h->blob.size();
    if (buf) {
        if (cap < h->blob.size())
            return fail(DVBS2B200_EINVAL, "buffer too small");
        memcpy(buf, h->blob.data(), h->blob.size());
    }
    return DVBS2B200_OK;
}

void dvbs2b200_code_destroy(dvbs2b200_code* h)
{
    if (!h)
        return;
    DeviceGuard g(h->device);
    if (h->ev_valid && h->ev_last)
        cudaEventSynchronize(h->ev_last); // asynchronous calls on a caller's stream may still use the scratch
    for (cudaStream_t st : { h->stream, h->s_in, h->s_out })
        if (st)
            cudaStreamSynchronize(st);
    for (DevBuf* b : { &h->d_in, &h->d_mid, &h->d_out, &h->d_post, &h->d_i32a, &h->d_i32b, &h->d_n0, &h->d_llr, &h->d_sync, &h->d_scratch, &h->d_flag, &h->d_prof, &h->d_next,
                       &h->d_prbs, &h->d_bbstate, &h->d_bbrec, &h->d_bbplan, &h->d_ts, &h->d_points, &h->d_err })
        b->release();
    for (HostBuf* b : { &h->h_ring, &h->h_out, &h->h_cnt })
        b->release();
    for (cudaEvent_t& e : h->ev_slot)
        if (e)
            cudaEventDestroy(e);
    if (h->ev_last)
        cudaEventDestroy(h->ev_last);
    if (h->d_blob)
        cudaFree(h->d_blob);
    for (cudaStream_t st : { h->stream, h->s_in, h->s_out })
        if (st)
            cudaStreamDestroy(st);
    delete h;
}

int dvbs2b200_code_info_get(const dvbs2b200_code* h, dvbs2b200_code_info* info)
{
    if (!h || !info)
        return fail(DVBS2B200_EINVAL, "null argument");
    info_from_header(h->hdr, info);
    return DVBS2B200_OK;
}

uint64_t dvbs2b200_launch_count(const dvbs2b200_code* h) { return h ? h->launches : 0; }

// ---- LDPC -------------------------------------------------------------------------------------
int dvbs2b200_ldpc_decode_dev(dvbs2b200_code* h, const int8_t* d_llr, int frames, int max_trials, int term_group,
                              int output_mode, uint8_t* d_hard, int8_t* d_llr_post, int32_t* d_trials_left,
                              void* stream)
{
    if (!h)
        return fail(DVBS2B200_EINVAL, "null handle");
    DeviceGuard g(h->device);
    StreamOrder so(h, (cudaStream_t)stream);
    return ldpc_dev(h, d_llr, frames, max_trials, term_group, output_mode, d_hard, d_llr_post, d_trials_left,
                    (cudaStream_t)stream);
}

int dvbs2b200_ldpc_decode(dvbs2b200_code* h, const int8_t* llr, int frames, int max_trials, int term_group,
                          int output_mode, uint8_t* hard, int8_t* llr_post, int32_t* trials_left)
{
    if (!h)
        return fail(DVBS2B200_EINVAL, "null handle");
    if (frames < 0)
        return fail(DVBS2B200_EINVAL, "negative frames");
    if (frames == 0)
        return DVBS2B200_OK;
    if (!llr)
        return fail(DVBS2B200_EINVAL, "llr is null");
    DeviceGuard g(h->device);
    StreamOrder so(h, h->stream);
    const BlobHeader& hd = h->hdr;
    const size_t in_bytes = (size_t)frames * hd.N;
    const size_t out_bytes = (size_t)frames * ldpc_out_bytes(hd, output_mode);
    int rc;
    if ((rc = h->d_in.ensure(in_bytes)))
        return rc;
    if (hard && (rc = h->d_out.ensure(out_bytes)))
        return rc;
    if (llr_post && (rc = h->d_post.ensure(in_bytes)))
        return rc;
    if (trials_left && (rc = h->d_i32a.ensure((size_t)frames * 4)))
        return rc;
    cudaStream_t s = h->stream;
    CU(cudaMemcpyAsync(h->d_in.p, llr, in_bytes, cudaMemcpyHostToDevice, s));
    rc = ldpc_dev(h, (const int8_t*)h->d_in.p, frames, max_trials, term_group, output_mode,
                  hard ? (uint8_t*)h->d_out.p : nullptr, llr_post ? (int8_t*)h->d_post.p : nullptr,
                  trials_left ? (int32_t*)h->d_i32a.p : nullptr, s);
    if (rc)
        return rc;
    if (hard)
        CU(cudaMemcpyAsync(hard, h->d_out.p, out_bytes, cudaMemcpyDeviceToHost, s));
    if (llr_post)
        CU(cudaMemcpyAsync(llr_post, h->d_post.p, in_bytes, cudaMemcpyDeviceToHost, s));
    if (trials_left)
        CU(cudaMemcpyAsync(trials_left, h->d_i32a.p, (size_t)frames * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return DVBS2B200_OK;
}

// ---- BCH --------------------------------------------------------------------------------------
int dvbs2b200_bch_decode_dev(dvbs2b200_code* h, const uint8_t* d_cw, int frames, uint8_t* d_msg,
                             int32_t* d_corrections, void* stream)
{
    if (!h)
        return fail(DVBS2B200_EINVAL, "null handle");
    DeviceGuard g(h->device);
    StreamOrder so(h, (cudaStream_t)stream);
    return bch_dev(h, d_cw, h->hdr.nbch / 8, frames, d_msg, d_corrections, (cudaStream_t)stream);
}

int dvbs2b200_bch_decode(dvbs2b200_code* h, const uint8_t* cw, int frames, uint8_t* msg, int32_t* corrections)
{
    if (!h)
        return fail(DVBS2B200_EINVAL, "null handle");
    if (frames < 0)
        return fail(DVBS2B200_EINVAL, "negative frames");
    if (frames == 0)
        return DVBS2B200_OK;
    if (!cw || !msg)
        return fail(DVBS2B200_EINVAL, "null buffer");
    DeviceGuard g(h->device);
    StreamOrder so(h, h->stream);
    const BlobHeader& hd = h->hdr;
    const size_t in_bytes = (size_t)frames * (hd.nbch / 8), out_bytes = (size_t)frames * (hd.kbch / 8);
    int rc;
    if ((rc = h->d_mid.ensure(in_bytes)) || (rc = h->d_out.ensure(out_bytes)) ||
        (rc = h->d_i32b.ensure((size_t)frames * 4)))
        return rc;
    cudaStream_t s = h->stream;
    CU(cudaMemcpyAsync(h->d_mid.p, cw, in_bytes, cudaMemcpyHostToDevice, s));
    rc = bch_dev(h, (const uint8_t*)h->d_mid.p, hd.nbch / 8, frames, (uint8_t*)h->d_out.p, (int32_t*)h->d_i32b.p, s);
    if (rc)
        return rc;
    CU(cudaMemcpyAsync(msg, h->d_out.p, out_bytes, cudaMemcpyDeviceToHost, s));
    if (corrections)
        CU(cudaMemcpyAsync(corrections, h->d_i32b.p, (size_t)frames * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return DVBS2B200_OK;
}

// ---- demapper ---------------------------------------------------------------------------------
int dvbs2b200_demap_dev(dvbs2b200_code* h, int constellation, const float* d_iq, int frames, const float* d_n0,
                        int8_t* d_llr, void* stream)
{
    if (!h)
        return fail(DVBS2B200_EINVAL, "null handle");
    DeviceGuard g(h->device);
    StreamOrder so(h, (cudaStream_t)stream);
    return demap_dev(h, constellation, d_iq, frames, d_n0, d_llr, (cudaStream_t)stream);
}

int dvbs2b200_demap(dvbs2b200_code* h, int constellation, const float* iq, int frames, const float* n0, int8_t* llr)
{
    if (!h)
        return fail(DVBS2B200_EINVAL, "null handle");
    const int bits = bits_per_symbol(constellation);
    if (!bits)
        return fail(DVBS2B200_EUNSUPPORTED, "Unsupported constellation");
    if (frames < 0)
        return fail(DVBS2B200_EINVAL, "negative frames");
    if (frames == 0)
        return DVBS2B200_OK;
    if (!iq || !n0 || !llr)
        return fail(DVBS2B200_EINVAL, "null buffer");
    DeviceGuard g(h->device);
    StreamOrder so(h, h->stream);
    const BlobHeader& hd = h->hdr;
    const size_t iq_bytes = (size_t)frames * (hd.N / bits) * 8, llr_bytes = (size_t)frames * hd.N;
    int rc;
    if ((rc = h->d_in.ensure(iq_bytes)) || (rc = h->d_llr.ensure(llr_bytes)) || (rc = h->d_n0.ensure((size_t)frames * 4)))
        return rc;
    cudaStream_t s = h->stream;
    CU(cudaMemcpyAsync(h->d_in.p, iq, iq_bytes, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(h->d_n0.p, n0, (size_t)frames * 4, cudaMemcpyHostToDevice, s));
    rc = demap_dev(h, constellation, (const float*)h->d_in.p, frames, (const float*)h->d_n0.p, (int8_t*)h->d_llr.p, s);
    if (rc)
        return rc;
    CU(cudaMemcpyAsync(llr, h->d_llr.p, llr_bytes, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return DVBS2B200_OK;
}

// ---- fused chain ------------------------------------------------------------------------------
} // extern "C"

namespace {
// demap -> LDPC -> BCH on frames [f0, f0 + nf) of a batch whose buffers hold `frames` frames
int fec_dev_range(dvbs2b200_code* h, int constellation, const float* d_iq, const float* d_n0, const int8_t* d_llr,
                  int f0, int nf, int max_trials, int term_group, uint8_t* d_llr_scratch, uint8_t* d_mid,
                  uint8_t* d_msg, int32_t* d_trials_left, int32_t* d_corrections, cudaStream_t s,
                  const unsigned int* d_ready = nullptr, int ready_chunk = 0, unsigned int* d_err = nullptr)
{
    const BlobHeader& hd = h->hdr;
    int rc;
    const int8_t* llr = nullptr;
    if (d_iq) {
        const int bits = bits_per_symbol(constellation);
        if (!bits)
            return fail(DVBS2B200_EUNSUPPORTED, "Unsupported constellation");
        int8_t* out = (int8_t*)d_llr_scratch + (size_t)f0 * hd.N;
        if ((rc = demap_dev(h, constellation, d_iq + (size_t)f0 * (hd.N / bits) * 2, nf, d_n0 + f0, out, s)))
            return rc;
        llr = out;
    } else {
        llr = d_llr + (size_t)f0 * hd.N;
    }
    const int mid_stride = hd.kldpc_out / 8; // OM_MESSAGE: BCH codeword bytes
    uint8_t* mid = d_mid + (size_t)f0 * mid_stride;
    if ((rc = ldpc_dev(h, llr, nf, max_trials, term_group, /*OM_MESSAGE*/ 1, mid, nullptr,
                       d_trials_left ? d_trials_left + f0 : nullptr, s, d_ready, ready_chunk, d_err)))
        return rc;
    return bch_dev(h, mid, mid_stride, nf, d_msg + (size_t)f0 * (hd.kbch / 8),
                   d_corrections ? d_corrections + f0 : nullptr, s);
}
} // namespace

extern "C" {

int dvbs2b200_fec_decode_dev(dvbs2b200_code* h, int constellation, const float* d_iq, const float* d_n0,
                             const int8_t* d_llr, int frames, int max_trials, int term_group, uint8_t* d_msg,
                             int32_t* d_trials_left, int32_t* d_corrections, void* stream)
{
    if (!h)
        return fail(DVBS2B200_EINVAL, "null handle");
    if (frames < 0)
        return fail(DVBS2B200_EINVAL, "negative frames");
    if (frames == 0)
        return DVBS2B200_OK;
    if (!d_msg || (!d_iq && !d_llr))
        return fail(DVBS2B200_EINVAL, "null buffer");
    if (d_iq && !d_n0)
        return fail(DVBS2B200_EINVAL, "n0 is null");
    DeviceGuard g(h->device);
    StreamOrder so(h, (cudaStream_t)stream);
    const BlobHeader& hd = h->hdr;
    int rc;
    if (d_iq && (rc = h->d_llr.ensure((size_t)frames * hd.N)))
        return rc;
    if ((rc = h->d_mid.ensure((size_t)frames * (hd.kldpc_out / 8))))
        return rc;
    return fec_dev_range(h, constellation, d_iq, d_n0, d_llr, 0, frames, max_trials, term_group, (uint8_t*)h->d_llr.p,
                         (uint8_t*)h->d_mid.p, d_msg, d_trials_left, d_corrections, (cudaStream_t)stream);
}

int dvbs2b200_fec_decode(dvbs2b200_code* h, int constellation, const float* iq, const float* n0, const int8_t* llr,
                         int frames, int max_trials, int term_group, uint8_t* msg, int32_t* trials_left,
                         int32_t* corrections)
{
    if (!h)
        return fail(DVBS2B200_EINVAL, "null handle");
    if (frames < 0)
        return fail(DVBS2B200_EINVAL, "negative frames");
    if (frames == 0)
        return DVBS2B200_OK;
    if (!msg || (!iq && !llr))
        return fail(DVBS2B200_EINVAL, "null buffer");
    if (iq && !n0)
        return fail(DVBS2B200_EINVAL, "n0 is null");
    DeviceGuard g(h->device);
    StreamOrder so(h, h->stream);
    const BlobHeader& hd = h->hdr;
    const int bits = iq ? bits_per_symbol(constellation) : 0;
    if (iq && !bits)
        return fail(DVBS2B200_EUNSUPPORTED, "Unsupported constellation");
    if (term_group != 0 && term_group != 1 && term_group != 16 && term_group != 32)
        return fail(DVBS2B200_EINVAL, "term_group must be 0, 16 or 32");
    if (term_group > 1 && frames % term_group)
        return fail(DVBS2B200_EINVAL, "frames must be a multiple of term_group");
    const size_t in_stride = iq ? (size_t)(hd.N / bits) * 8 : (size_t)hd.N; // bytes per frame of input
    const size_t out_stride = hd.kbch / 8;
    int rc;
    if ((rc = h->d_in.ensure((size_t)frames * in_stride)) || (rc = h->d_out.ensure((size_t)frames * out_stride)) ||
        (rc = h->d_i32a.ensure((size_t)frames * 4)) || (rc = h->d_i32b.ensure((size_t)frames * 4)) ||
        (rc = h->d_mid.ensure((size_t)frames * (hd.kldpc_out / 8))))
        return rc;
    if (iq && ((rc = h->d_n0.ensure((size_t)frames * 4)) || (rc = h->d_llr.ensure((size_t)frames * hd.N))))
        return rc;
    if (!iq && term_group <= 1) {
        // Streaming path (LLR input): ONE persistent LDPC launch over the whole batch, issued FIRST; the input
        // then arrives in chunks on the copy-in stream, each followed by a 4-byte DMA copy that publishes the
        // number of chunks that have landed; a CTA waits on that counter before it loads a frame.  The
        // counter is advanced by the copy engine, not by a kernel, so the resident, spinning CTAs cannot starve
        // it of SM resources; a CTA that waits for more than ~17 s sets the error word and the call fails.
        // Pageable host buffers (GNU Radio's ring buffers) are staged chunk by chunk through a ring of pinned
        // slots while the device decodes the chunks that are already there.
        const int sc = std::max(32, h->sm_count); // frames per arrival flag
        const int n_sc = (frames + sc - 1) / sc;
        const bool in_pinned = host_ptr_is_pinned(llr), out_pinned = host_ptr_is_pinned(msg);
        constexpr int kSlots = 3;
        if ((rc = h->d_flag.ensure(16)) || (rc = h->d_err.ensure(16)) || (rc = h->h_cnt.ensure((size_t)n_sc * sizeof(unsigned int))))
            return rc;
        if (!in_pinned && (rc = h->h_ring.ensure((size_t)kSlots * sc * in_stride)))
            return rc;
        const size_t out_bytes_al = ((size_t)frames * out_stride + 15) & ~(size_t)15;
        if (!out_pinned && (rc = h->h_out.ensure(out_bytes_al + (size_t)frames * 8)))
            return rc;
        for (int k = 0; k < kSlots; ++k)
            if (!h->ev_slot[k])
                CU(cudaEventCreateWithFlags(&h->ev_slot[k], cudaEventDisableTiming));
        unsigned int* cnt = (unsigned int*)h->h_cnt.p;
        for (int c = 0; c < n_sc; ++c)
            cnt[c] = (unsigned int)(c + 1);
        cudaEvent_t ev_zero = nullptr;
        cudaError_t e;
        unsigned int* flag = (unsigned int*)h->d_flag.p;
        auto bail = [&](cudaError_t err, const char* what) {
            // let the resident kernel out: everything has "arrived" (its results are discarded with the error)
            unsigned int all = 0xffffffffu;
            cudaMemcpyAsync(flag, &all, 4, cudaMemcpyHostToDevice, h->s_in);
            cudaDeviceSynchronize();
            if (ev_zero)
                cudaEventDestroy(ev_zero);
            return cuda_fail(err, what);
        };
        if ((e = cudaMemsetAsync(flag, 0, 4, h->s_in)) != cudaSuccess)
            return cuda_fail(e, "cudaMemsetAsync(flag)");
        if ((e = cudaMemsetAsync(h->d_err.p, 0, 4, h->s_in)) != cudaSuccess)
            return cuda_fail(e, "cudaMemsetAsync(err)");
        if ((e = cudaEventCreateWithFlags(&ev_zero, cudaEventDisableTiming)) != cudaSuccess)
            return cuda_fail(e, "cudaEventCreate");
        if ((e = cudaEventRecord(ev_zero, h->s_in)) != cudaSuccess || (e = cudaStreamWaitEvent(h->stream, ev_zero, 0)) != cudaSuccess) {
            cudaEventDestroy(ev_zero);
            return cuda_fail(e, "cudaEventRecord");
        }
        rc = fec_dev_range(h, constellation, nullptr, nullptr, (const int8_t*)h->d_in.p, 0, frames, max_trials, term_group,
                           (uint8_t*)h->d_llr.p, (uint8_t*)h->d_mid.p, (uint8_t*)h->d_out.p, (int32_t*)h->d_i32a.p,
                           (int32_t*)h->d_i32b.p, h->stream, flag, sc, (unsigned int*)h->d_err.p);
        if (rc) {
            bail(cudaSuccess, "launch");
            return rc;
        }
        for (int c = 0; c < n_sc; ++c) {
            const int f0 = c * sc, nf = std::min(sc, frames - f0);
            const uint8_t* src = (const uint8_t*)llr + (size_t)f0 * in_stride;
            if (!in_pinned) {
                const int slot = c % kSlots;
                uint8_t* stage = (uint8_t*)h->h_ring.p + (size_t)slot * sc * in_stride;
                if (c >= kSlots && (e = cudaEventSynchronize(h->ev_slot[slot])) != cudaSuccess)
                    return bail(e, "cudaEventSynchronize(slot)");
                memcpy(stage, src, (size_t)nf * in_stride);
                src = stage;
            }
            if ((e = cudaMemcpyAsync((uint8_t*)h->d_in.p + (size_t)f0 * in_stride, src, (size_t)nf * in_stride, cudaMemcpyHostToDevice,
                                     h->s_in)) != cudaSuccess)
                return bail(e, "cudaMemcpyAsync(llr chunk)");
            if (!in_pinned && (e = cudaEventRecord(h->ev_slot[c % kSlots], h->s_in)) != cudaSuccess)
                return bail(e, "cudaEventRecord(slot)");
            if ((e = cudaMemcpyAsync(flag, &cnt[c], 4, cudaMemcpyHostToDevice, h->s_in)) != cudaSuccess)
                return bail(e, "cudaMemcpyAsync(arrival counter)");
        }
        uint8_t* o_msg = out_pinned ? msg : (uint8_t*)h->h_out.p;
        int32_t* o_tr = out_pinned ? trials_left : (int32_t*)((uint8_t*)h->h_out.p + out_bytes_al);
        int32_t* o_co = out_pinned ? corrections : o_tr + frames;
        if ((e = cudaMemcpyAsync(o_msg, h->d_out.p, (size_t)frames * out_stride, cudaMemcpyDeviceToHost, h->stream)) != cudaSuccess)
            return bail(e, "cudaMemcpyAsync(msg)");
        if (trials_left && (e = cudaMemcpyAsync(o_tr, h->d_i32a.p, (size_t)frames * 4, cudaMemcpyDeviceToHost, h->stream)) != cudaSuccess)
            return bail(e, "cudaMemcpyAsync(trials)");
        if (corrections && (e = cudaMemcpyAsync(o_co, h->d_i32b.p, (size_t)frames * 4, cudaMemcpyDeviceToHost, h->stream)) != cudaSuccess)
            return bail(e, "cudaMemcpyAsync(corrections)");
        unsigned int err_word = 0;
        if ((e = cudaMemcpyAsync(&cnt[0], h->d_err.p, 4, cudaMemcpyDeviceToHost, h->stream)) != cudaSuccess) // cnt[0] is free by now
            return bail(e, "cudaMemcpyAsync(err)");
        if ((e = cudaStreamSynchronize(h->s_in)) != cudaSuccess)
            return bail(e, "cudaStreamSynchronize");
        if ((e = cudaStreamSynchronize(h->stream)) != cudaSuccess)
            return bail(e, "cudaStreamSynchronize");
        err_word = cnt[0];
        cudaEventDestroy(ev_zero);
        if (err_word)
            return fail(DVBS2B200_ECUDA, "LDPC kernel gave up waiting for its input (host-to-device copy never arrived)");
        if (!out_pinned) {
            memcpy(msg, o_msg, (size_t)frames * out_stride);
            if (trials_left)
                memcpy(trials_left, o_tr, (size_t)frames * 4);
            if (corrections)
                memcpy(corrections, o_co, (size_t)frames * 4);
        }
        return DVBS2B200_OK;
    }
    // Pipeline (symbol input / group mode): the batch is cut into chunks of full waves of resident CTAs; chunk c+1 is copied in
    int chunk = std::max(1, h->sm_count * std::max(1, h->ldpc_ctas)) * 2;
    if (term_group > 1)
        chunk = std::max(term_group, chunk / term_group * term_group);
    const int n_chunks = (frames + chunk - 1) / chunk;
    std::vector<cudaEvent_t> ev_in(n_chunks, nullptr), ev_done(n_chunks, nullptr);
    auto cleanup = [&]() {
        for (cudaEvent_t e : ev_in)
            if (e)
                cudaEventDestroy(e);
        for (cudaEvent_t e : ev_done)
            if (e)
                cudaEventDestroy(e);
    };
#define CUP(call)                                \
    do {                                         \
        cudaError_t e__ = (call);                \
        if (e__ != cudaSuccess) {                \
            cudaDeviceSynchronize();             \
            cleanup();                           \
            return cuda_fail(e__, #call);        \
        }                                        \
    } while (0)
    if (iq)
        CUP(cudaMemcpyAsync(h->d_n0.p, n0, (size_t)frames * 4, cudaMemcpyHostToDevice, h->s_in));
    const uint8_t* src = iq ? (const uint8_t*)iq : (const uint8_t*)llr;
    for (int c = 0; c < n_chunks; ++c) {
        const int f0 = c * chunk, nf = std::min(chunk, frames - f0);
        CUP(cudaEventCreateWithFlags(&ev_in[c], cudaEventDisableTiming));
        CUP(cudaEventCreateWithFlags(&ev_done[c], cudaEventDisableTiming));
        CUP(cudaMemcpyAsync((uint8_t*)h->d_in.p + (size_t)f0 * in_stride, src + (size_t)f0 * in_stride,
                            (size_t)nf * in_stride, cudaMemcpyHostToDevice, h->s_in));
        CUP(cudaEventRecord(ev_in[c], h->s_in));
    }
    for (int c = 0; c < n_chunks; ++c) {
        const int f0 = c * chunk, nf = std::min(chunk, frames - f0);
        CUP(cudaStreamWaitEvent(h->stream, ev_in[c], 0));
        rc = fec_dev_range(h, constellation, iq ? (const float*)h->d_in.p : nullptr, (const float*)h->d_n0.p,
                           iq ? nullptr : (const int8_t*)h->d_in.p, f0, nf, max_trials, term_group, (uint8_t*)h->d_llr.p,
                           (uint8_t*)h->d_mid.p, (uint8_t*)h->d_out.p, (int32_t*)h->d_i32a.p, (int32_t*)h->d_i32b.p,
                           h->stream);
        if (rc) {
            cudaDeviceSynchronize();
            cleanup();
            return rc;
        }
        CUP(cudaEventRecord(ev_done[c], h->stream));
        CUP(cudaStreamWaitEvent(h->s_out, ev_done[c], 0));
        CUP(cudaMemcpyAsync(msg + (size_t)f0 * out_stride, (uint8_t*)h->d_out.p + (size_t)f0 * out_stride,
                            (size_t)nf * out_stride, cudaMemcpyDeviceToHost, h->s_out));
        if (trials_left)
            CUP(cudaMemcpyAsync(trials_left + f0, (int32_t*)h->d_i32a.p + f0, (size_t)nf * 4, cudaMemcpyDeviceToHost, h->s_out));
        if (corrections)
            CUP(cudaMemcpyAsync(corrections + f0, (int32_t*)h->d_i32b.p + f0, (size_t)nf * 4, cudaMemcpyDeviceToHost, h->s_out));
    }
    CUP(cudaStreamSynchronize(h->s_out));
    CUP(cudaStreamSynchronize(h->stream));
#undef CUP
    cleanup();
    return DVBS2B200_OK;
}

// ---- mixed-MODCOD batches -------------------------------------------------------------------------------
int dvbs2b200_mixed_create(dvbs2b200_mixed** out, int device, int n_codes, const int* standard, const int* framesize,
                           const int* rate)
{
    if (!out || n_codes <= 0 || n_codes > 255 || !standard || !framesize || !rate)
        return fail(DVBS2B200_EINVAL, "bad argument");
    *out = nullptr;
    dvbs2b200_mixed* m = new (std::nothrow) dvbs2b200_mixed();
    if (!m)
        return fail(DVBS2B200_ENOMEM, "out of host memory");
    m->device = device;
    for (int c = 0; c < n_codes; ++c) {
        dvbs2b200_code* h = nullptr;
        int rc = dvbs2b200_code_create(&h, device, standard[c], framesize[c], rate[c]);
        if (rc) {
            dvbs2b200_mixed_destroy(m);
            return rc;
        }
        m->codes.push_back(h);
    }
    m->per.resize(n_codes);
    DeviceGuard g(device);
    cudaError_t e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        dvbs2b200_mixed_destroy(m);
        return cuda_fail(e, "cudaStreamCreate");
    }
    *out = m;
    return DVBS2B200_OK;
}

int dvbs2b200_mixed_create_from_tables(dvbs2b200_mixed** out, int device, int n_codes, const void* const* blobs, const size_t* sizes)
{
    if (!out || n_codes <= 0 || n_codes > 255 || !blobs || !sizes)
        return fail(DVBS2B200_EINVAL, "bad argument");
    *out = nullptr;
    dvbs2b200_mixed* m = new (std::nothrow) dvbs2b200_mixed();
    if (!m)
        return fail(DVBS2B200_ENOMEM, "out of host memory");
    m->device = device;
    for (int c = 0; c < n_codes; ++c) {
        dvbs2b200_code* h = nullptr;
        int rc = dvbs2b200_code_create_from_tables(&h, device, blobs[c], sizes[c]);
        if (rc) {
            dvbs2b200_mixed_destroy(m);
            return rc;
        }
        m->codes.push_back(h);
    }
    m->per.resize(n_codes);
    DeviceGuard g(device);
    cudaError_t e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        dvbs2b200_mixed_destroy(m);
        return cuda_fail(e, "cudaStreamCreate");
    }
    *out = m;
    return DVBS2B200_OK;
}

void dvbs2b200_mixed_destroy(dvbs2b200_mixed* m)
{
    if (!m)
        return;
    {
        DeviceGuard g(m->device);
        if (m->stream) {
            cudaStreamSynchronize(m->stream);
            cudaStreamDestroy(m->stream);
        }
        for (DevBuf* b : { &m->d_in, &m->d_out, &m->d_tr, &m->d_co })
            b->release();
        for (auto& pc : m->per)
            for (DevBuf* b : { &pc.d_in_off, &pc.d_out_off, &pc.d_pos, &pc.d_stage_in, &pc.d_stage_out, &pc.d_tr, &pc.d_co })
                b->release();
    }
    for (dvbs2b200_code* h : m->codes)
        dvbs2b200_code_destroy(h);
    delete m;
}

int dvbs2b200_mixed_code_info(const dvbs2b200_mixed* m, int code, dvbs2b200_code_info* info)
{
    if (!m || code < 0 || code >= (int)m->codes.size())
        return fail(DVBS2B200_EINVAL, "bad argument");
    return dvbs2b200_code_info_get(m->codes[code], info);
}

} // extern "C"

namespace {
// The batch with its input and outputs in device memory: bucket the frames by code (host), gather each bucket,
// decode it on its code's stream, scatter the results back; `stream` waits for every bucket before it goes on.
int mixed_dev(dvbs2b200_mixed* m, int frames, const uint8_t* code_id, const int8_t* d_llr, int max_trials, uint8_t* d_msg,
              int32_t* d_trials_left, int32_t* d_corrections, cudaStream_t stream, unsigned long long in_total_expect)
{
    const int nc = (int)m->codes.size();
    std::vector<std::vector<unsigned long long>> in_off(nc), out_off(nc);
    std::vector<std::vector<int32_t>> pos(nc);
    unsigned long long in_total = 0, out_total = 0;
    for (int f = 0; f < frames; ++f) {
        const int c = code_id[f];
        const BlobHeader& hd = m->codes[c]->hdr;
        in_off[c].push_back(in_total);
        out_off[c].push_back(out_total);
        pos[c].push_back(f);
        in_total += (unsigned long long)hd.N;
        out_total += (unsigned long long)hd.kbch / 8;
    }
    (void)in_total_expect;
    int rc;
    // status words are always produced on the device (the scatter kernel writes them); the caller may not want them
    if (!d_trials_left) {
        if ((rc = m->d_tr.ensure((size_t)frames * 4)))
            return rc;
        d_trials_left = (int32_t*)m->d_tr.p;
    }
    if (!d_corrections) {
        if ((rc = m->d_co.ensure((size_t)frames * 4)))
            return rc;
        d_corrections = (int32_t*)m->d_co.p;
    }
    cudaEvent_t ev_in = nullptr;
    std::vector<cudaEvent_t> ev_done(nc, nullptr);
    auto cleanup = [&]() {
        if (ev_in)
            cudaEventDestroy(ev_in);
        for (cudaEvent_t e : ev_done)
            if (e)
                cudaEventDestroy(e);
    };
#define CUM(call)                                \
    do {                                         \
        cudaError_t e__ = (call);                \
        if (e__ != cudaSuccess) {                \
            cudaDeviceSynchronize();             \
            cleanup();                           \
            return cuda_fail(e__, #call);        \
        }                                        \
    } while (0)
    CUM(cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming));
    CUM(cudaEventRecord(ev_in, stream));
    for (int c = 0; c < nc; ++c) {
        const int n = (int)pos[c].size();
        if (n == 0)
            continue;
        dvbs2b200_code* h = m->codes[c];
        const BlobHeader& hd = h->hdr;
        auto& pc = m->per[c];
        const int kb = hd.kbch / 8;
        if ((rc = pc.d_in_off.ensure((size_t)n * 8)) || (rc = pc.d_out_off.ensure((size_t)n * 8)) || (rc = pc.d_pos.ensure((size_t)n * 4)) ||
            (rc = pc.d_stage_in.ensure((size_t)n * hd.N)) || (rc = pc.d_stage_out.ensure((size_t)n * kb)) ||
            (rc = pc.d_tr.ensure((size_t)n * 4)) || (rc = pc.d_co.ensure((size_t)n * 4))) {
            cudaDeviceSynchronize();
            cleanup();
            return rc;
        }
        cudaStream_t s = h->stream; // every code decodes on its own stream: the codes of a batch overlap
        // (pageable sources: the copies are staged before the calls return, the vectors may go out of scope)
        CUM(cudaMemcpyAsync(pc.d_in_off.p, in_off[c].data(), (size_t)n * 8, cudaMemcpyHostToDevice, s));
        CUM(cudaMemcpyAsync(pc.d_out_off.p, out_off[c].data(), (size_t)n * 8, cudaMemcpyHostToDevice, s));
        CUM(cudaMemcpyAsync(pc.d_pos.p, pos[c].data(), (size_t)n * 4, cudaMemcpyHostToDevice, s));
        CUM(cudaStreamWaitEvent(s, ev_in, 0));
        CUM(gather_launch((const uint8_t*)d_llr, (const unsigned long long*)pc.d_in_off.p, (uint8_t*)pc.d_stage_in.p, hd.N, n, s));
        rc = dvbs2b200_fec_decode_dev(h, 0, nullptr, nullptr, (const int8_t*)pc.d_stage_in.p, n, max_trials, 0, (uint8_t*)pc.d_stage_out.p,
                                      (int32_t*)pc.d_tr.p, (int32_t*)pc.d_co.p, s);
        if (rc) {
            cudaDeviceSynchronize();
            cleanup();
            return rc;
        }
        CUM(scatter_launch((const uint8_t*)pc.d_stage_out.p, (const unsigned long long*)pc.d_out_off.p, d_msg, kb, n,
                           (const int32_t*)pc.d_tr.p, (const int32_t*)pc.d_co.p, (const int32_t*)pc.d_pos.p, d_trials_left,
                           d_corrections, s));
        h->launches += 2;
        CUM(cudaEventCreateWithFlags(&ev_done[c], cudaEventDisableTiming));
        CUM(cudaEventRecord(ev_done[c], s));
        CUM(cudaStreamWaitEvent(stream, ev_done[c], 0));
    }
#undef CUM
    cleanup();
    return DVBS2B200_OK;
}

int mixed_sizes(const dvbs2b200_mixed* m, int frames, const uint8_t* code_id, unsigned long long* in_total, unsigned long long* out_total)
{
    const int nc = (int)m->codes.size();
    *in_total = *out_total = 0;
    for (int f = 0; f < frames; ++f) {
        if (code_id[f] >= nc)
            return fail(DVBS2B200_EINVAL, "code_id out of range");
        const BlobHeader& hd = m->codes[code_id[f]]->hdr;
        *in_total += (unsigned long long)hd.N;
        *out_total += (unsigned long long)hd.kbch / 8;
    }
    return DVBS2B200_OK;
}
} // namespace

extern "C" {

int dvbs2b200_mixed_fec_decode_dev(dvbs2b200_mixed* m, int frames, const uint8_t* code_id, const int8_t* d_llr, int max_trials,
                                   uint8_t* d_msg, int32_t* d_trials_left, int32_t* d_corrections, void* stream)
{
    if (!m)
        return fail(DVBS2B200_EINVAL, "null handle");
    if (frames < 0)
        return fail(DVBS2B200_EINVAL, "negative frames");
    if (frames == 0)
        return DVBS2B200_OK;
    if (!code_id || !d_llr || !d_msg)
        return fail(DVBS2B200_EINVAL, "null buffer");
    DeviceGuard g(m->device);
    unsigned long long in_total, out_total;
    int rc = mixed_sizes(m, frames, code_id, &in_total, &out_total);
    if (rc)
        return rc;
    return mixed_dev(m, frames, code_id, d_llr, max_trials, d_msg, d_trials_left, d_corrections, (cudaStream_t)stream, in_total);
}

int dvbs2b200_mixed_fec_decode(dvbs2b200_mixed* m, int frames, const uint8_t* code_id, const int8_t* llr, int max_trials,
                               uint8_t* msg, int32_t* trials_left, int32_t* corrections)
{
    if (!m)
        return fail(DVBS2B200_EINVAL, "null handle");
    if (frames < 0)
        return fail(DVBS2B200_EINVAL, "negative frames");
    if (frames == 0)
        return DVBS2B200_OK;
    if (!code_id || !llr || !msg)
        return fail(DVBS2B200_EINVAL, "null buffer");
    DeviceGuard g(m->device);
    unsigned long long in_total, out_total;
    int rc = mixed_sizes(m, frames, code_id, &in_total, &out_total);
    if (rc)
        return rc;
    if ((rc = m->d_in.ensure(in_total)) || (rc = m->d_out.ensure(out_total)) || (rc = m->d_tr.ensure((size_t)frames * 4)) ||
        (rc = m->d_co.ensure((size_t)frames * 4)))
        return rc;
    auto bail = [&](cudaError_t e, const char* what) {
        cudaDeviceSynchronize();
        return cuda_fail(e, what);
    };
    cudaError_t e;
    if ((e = cudaMemcpyAsync(m->d_in.p, llr, in_total, cudaMemcpyHostToDevice, m->stream)) != cudaSuccess)
        return bail(e, "cudaMemcpyAsync(llr)");
    if ((rc = mixed_dev(m, frames, code_id, (const int8_t*)m->d_in.p, max_trials, (uint8_t*)m->d_out.p, (int32_t*)m->d_tr.p,
                        (int32_t*)m->d_co.p, m->stream, in_total)))
        return rc;
    if ((e = cudaMemcpyAsync(msg, m->d_out.p, out_total, cudaMemcpyDeviceToHost, m->stream)) != cudaSuccess)
        return bail(e, "cudaMemcpyAsync(msg)");
    if (trials_left && (e = cudaMemcpyAsync(trials_left, m->d_tr.p, (size_t)frames * 4, cudaMemcpyDeviceToHost, m->stream)) != cudaSuccess)
        return bail(e, "cudaMemcpyAsync(trials)");
    if (corrections && (e = cudaMemcpyAsync(corrections, m->d_co.p, (size_t)frames * 4, cudaMemcpyDeviceToHost, m->stream)) != cudaSuccess)
        return bail(e, "cudaMemcpyAsync(corrections)");
    if ((e = cudaStreamSynchronize(m->stream)) != cudaSuccess)
        return bail(e, "cudaStreamSynchronize");
    return DVBS2B200_OK;
}

uint64_t dvbs2b200_mixed_launch_count(const dvbs2b200_mixed* m)
{
    uint64_t n = 0;
    if (m)
        for (const dvbs2b200_code* h : m->codes)
            n += h->launches;
    return n;
}

// ---- table-driven demapper (16APSK / 32APSK / any constellation of up to 32 points) ------------------
int dvbs2b200_demap_table_dev(dvbs2b200_code* h, int bits, const float* points, const int* row_offsets, const float* d_iq,
                              int frames, const float* d_n0, int8_t* d_llr, void* stream)
{
    if (!h)
        return fail(DVBS2B200_EINVAL, "null handle");
    if (bits < 1 || bits > 5 || !points || !row_offsets)
        return fail(DVBS2B200_EINVAL, "bits must be 1..5 with a constellation and row offsets");
    const BlobHeader& hd = h->hdr;
    if (hd.N % bits)
        return fail(DVBS2B200_EINVAL, "frame length is not a multiple of the bits per symbol");
    if (frames < 0)
        return fail(DVBS2B200_EINVAL, "negative frames");
    if (frames == 0)
        return DVBS2B200_OK;
    if (!d_iq || !d_n0 || !d_llr)
        return fail(DVBS2B200_EINVAL, "null buffer");
    const int n_syms = hd.N / bits;
    for (int k = 0; k < bits; ++k)
        if (row_offsets[k] < 0 || row_offsets[k] + n_syms > hd.N)
            return fail(DVBS2B200_EINVAL, "row offset out of range");
    if (((uintptr_t)d_iq & 15) || ((uintptr_t)d_llr & 1)) // 128-bit symbol loads, 16-bit packed LLR stores
        return fail(DVBS2B200_EINVAL, "symbol buffer must be 16-byte aligned and the LLR buffer 2-byte aligned");
    if (n_syms % 2)
        return fail(DVBS2B200_EUNSUPPORTED, "symbols per frame must be even");
    for (int k = 0; k < bits; ++k)
        if (row_offsets[k] & 1)
            return fail(DVBS2B200_EINVAL, "row offsets must be even");
    DeviceGuard g(h->device);
    StreamOrder so(h, (cudaStream_t)stream);
    cudaStream_t s = (cudaStream_t)stream;
    // the constellation travels in the kernel parameters: nothing to upload, nothing cached on the device
    TableDemapConst tc;
    memset(&tc, 0, sizeof(tc));
    for (int i = 0; i < (1 << bits); ++i) {
        const float x = points[2 * i], y = points[2 * i + 1];
        tc.a[i] = -2.0f * x;
        tc.b[i] = -2.0f * y;
        tc.c[i] = x * x + y * y;
    }
    for (int k = 0; k < bits; ++k)
        tc.row[k] = row_offsets[k];
    TableDemapLaunch p;
    memset(&p, 0, sizeof(p));
    p.n_syms = n_syms;
    p.bits = bits;
    for (int f0 = 0; f0 < frames; f0 += 32768) {
        p.frames = std::min(32768, frames - f0);
        p.iq = d_iq + (size_t)f0 * n_syms * 2;
        p.n0 = d_n0 + f0;
        p.llr = d_llr + (size_t)f0 * hd.N;
        cudaError_t e = demap_table_launch(p, tc, s);
        if (e != cudaSuccess)
            return cuda_fail(e, "demap_table_launch");
        h->launches += 1;
    }
    return DVBS2B200_OK;
}

int dvbs2b200_demap_table(dvbs2b200_code* h, int bits, const float* points, const int* row_offsets, const float* iq, int frames,
                          const float* n0, int8_t* llr)
{
    if (!h)
        return fail(DVBS2B200_EINVAL, "null handle");
    if (bits < 1 || bits > 5)
        return fail(DVBS2B200_EINVAL, "bits must be 1..5");
    if (frames < 0)
        return fail(DVBS2B200_EINVAL, "negative frames");
    if (frames == 0)
        return DVBS2B200_OK;
    if (!iq || !n0 || !llr)
        return fail(DVBS2B200_EINVAL, "null buffer");
    DeviceGuard g(h->device);
    StreamOrder so(h, h->stream);
    const BlobHeader& hd = h->hdr;
    if (hd.N % bits)
        return fail(DVBS2B200_EINVAL, "frame length is not a multiple of the bits per symbol");
    const size_t iq_bytes = (size_t)frames * (hd.N / bits) * 8, llr_bytes = (size_t)frames * hd.N;
    int rc;
    if ((rc = h->d_in.ensure(iq_bytes)) || (rc = h->d_llr.ensure(llr_bytes)) || (rc = h->d_n0.ensure((size_t)frames * 4)))
        return rc;
    cudaStream_t s = h->stream;
    CU(cudaMemcpyAsync(h->d_in.p, iq, iq_bytes, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(h->d_n0.p, n0, (size_t)frames * 4, cudaMemcpyHostToDevice, s));
    rc = dvbs2b200_demap_table_dev(h, bits, points, row_offsets, (const float*)h->d_in.p, frames, (const float*)h->d_n0.p,
                                   (int8_t*)h->d_llr.p, s);
    if (rc)
        return rc;
    CU(cudaMemcpyAsync(llr, h->d_llr.p, llr_bytes, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return DVBS2B200_OK;
}

// ---- SNR estimate ---------------------------------------------------------------------------------
int dvbs2b200_estimate_snr_dev(dvbs2b200_code* h, int constellation, const float* d_iq, const int8_t* d_llr_post, int frames,
                               float* d_snr_lin, void* stream)
{
    if (!h)
        return fail(DVBS2B200_EINVAL, "null handle");
    DeviceGuard g(h->device);
    StreamOrder so(h, (cudaStream_t)stream);
    return snr_dev(h, constellation, d_iq, d_llr_post, frames, d_snr_lin, (cudaStream_t)stream);
}

int dvbs2b200_estimate_snr(dvbs2b200_code* h, int constellation, const float* iq, const int8_t* llr_post, int frames,
                           float* snr_lin)
{
    if (!h)
        return fail(DVBS2B200_EINVAL, "null handle");
    const int bits = bits_per_symbol(constellation);
    if (!bits)
        return fail(DVBS2B200_EUNSUPPORTED, "Unsupported constellation");
    if (frames < 0)
        return fail(DVBS2B200_EINVAL, "negative frames");
    if (frames == 0)
        return DVBS2B200_OK;
    if (!iq || !snr_lin)
        return fail(DVBS2B200_EINVAL, "null buffer");
    DeviceGuard g(h->device);
    StreamOrder so(h, h->stream);
    const BlobHeader& hd = h->hdr;
    const size_t iq_bytes = (size_t)frames * (hd.N / bits) * 8, llr_bytes = (size_t)frames * hd.N;
    int rc;
    if ((rc = h->d_in.ensure(iq_bytes)) || (rc = h->d_n0.ensure((size_t)frames * 4)))
        return rc;
    if (llr_post && (rc = h->d_llr.ensure(llr_bytes)))
        return rc;
    cudaStream_t s = h->stream;
    CU(cudaMemcpyAsync(h->d_in.p, iq, iq_bytes, cudaMemcpyHostToDevice, s));
    if (llr_post)
        CU(cudaMemcpyAsync(h->d_llr.p, llr_post, llr_bytes, cudaMemcpyHostToDevice, s));
    rc = snr_dev(h, constellation, (const float*)h->d_in.p, llr_post ? (const int8_t*)h->d_llr.p : nullptr, frames, (float*)h->d_n0.p, s);
    if (rc)
        return rc;
    CU(cudaMemcpyAsync(snr_lin, h->d_n0.p, (size_t)frames * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return DVBS2B200_OK;
}

// ---- BB layer: descrambler, deheader, and the chain down to TS bytes --------------------------------
int dvbs2b200_bb_descramble_dev(dvbs2b200_code* h, const uint8_t* d_in, int frames, uint8_t* d_out, void* stream)
{
    if (!h)
        return fail(DVBS2B200_EINVAL, "null handle");
    DeviceGuard g(h->device);
    StreamOrder so(h, (cudaStream_t)stream);
    return bb_descramble_dev(h, d_in, frames, d_out, (cudaStream_t)stream);
}

int dvbs2b200_bb_descramble(dvbs2b200_code* h, const uint8_t* in, int frames, uint8_t* out)
{
    if (!h)
        return fail(DVBS2B200_EINVAL, "null handle");
    if (frames < 0)
        return fail(DVBS2B200_EINVAL, "negative frames");
    if (frames == 0)
        return DVBS2B200_OK;
    if (!in || !out)
        return fail(DVBS2B200_EINVAL, "null buffer");
    DeviceGuard g(h->device);
    StreamOrder so(h, h->stream);
    const size_t bytes = (size_t)frames * (h->hdr.kbch / 8);
    int rc;
    if ((rc = h->d_mid.ensure(bytes)) || (rc = h->d_out.ensure(bytes)))
        return rc;
    cudaStream_t s = h->stream;
    CU(cudaMemcpyAsync(h->d_mid.p, in, bytes, cudaMemcpyHostToDevice, s));
    if ((rc = bb_descramble_dev(h, (const uint8_t*)h->d_mid.p, frames, (uint8_t*)h->d_out.p, s)))
        return rc;
    CU(cudaMemcpyAsync(out, h->d_out.p, bytes, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return DVBS2B200_OK;
}

size_t dvbs2b200_bb_ts_capacity(const dvbs2b200_code* h, int frames)
{
    return (h && frames > 0) ? bb_ts_capacity(h->hdr, frames) : 0;
}

int dvbs2b200_bb_deheader_dev(dvbs2b200_code* h, const uint8_t* d_bbframes, int frames, int scrambled, uint8_t* d_ts,
                              size_t ts_cap, void* stream)
{
    if (!h)
        return fail(DVBS2B200_EINVAL, "null handle");
    DeviceGuard g(h->device);
    StreamOrder so(h, (cudaStream_t)stream);
    return bb_deheader_dev(h, d_bbframes, frames, scrambled, d_ts, ts_cap, (cudaStream_t)stream);
}

int dvbs2b200_bb_deheader(dvbs2b200_code* h, const uint8_t* bbframes, int frames, int scrambled, uint8_t* ts, size_t ts_cap,
                          size_t* ts_bytes)
{
    if (!h)
        return fail(DVBS2B200_EINVAL, "null handle");
    if (frames < 0)
        return fail(DVBS2B200_EINVAL, "negative frames");
    if (ts_bytes)
        *ts_bytes = 0;
    if (frames == 0)
        return DVBS2B200_OK;
    if (!bbframes || !ts)
        return fail(DVBS2B200_EINVAL, "null buffer");
    DeviceGuard g(h->device);
    StreamOrder so(h, h->stream);
    const size_t bytes = (size_t)frames * (h->hdr.kbch / 8);
    const size_t cap = bb_ts_capacity(h->hdr, frames);
    if (ts_cap < cap)
        return fail(DVBS2B200_EINVAL, "ts buffer smaller than dvbs2b200_bb_ts_capacity(frames)");
    int rc;
    if ((rc = h->d_mid.ensure(bytes)) || (rc = h->d_ts.ensure(cap + 256)))
        return rc;
    cudaStream_t s = h->stream;
    CU(cudaMemcpyAsync(h->d_mid.p, bbframes, bytes, cudaMemcpyHostToDevice, s));
    if ((rc = bb_deheader_dev(h, (const uint8_t*)h->d_mid.p, frames, scrambled, (uint8_t*)h->d_ts.p, cap, s)))
        return rc;
    return bb_fetch_ts(h, (const uint8_t*)h->d_ts.p, ts, cap, ts_bytes, s);
}

int dvbs2b200_bb_produced_dev(dvbs2b200_code* h, void* stream, size_t* ts_bytes)
{
    if (!h || !ts_bytes)
        return fail(DVBS2B200_EINVAL, "null argument");
    DeviceGuard g(h->device);
    int rc = bb_ensure(h);
    if (rc)
        return rc;
    unsigned long long produced = 0;
    StreamOrder so(h, (cudaStream_t)stream);
    CU(cudaMemcpyAsync(&produced, &((BbState*)h->d_bbstate.p)->produced, sizeof(produced), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    *ts_bytes = (size_t)produced;
    return DVBS2B200_OK;
}

int dvbs2b200_bb_reset(dvbs2b200_code* h)
{
    if (!h)
        return fail(DVBS2B200_EINVAL, "null handle");
    DeviceGuard g(h->device);
    int rc = bb_ensure(h);
    if (rc)
        return rc;
    StreamOrder so(h, h->stream); // after whatever an earlier call left running on another stream
    CU(cudaMemsetAsync(h->d_bbstate.p, 0, sizeof(BbState), h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return DVBS2B200_OK;
}

int dvbs2b200_bb_counters_get(dvbs2b200_code* h, dvbs2b200_bb_counters* c)
{
    if (!h || !c)
        return fail(DVBS2B200_EINVAL, "null argument");
    DeviceGuard g(h->device);
    int rc = bb_ensure(h);
    if (rc)
        return rc;
    BbState st;
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaMemcpy(&st, h->d_bbstate.p, sizeof(st), cudaMemcpyDeviceToHost));
    c->packets = st.packet_cnt;
    c->errors = st.error_cnt;
    c->bbframes = st.bbframe_cnt;
    c->dropped = st.bbframe_drop_cnt;
    c->gaps = st.bbframe_gap_cnt;
    return DVBS2B200_OK;
}

int dvbs2b200_fec_decode_ts_dev(dvbs2b200_code* h, int constellation, const float* d_iq, const float* d_n0, const int8_t* d_llr,
                                int frames, int max_trials, int term_group, uint8_t* d_ts, size_t ts_cap, int32_t* d_trials_left,
                                int32_t* d_corrections, void* stream)
{
    if (!h)
        return fail(DVBS2B200_EINVAL, "null handle");
    if (frames < 0)
        return fail(DVBS2B200_EINVAL, "negative frames");
    if (frames == 0)
        return DVBS2B200_OK;
    DeviceGuard g(h->device);
    StreamOrder so(h, (cudaStream_t)stream);
    int rc;
    if ((rc = h->d_out.ensure((size_t)frames * (h->hdr.kbch / 8))))
        return rc;
    if ((rc = dvbs2b200_fec_decode_dev(h, constellation, d_iq, d_n0, d_llr, frames, max_trials, term_group, (uint8_t*)h->d_out.p,
                                       d_trials_left, d_corrections, stream)))
        return rc;
    return bb_deheader_dev(h, (const uint8_t*)h->d_out.p, frames, /*scrambled*/ 1, d_ts, ts_cap, (cudaStream_t)stream);
}

int dvbs2b200_fec_decode_ts(dvbs2b200_code* h, int constellation, const float* iq, const float* n0, const int8_t* llr, int frames,
                            int max_trials, int term_group, uint8_t* ts, size_t ts_cap, size_t* ts_bytes, int32_t* trials_left,
                            int32_t* corrections)
{
    if (!h)
        return fail(DVBS2B200_EINVAL, "null handle");
    if (frames < 0)
        return fail(DVBS2B200_EINVAL, "negative frames");
    if (ts_bytes)
        *ts_bytes = 0;
    if (frames == 0)
        return DVBS2B200_OK;
    if (!ts || (!iq && !llr))
        return fail(DVBS2B200_EINVAL, "null buffer");
    if (iq && !n0)
        return fail(DVBS2B200_EINVAL, "n0 is null");
    DeviceGuard g(h->device);
    StreamOrder so(h, h->stream);
    const BlobHeader& hd = h->hdr;
    const int bits = iq ? bits_per_symbol(constellation) : 0;
    if (iq && !bits)
        return fail(DVBS2B200_EUNSUPPORTED, "Unsupported constellation");
    const size_t in_bytes = (size_t)frames * (iq ? (size_t)(hd.N / bits) * 8 : (size_t)hd.N);
    const size_t cap = bb_ts_capacity(hd, frames);
    if (ts_cap < cap)
        return fail(DVBS2B200_EINVAL, "ts buffer smaller than dvbs2b200_bb_ts_capacity(frames)");
    int rc;
    if ((rc = h->d_in.ensure(in_bytes)) || (rc = h->d_ts.ensure(cap + 256)) || (rc = h->d_i32a.ensure((size_t)frames * 4)) ||
        (rc = h->d_i32b.ensure((size_t)frames * 4)))
        return rc;
    if (iq && (rc = h->d_n0.ensure((size_t)frames * 4)))
        return rc;
    cudaStream_t s = h->stream;
    CU(cudaMemcpyAsync(h->d_in.p, iq ? (const void*)iq : (const void*)llr, in_bytes, cudaMemcpyHostToDevice, s));
    if (iq)
        CU(cudaMemcpyAsync(h->d_n0.p, n0, (size_t)frames * 4, cudaMemcpyHostToDevice, s));
    // the BBFRAMEs stay in device memory: only TS bytes (and the per-frame status words) travel back
    rc = dvbs2b200_fec_decode_ts_dev(h, constellation, iq ? (const float*)h->d_in.p : nullptr, (const float*)h->d_n0.p,
                                     iq ? nullptr : (const int8_t*)h->d_in.p, frames, max_trials, term_group, (uint8_t*)h->d_ts.p, cap,
                                     (int32_t*)h->d_i32a.p, (int32_t*)h->d_i32b.p, s);
    if (rc)
        return rc;
    if (trials_left)
        CU(cudaMemcpyAsync(trials_left, h->d_i32a.p, (size_t)frames * 4, cudaMemcpyDeviceToHost, s));
    if (corrections)
        CU(cudaMemcpyAsync(corrections, h->d_i32b.p, (size_t)frames * 4, cudaMemcpyDeviceToHost, s));
    return bb_fetch_ts(h, (const uint8_t*)h->d_ts.p, ts, cap, ts_bytes, s);
}

// ---- one code on several devices of one process -----------------------------------------------------------
int dvbs2b200_multi_create(dvbs2b200_multi** out, const int* devices, int n_devices, int standard, int framesize, int rate)
{
    if (!out || !devices || n_devices <= 0 || n_devices > 64)
        return fail(DVBS2B200_EINVAL, "bad argument");
    *out = nullptr;
    // the tables are built ONCE on the host; every device gets the same blob with one host-to-device copy
    std::vector<uint8_t> blob;
    std::string err;
    if (!build_blob(standard, framesize, rate, blob, err))
        return fail(DVBS2B200_EUNSUPPORTED, err);
    dvbs2b200_multi* m = new (std::nothrow) dvbs2b200_multi();
    if (!m)
        return fail(DVBS2B200_ENOMEM, "out of host memory");
    for (int i = 0; i < n_devices; ++i) {
        dvbs2b200_code* h = nullptr;
        int rc = create_from_blob(&h, devices[i], std::vector<uint8_t>(blob), /*validated=*/true);
        if (rc) {
            dvbs2b200_multi_destroy(m);
            return rc;
        }
        m->codes.push_back(h);
    }
    *out = m;
    return DVBS2B200_OK;
}

void dvbs2b200_multi_destroy(dvbs2b200_multi* m)
{
    if (!m)
        return;
    for (dvbs2b200_code* h : m->codes)
        dvbs2b200_code_destroy(h);
    delete m;
}

int dvbs2b200_multi_device_count(const dvbs2b200_multi* m) { return m ? (int)m->codes.size() : 0; }

dvbs2b200_code* dvbs2b200_multi_code(dvbs2b200_multi* m, int index)
{
    return (m && index >= 0 && index < (int)m->codes.size()) ? m->codes[index] : nullptr;
}

int dvbs2b200_multi_shard(const dvbs2b200_multi* m, int frames, int index, int* first, int* count)
{
    if (!m || !first || !count || index < 0 || index >= (int)m->codes.size() || frames < 0)
        return fail(DVBS2B200_EINVAL, "bad argument");
    // contiguous ranges whose boundaries fall on multiples of 32 frames, so that a batch-coupled termination
    // group (16 / 32 frames, lib/ldpc_decoder/layered_decoder.hh:153) never straddles two devices
    const int world = (int)m->codes.size();
    const int units = (frames + 31) / 32, per = units / world, extra = units % world;
    const int lo = index * per + std::min(index, extra), hi = lo + per + (index < extra ? 1 : 0);
    *first = std::min(lo * 32, frames);
    *count = std::min(hi * 32, frames) - *first;
    return DVBS2B200_OK;
}

int dvbs2b200_multi_fec_decode(dvbs2b200_multi* m, int constellation, const float* iq, const float* n0, const int8_t* llr, int frames,
                               int max_trials, int term_group, uint8_t* msg, int32_t* trials_left, int32_t* corrections)
{
    if (!m || m->codes.empty())
        return fail(DVBS2B200_EINVAL, "null handle");
    if (frames < 0)
        return fail(DVBS2B200_EINVAL, "negative frames");
    if (frames == 0)
        return DVBS2B200_OK;
    if (!msg || (!iq && !llr))
        return fail(DVBS2B200_EINVAL, "null buffer");
    const int world = (int)m->codes.size();
    const BlobHeader& hd = m->codes[0]->hdr;
    const int bits = iq ? bits_per_symbol(constellation) : 0;
    if (iq && !bits)
        return fail(DVBS2B200_EUNSUPPORTED, "Unsupported constellation");
    std::vector<int> rcs(world, DVBS2B200_OK);
    std::vector<std::string> errs(world);
    // one worker per device: the caller stays one thread; the workers stage and launch concurrently
    auto work = [&](int i) {
        int f0 = 0, nf = 0;
        dvbs2b200_multi_shard(m, frames, i, &f0, &nf);
        if (nf == 0)
            return;
        rcs[i] = dvbs2b200_fec_decode(m->codes[i], constellation, iq ? iq + (size_t)f0 * (hd.N / bits) * 2 : nullptr, n0 ? n0 + f0 : nullptr,
                                      llr ? llr + (size_t)f0 * hd.N : nullptr, nf, max_trials, term_group,
                                      msg + (size_t)f0 * (hd.kbch / 8), trials_left ? trials_left + f0 : nullptr,
                                      corrections ? corrections + f0 : nullptr);
        if (rcs[i])
            errs[i] = g_err; // thread-local in the worker
    };
    std::vector<std::thread> th;
    for (int i = 1; i < world; ++i)
        th.emplace_back(work, i);
    work(0);
    for (auto& t : th)
        t.join();
    for (int i = 0; i < world; ++i)
        if (rcs[i])
            return fail(rcs[i], "device " + std::to_string(m->codes[i]->device) + ": " + errs[i]);
    return DVBS2B200_OK;
}

// ---- PL descrambler + pilot-segment de-rotation ---------------------------------------------------------------
struct dvbs2b200_pl {
    int device = 0;
    int gold_code = 0;
    cudaStream_t stream = nullptr;
    DevBuf d_rn, d_in, d_out, d_info;
};

} // extern "C"

namespace {
constexpr int kPlMaxPayload = 360 * 90 + 22 * 36; // MAX_PLFRAME_PAYLOAD, lib/pl_defs.h:29
// lib/pl_descrambler.cc:36-98: x and y shift registers of EN 302 307-1 5.5.4, x advanced by the Gold code
int pl_parity18(long a, long b)
{
    a &= b;
    int c = 0;
    for (int i = 0; i < 18; ++i)
        c += (int)((a >> i) & 1);
    return c & 1;
}
void pl_scrambling_codes(int gold_code, std::vector<uint8_t>& rn)
{
    rn.assign(kPlMaxPayload + 8, 0);
    long x = 0x00001, y = 0x3FFFF;
    for (int n = 0; n < gold_code; ++n) {
        const int xb = pl_parity18(x, 0x0081);
        x >>= 1;
        if (xb)
            x |= 0x20000;
    }
    for (int i = 0; i < kPlMaxPayload; ++i) {
        const int xa = pl_parity18(x, 0x8050), xb = pl_parity18(x, 0x0081), xc = (int)(x & 1);
        x >>= 1;
        if (xb)
            x |= 0x20000;
        const int ya = pl_parity18(y, 0x04A1), yb = pl_parity18(y, 0xFF60), yc = (int)(y & 1);
        y >>= 1;
        if (ya)
            y |= 0x20000;
        rn[i] = (uint8_t)((((xa ^ yb) & 1) << 1) | ((xc ^ yc) & 1));
    }
}
int pl_payload_len(int n_slots, int has_pilots) { return n_slots * 90 + (has_pilots ? ((n_slots - 1) / 16) * 36 : 0); }

int pl_dev(dvbs2b200_pl* h, const float* d_payload, int frames, int n_slots, int has_pilots, const dvbs2b200_pl_frame* d_info,
           float* d_out, cudaStream_t stream)
{
    if (frames < 0)
        return fail(DVBS2B200_EINVAL, "negative frames");
    if (n_slots < 1 || n_slots > 360)
        return fail(DVBS2B200_EINVAL, "n_slots must be 1..360");
    if (frames == 0)
        return DVBS2B200_OK;
    if (!d_payload || !d_info || !d_out)
        return fail(DVBS2B200_EINVAL, "null buffer");
    if (((uintptr_t)d_payload & 15) || ((uintptr_t)d_out & 15))
        return fail(DVBS2B200_EINVAL, "symbol buffers must be 16-byte aligned");
    static_assert(sizeof(dvbs2b200_pl_frame) == sizeof(PlFrameInfo), "dvbs2b200_pl_frame layout");
    PlLaunch p;
    memset(&p, 0, sizeof(p));
    p.payload = d_payload;
    p.out = d_out;
    p.rn = (const uint8_t*)h->d_rn.p;
    p.info = reinterpret_cast<const PlFrameInfo*>(d_info);
    p.frames = frames;
    p.n_slots = n_slots;
    p.has_pilots = has_pilots ? 1 : 0;
    p.payload_len = pl_payload_len(n_slots, has_pilots);
    cudaError_t e = pl_launch(p, stream);
    if (e != cudaSuccess)
        return cuda_fail(e, "pl_launch");
    return DVBS2B200_OK;
}
} // namespace

extern "C" {

int dvbs2b200_pl_create(dvbs2b200_pl** out, int device, int gold_code)
{
    if (!out || gold_code < 0 || gold_code > 262141)
        return fail(DVBS2B200_EINVAL, "gold_code must be 0..262141");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0)
        return fail(DVBS2B200_ECUDA, std::string("no CUDA device available (") + cudaGetErrorString(e) + "); libdvbs2_b200 has no CPU fallback");
    if (device < 0 || device >= ndev)
        return fail(DVBS2B200_EINVAL, "device index out of range");
    dvbs2b200_pl* h = new (std::nothrow) dvbs2b200_pl();
    if (!h)
        return fail(DVBS2B200_ENOMEM, "out of host memory");
    h->device = device;
    h->gold_code = gold_code;
    DeviceGuard g(device);
    std::vector<uint8_t> rn;
    pl_scrambling_codes(gold_code, rn);
    int rc = h->d_rn.ensure(rn.size());
    if (!rc && (e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess)
        rc = cuda_fail(e, "cudaStreamCreate");
    if (!rc && (e = cudaMemcpyAsync(h->d_rn.p, rn.data(), rn.size(), cudaMemcpyHostToDevice, h->stream)) != cudaSuccess)
        rc = cuda_fail(e, "cudaMemcpy(scrambling codes)");
    if (!rc && (e = cudaStreamSynchronize(h->stream)) != cudaSuccess)
        rc = cuda_fail(e, "cudaStreamSynchronize");
    if (!rc && (e = pl_preload()) != cudaSuccess)
        rc = cuda_fail(e, "kernel preload");
    if (rc) {
        dvbs2b200_pl_destroy(h);
        return rc;
    }
    *out = h;
    return DVBS2B200_OK;
}

void dvbs2b200_pl_destroy(dvbs2b200_pl* h)
{
    if (!h)
        return;
    DeviceGuard g(h->device);
    if (h->stream) {
        cudaStreamSynchronize(h->stream);
        cudaStreamDestroy(h->stream);
    }
    for (DevBuf* b : { &h->d_rn, &h->d_in, &h->d_out, &h->d_info })
        b->release();
    delete h;
}

int dvbs2b200_pl_scrambling_codes(int gold_code, uint8_t* rn, int n)
{
    if (!rn || n < 0 || n > kPlMaxPayload || gold_code < 0 || gold_code > 262141)
        return fail(DVBS2B200_EINVAL, "bad argument");
    std::vector<uint8_t> all;
    pl_scrambling_codes(gold_code, all);
    memcpy(rn, all.data(), (size_t)n);
    return DVBS2B200_OK;
}

int dvbs2b200_pl_payload_len(int n_slots, int has_pilots) { return pl_payload_len(n_slots, has_pilots); }

int dvbs2b200_pl_descramble_derotate_dev(dvbs2b200_pl* h, const float* d_payload, int frames, int n_slots, int has_pilots,
                                         const dvbs2b200_pl_frame* d_info, float* d_xfecframe, void* stream)
{
    if (!h)
        return fail(DVBS2B200_EINVAL, "null handle");
    DeviceGuard g(h->device);
    return pl_dev(h, d_payload, frames, n_slots, has_pilots, d_info, d_xfecframe, (cudaStream_t)stream);
}

int dvbs2b200_pl_descramble_derotate(dvbs2b200_pl* h, const float* payload, int frames, int n_slots, int has_pilots,
                                     const dvbs2b200_pl_frame* info, float* xfecframe)
{
    if (!h)
        return fail(DVBS2B200_EINVAL, "null handle");
    if (frames < 0)
        return fail(DVBS2B200_EINVAL, "negative frames");
    if (n_slots < 1 || n_slots > 360)
        return fail(DVBS2B200_EINVAL, "n_slots must be 1..360");
    if (frames == 0)
        return DVBS2B200_OK;
    if (!payload || !info || !xfecframe)
        return fail(DVBS2B200_EINVAL, "null buffer");
    DeviceGuard g(h->device);
    const size_t in_bytes = (size_t)frames * pl_payload_len(n_slots, has_pilots) * 8, out_bytes = (size_t)frames * n_slots * 90 * 8;
    int rc;
    if ((rc = h->d_in.ensure(in_bytes)) || (rc = h->d_out.ensure(out_bytes)) || (rc = h->d_info.ensure((size_t)frames * sizeof(dvbs2b200_pl_frame))))
        return rc;
    cudaStream_t s = h->stream;
    CU(cudaMemcpyAsync(h->d_in.p, payload, in_bytes, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(h->d_info.p, info, (size_t)frames * sizeof(dvbs2b200_pl_frame), cudaMemcpyHostToDevice, s));
    rc = pl_dev(h, (const float*)h->d_in.p, frames, n_slots, has_pilots, (const dvbs2b200_pl_frame*)h->d_info.p, (float*)h->d_out.p, s);
    if (rc) {
        cudaStreamSynchronize(s);
        return rc;
    }
    CU(cudaMemcpyAsync(xfecframe, h->d_out.p, out_bytes, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return DVBS2B200_OK;
}

} // extern "C"
