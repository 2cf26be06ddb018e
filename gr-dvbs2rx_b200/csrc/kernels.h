// kernels.h -- launch descriptors shared by the sm_100a kernels and the C ABI (capi.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dvbs2b200 {

constexpr int kLdpcThreads = 192;  // 180 check-node pairs of a layer + 12 spare lanes = 6 warps
constexpr int kLdpcCtasPerSm = 3;  // 3 x (N + tables) bytes of shared memory fit one SM

struct LdpcLaunch {
    // code
    int N, K, R, q, n_circ, n_steps;
    const uint8_t* tab;   // device: [LayerRec q][EdgeRec n_circ][StepRec n_steps], 16-byte aligned
    uint32_t tab_bytes;   // multiple of 16
    const uint16_t* work; // level tables of the split steps
    // shared-memory carve-up (bytes from the start of dynamic shared memory)
    uint32_t smem_tab_off, smem_bar_off, smem_rec_off; // rec: scratch of the split steps
    // per-CTA check-node state, [grid][R/2 pairs x (1 + ceil(deg/8)) words] uint32, L2 resident
    uint32_t* msg_scratch;
    // two constants the compiler must not see through (they keep shifts / adds on the FMA pipe, ldpc_core.cuh)
    uint32_t two, four, neg1, c30, c16, c32;
    // batch
    const int8_t* llr; // [frames][N], 4-byte aligned
    // streaming input: *ready counts the chunks of ready_chunk frames that have arrived (null: all there)
    const unsigned int* ready;
    int ready_chunk;
    unsigned int* err; // set to 1 by a CTA that gave up waiting on *ready (null: not reported)
    int frames;
    int max_trials;
    int group;           // 0 per-frame termination, else frames per coupled group
    unsigned int* gsync; // [frames/group][max_trials + 2], zeroed (group mode only)
    // per-frame termination: frames beyond the first wave are handed out through this counter (zeroed before the launch)
    // instead of by stride, so a CTA whose frames converge early takes more of them (null: by stride)
    unsigned int* next_frame;
    uint8_t* hard;       // [frames][out_bytes] or null
    int out_bytes;
    int8_t* llr_post;     // [frames][N] or null, 4-byte aligned
    int32_t* trials_left; // [frames] or null
    // optional per-CTA cycle counters [grid][16] (diagnostics build: -DDVBS2_PHASE_PROFILE, DVBS2B200_PHASE_PROFILE)
    unsigned long long* prof;
    int sm_count;
};

// fills the smem_* offsets of p (if non-null) and returns the dynamic shared memory size
size_t ldpc_smem_bytes(int N, uint32_t tab_bytes, uint32_t scratch_bytes, LdpcLaunch* p);
cudaError_t ldpc_launch(const LdpcLaunch& p, int max_cnt, bool uniform, int grid, size_t smem, cudaStream_t stream);
// resident CTAs per SM for this code's kernel instantiation (occupancy query)
int ldpc_ctas_per_sm(int N, int max_cnt, bool uniform, size_t smem);

struct BchLaunch {
    const uint8_t* cw; // [frames][n_bytes]
    uint8_t* msg;      // [frames][k_bytes]
    int32_t* corrections; // [frames] or null
    int frames;
    int n, k, t, m;    // bits, bits, capability, GF(2^m)
    uint32_t shorten;  // s = 2^m - 1 - n
    const uint16_t* antilog; // [2^m]: alpha^i for i <= 2^m - 1
    const uint16_t* log;     // [2^m]
    int cw_stride, msg_stride; // bytes between consecutive frames
};
constexpr int kBchWarpsPerBlock = 32; // persistent CTAs, one per SM
cudaError_t bch_launch(const BchLaunch& p, cudaStream_t stream);

struct DemapLaunch {
    const float* iq; // [frames][n_syms][2]
    const float* n0; // [frames]
    int8_t* llr;     // [frames][N]
    int frames, n_syms;
    int constellation; // 0 QPSK, 4 8PSK
    int row0, row1, row2; // 8PSK deinterleaver row offsets
};
cudaError_t demap_launch(const DemapLaunch& p, cudaStream_t stream);

struct TableDemapLaunch {
    const float* iq;     // [frames][n_syms][2], 16-byte aligned
    const float* n0;     // [frames]
    int8_t* llr;         // [frames][n_syms * bits], 4-byte aligned
    int frames, n_syms, bits;
};
// the constellation, by value in the kernel parameters: point s = the symbol's bits, first bit = MSB
struct TableDemapConst {
    float a[32], b[32], c[32]; // -2 Re s, -2 Im s, |s|^2
    int row[5];                // bit k of symbol j -> llr[row[k] + j]
};
cudaError_t demap_table_launch(const TableDemapLaunch& p, const TableDemapConst& t, cudaStream_t stream);

struct SnrLaunch {
    const float* iq;   // [frames][n_syms][2]
    const int8_t* llr; // null: slice the symbols; else [frames][N] posterior LLRs (codeword order)
    float* snr_lin;    // [frames] linear Es/N0
    int frames, n_syms;
    int constellation; // 0 QPSK, 4 8PSK
    int row0, row1, row2;
};
cudaError_t snr_launch(const SnrLaunch& p, cudaStream_t stream);

// ---- BB layer: descrambler + deheader (bb_kernel.cu) --------------------------------------------------
// Stream state of the deheader (lib/bbdeheader_bb_impl.h:40-53), in device memory, one per handle.
struct BbState {
    int synched;          // d_synched
    unsigned int partial; // d_partial_ts_bytes
    int carry_idx;        // which carry buffer holds d_partial_pkt after the call
    int read_idx;         // which one held it before (read by the call's first completed packet)
    int carry_src_frame;  // frame of this call whose tail becomes the new partial packet, or -1
    unsigned int carry_src_off, carry_len;
    unsigned int pad;
    unsigned long long packet_cnt, error_cnt, bbframe_cnt, bbframe_drop_cnt, bbframe_gap_cnt;
    unsigned long long produced; // TS bytes written by the last call
    uint8_t carry[2][192];
};
// Per BBFRAME: what the frame contributes to the output.
struct BbPlan {
    uint32_t out_pkt;  // index of its first packet in the output of this call
    uint32_t n_pkts;
    uint32_t p_in;     // bytes of a carried partial packet in front of its first packet
    uint32_t skip;     // DATAFIELD bytes skipped on re-synchronisation (syncd/8 + 1)
    int32_t src_frame; // where those p_in bytes are: a frame of this call, or -1 = the state's carry buffer
    uint32_t src_off;  // offset into that frame's DATAFIELD
};
struct BbLaunch {
    const uint8_t* bb;   // [frames][kbytes] BBFRAMEs (BCH output)
    const uint8_t* prbs; // [kbytes] descrambling sequence
    int scrambled;       // 1: bb is still scrambled (descramble on the fly), 0: already descrambled
    int frames, kbytes;
    uint32_t* rec;       // [frames] scratch
    BbPlan* plan;        // [frames] scratch
    BbState* state;
    uint8_t* ts;         // output, capacity ts_cap bytes
    unsigned long long ts_cap;
};
cudaError_t bb_descramble_launch(const uint8_t* in, uint8_t* out, const uint8_t* prbs, int frames, int kbytes, cudaStream_t stream);
// three launches: header records, state scan, packet extraction
cudaError_t bb_deheader_launch(const BbLaunch& p, cudaStream_t stream);

// ---- mixed-MODCOD batches: gather / scatter of variable-size frames (mixed_kernel.cu) --------------------
cudaError_t gather_launch(const uint8_t* src, const unsigned long long* off, uint8_t* dst, int bytes, int frames, cudaStream_t stream);
cudaError_t scatter_launch(const uint8_t* src, const unsigned long long* off, uint8_t* dst, int bytes, int frames, const int32_t* v0,
                           const int32_t* v1, const int32_t* pos, int32_t* o0, int32_t* o1, cudaStream_t stream);


// ---- PL descrambler + pilot-segment de-rotation (pl_kernel.cu) ----------------------------------------------
// per PLFRAME: what the frame synchroniser estimated (lib/plsync_cc_impl.cc: plframe_info_t, freq_sync pilot phases);
// the layout is dvbs2b200_pl_frame of include/dvbs2_b200.h
struct PlFrameInfo {
    float plheader_phase;
    float fine_foffset;   // normalised (cycles per symbol); used only when coarse_corrected
    int coarse_corrected;
    int reserved;
    float pilot_phase[22];
};
struct PlLaunch {
    const float* payload;    // [frames][payload_len][2], 16-byte aligned rows (payload_len even)
    float* out;              // [frames][n_slots * 90][2]
    const uint8_t* rn;       // [>= payload_len] scrambling codes 0..3
    const PlFrameInfo* info; // [frames] (device)
    int frames, n_slots, has_pilots, payload_len;
};
cudaError_t pl_launch(const PlLaunch& p, cudaStream_t stream);
cudaError_t pl_preload();

// load the kernels of a translation unit ahead of their first launch (see bch_kernel.cu)
cudaError_t bch_preload();
cudaError_t bb_preload();
cudaError_t demap_preload();
cudaError_t mixed_preload();

} // namespace dvbs2b200
