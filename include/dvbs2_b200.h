/*
 * dvbs2_b200.h -- C ABI of libdvbs2_b200.so: the B200 (sm_100a) DVB-S2 FEC decode path
 * (soft demap -> layered offset-min-sum LDPC -> BCH) that drops in behind gr-dvbs2rx's
 * xfecframe_demapper_cb / ldpc_decoder_bb / bch_decoder_bb blocks.
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns 0 (DVBS2B200_OK) or a negative
 *     DVBS2B200_E*; nothing throws, nothing prints.  dvbs2b200_last_error() gives the text.
 *   - enum ordinals (standard, framesize, rate, constellation, output mode) are the
 *     reference's: include/gnuradio/dvbs2rx/dvb_config.h:15-121.
 *   - bit order is MSB first inside a byte, frames are contiguous
 *     (lib/ldpc_decoder_bb_impl.cc:432-442, lib/bch.cc:436-450).
 *   - a handle is bound to one CUDA device and one code (the reference is CCM: one MODCOD
 *     per block instance) and is used by one host thread at a time (GNU Radio calls
 *     general_work from one thread per block).
 *   - "host" entry points take host pointers and return when the results are in them.  Pinned /
 *     registered memory is copied from directly; pageable memory (GNU Radio's buffers) is staged
 *     through the handle's ring of pinned slots by dvbs2b200_fec_decode (the streaming entry
 *     point); the other host entry points hand pageable pointers to the driver's own staging.
 *   - "_dev" entry points take device pointers and a cudaStream_t (as void*) and are
 *     asynchronous on that stream.  A handle's scratch is shared by all its calls, so the library
 *     orders them: a call on another stream than the previous one first waits (on the device)
 *     for the previous call's work.  Symbol buffers must be 16-byte, LLR buffers 4-byte aligned.
 *   - there is NO CPU fallback: without a usable CUDA device every compute entry point
 *     fails with DVBS2B200_ECUDA.
 *
 * Reference interfaces replaced (file:line under the reference checkout):
 *   dvbs2b200_ldpc_decode   <- int (*decode)(void*, int8_t*, int)   lib/ldpc_decoder_bb_impl.h:41,
 *                              ldpc_<isa>::ldpc_dec_decode           lib/ldpc_decoder_bb_impl.cc:34-52,
 *                              + hard decision / packing loop        lib/ldpc_decoder_bb_impl.cc:406-447
 *   dvbs2b200_code_create   <- ldpc_<isa>::ldpc_dec_init + table/ISA dispatch
 *                                                                    lib/ldpc_decoder_bb_impl.cc:97-350,
 *                              galois_field / bch_codec construction lib/bch_decoder_bb_impl.cc:56-70
 *   dvbs2b200_bch_decode    <- d_codec->decode(in, out)              lib/bch_decoder_bb_impl.cc:94-111
 *                              (bch_codec::decode(u8)                lib/bch.cc:467-487)
 *   dvbs2b200_demap         <- d_qpsk->demap_soft / d_mod->soft + deinterleave
 *                                                                    lib/xfecframe_demapper_cb_impl.cc:151-176
 *   dvbs2b200_fec_decode    <- the three general_work bodies chained (demapper -> LDPC -> BCH)
 *   dvbs2b200_lookup        <- get_fec_info                          lib/fec_params.cc:16-344
 */
#ifndef DVBS2_B200_H
#define DVBS2_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DVBS2B200_VERSION 100 /* 0.1.0 */

enum {
    DVBS2B200_OK = 0,
    DVBS2B200_EINVAL = -1,       /* bad argument */
    DVBS2B200_EUNSUPPORTED = -2, /* code / constellation not supported (yet) */
    DVBS2B200_ECUDA = -3,        /* no device, launch or runtime failure */
    DVBS2B200_ENOMEM = -4
};

/* termination coupling of the LDPC iteration loop */
enum {
    DVBS2B200_TERM_PER_FRAME = 0 /* each frame stops when ITS syndrome is clean */
    /* 16 or 32: the reference's SIMD-batch semantics (lib/ldpc_decoder/layered_decoder.hh:153):
     * a group of that many consecutive frames iterates until all of them are clean. */
};

typedef struct dvbs2b200_code dvbs2b200_code; /* opaque */

typedef struct {
    int table;      /* index into the LDPC table list, dvbs2b200_table_name(table) */
    int n_ldpc;     /* LDPC codeword bits (16200 / 32400 / 64800) */
    int k_ldpc;     /* LDPC information bits of the table */
    int q;          /* (n_ldpc - k_ldpc) / 360 = number of layers */
    int n_circ;     /* 360x360 circulants in the data part */
    int links_total;
    int max_cn_deg;
    int kbch, nbch, t; /* lib/fec_params.cc; kldpc used for OM_MESSAGE output is nbch */
    int gf_m;          /* 16 normal, 14 short, 15 medium (lib/bch_decoder_bb_impl.cc:58-63) */
} dvbs2b200_code_info;

/* ---- library ---------------------------------------------------------------------------- */
int dvbs2b200_version(void);
const char* dvbs2b200_last_error(void); /* thread-local, never NULL */
int dvbs2b200_device_count(void);       /* >= 0, or DVBS2B200_ECUDA */
/* Page-lock a host buffer the caller keeps across calls (a GNU Radio block's ring buffers live as long as the
 * flowgraph: register them once in start(), unregister in stop()).  The host entry points then copy from / to it
 * directly instead of staging through the handle's pinned ring (bench.py: e2e vs e2e_pageable).  Thin wrappers of
 * cudaHostRegister(portable) / cudaHostUnregister; registering the same range twice is an error of the driver's. */
int dvbs2b200_host_register(void* ptr, size_t bytes);
int dvbs2b200_host_unregister(void* ptr);

/* ---- code parameters (pure host functions, usable without a GPU) -------------------------- */
int dvbs2b200_num_tables(void);
const char* dvbs2b200_table_name(int table);
int dvbs2b200_lookup(int standard, int framesize, int rate, dvbs2b200_code_info* info);
/* circulant list of a table: word = layer << 17 | group << 9 | shift; returns the count */
int dvbs2b200_table_circulants(int table, uint32_t* out, int cap);
/* BCH generator polynomial of a code, g[i] = coefficient of x^i; returns its degree */
int dvbs2b200_bch_genpoly(int framesize, int t, uint8_t* g, int cap);
/* per-layer serial-order schedule statistics (wavefront steps per iteration, deepest layer) */
int dvbs2b200_schedule_stats(int table, int* steps_per_iter, int* max_depth, int* conflict_layers);

/* ---- handles ------------------------------------------------------------------------------ */
/* Builds the packed code tables on the host (circulants, serial-order schedule, GF log/antilog,
 * generator data), uploads them to `device`, creates the stream and staging buffers. */
int dvbs2b200_code_create(dvbs2b200_code** h, int device, int standard, int framesize, int rate);
/* The packed tables as one relocatable blob, so that rank 0 can build them once and
 * broadcast them (one ncclBroadcast through torch.distributed) to the other ranks.
 * dvbs2b200_tables_build is the host-only half of dvbs2b200_code_create (no device needed):
 * call with buf == NULL to get the size. */
int dvbs2b200_tables_build(int standard, int framesize, int rate, void* buf, size_t cap, size_t* size);
int dvbs2b200_code_export_tables(const dvbs2b200_code* h, void* buf, size_t cap, size_t* size);
int dvbs2b200_code_create_from_tables(dvbs2b200_code** h, int device, const void* blob, size_t size);
void dvbs2b200_code_destroy(dvbs2b200_code* h);
int dvbs2b200_code_info_get(const dvbs2b200_code* h, dvbs2b200_code_info* info);
/* number of kernel launches issued through this handle so far (bench.py's gpu_launches) */
uint64_t dvbs2b200_launch_count(const dvbs2b200_code* h);

/* ---- LDPC --------------------------------------------------------------------------------- */
/* llr        [frames][n_ldpc] int8, frame-major (the block's input stream)
 * max_trials 0 -> 25 (lib/ldpc_decoder_bb_impl.cc:391,402)
 * term_group DVBS2B200_TERM_PER_FRAME, or 16 / 32 for the reference's batch-coupled loop
 *            (frames must then be a multiple of term_group)
 * output_mode OM_CODEWORD = 0 -> n_ldpc/8 bytes per frame, OM_MESSAGE = 1 -> nbch/8 bytes
 * hard       [frames][bytes] packed hard decisions (llr < 0 -> 1), MSB first
 * llr_post   NULL or [frames][n_ldpc] posterior LLRs (what the reference leaves in `code`
 *            and publishes on "llr_pdu")
 * trials_left NULL or [frames]: the reference's return value (>= 0 trials left, -1 = not
 *            converged); in group mode every frame of a group reports the group's value. */
int dvbs2b200_ldpc_decode(dvbs2b200_code* h, const int8_t* llr, int frames, int max_trials,
                          int term_group, int output_mode, uint8_t* hard, int8_t* llr_post,
                          int32_t* trials_left);
int dvbs2b200_ldpc_decode_dev(dvbs2b200_code* h, const int8_t* d_llr, int frames, int max_trials,
                              int term_group, int output_mode, uint8_t* d_hard,
                              int8_t* d_llr_post, int32_t* d_trials_left, void* stream);

/* ---- BCH ---------------------------------------------------------------------------------- */
/* cw [frames][nbch/8] -> msg [frames][kbch/8]; corrections[f] = #bits corrected, 0, or -1
 * (uncorrectable; the roots that were found are still flipped, lib/bch.cc:476-483). */
int dvbs2b200_bch_decode(dvbs2b200_code* h, const uint8_t* cw, int frames, uint8_t* msg,
                         int32_t* corrections);
int dvbs2b200_bch_decode_dev(dvbs2b200_code* h, const uint8_t* d_cw, int frames, uint8_t* d_msg,
                             int32_t* d_corrections, void* stream);

/* ---- soft demapper ------------------------------------------------------------------------ */
/* iq [frames][n_ldpc/bits][2] float (XFECFRAME symbols, pilots already removed)
 * n0 [frames] noise variance per frame -- explicit, the library has no hidden SNR feedback
 * llr [frames][n_ldpc] int8, deinterleaved to codeword order.
 * constellation: MOD_QPSK (0) or MOD_8PSK (4); others -> DVBS2B200_EUNSUPPORTED, as the
 * reference throws "Unsupported constellation" (lib/xfecframe_demapper_cb_impl.cc:70-72). */
int dvbs2b200_demap(dvbs2b200_code* h, int constellation, const float* iq, int frames,
                    const float* n0, int8_t* llr);
int dvbs2b200_demap_dev(dvbs2b200_code* h, int constellation, const float* d_iq, int frames,
                        const float* d_n0, int8_t* d_llr, void* stream);

/* ---- table-driven soft demapper: 16APSK, 32APSK, any constellation of up to 32 points ------------- */
/* NOT a replacement of reference code: gr-dvbs2rx demaps QPSK and 8PSK only and throws "Unsupported
 * constellation" otherwise (lib/xfecframe_demapper_cb_impl.cc:70-72), as dvbs2b200_demap does.  This entry
 * point is SURVEY 8f rank 3, defined by this library:
 *   llr_k = ( min_{s: bit k = 1} |y - s|^2 - min_{s: bit k = 0} |y - s|^2 ) / N0   (max-log; positive = bit 0,
 *   the convention of lib/qpsk.h:208-214; rounded to nearest even, saturated to int8)
 *   points      [2^bits][2] constellation, index = the symbol's bits with the first bit as MSB (host pointer)
 *   row_offsets [bits]: bit k of symbol j goes to llr[f][row_offsets[k] + j] -- the DVB-S2 bit interleaver read
 *               back (n_ldpc/bits rows, `bits` columns), whatever the column order of the MODCOD (host pointer)
 * dvbs2rx_b200.apsk (Python) builds the EN 302 307-1 16APSK / 32APSK tables for a ring ratio. */
int dvbs2b200_demap_table(dvbs2b200_code* h, int bits, const float* points, const int* row_offsets,
                          const float* iq, int frames, const float* n0, int8_t* llr);
int dvbs2b200_demap_table_dev(dvbs2b200_code* h, int bits, const float* points, const int* row_offsets,
                              const float* d_iq, int frames, const float* d_n0, int8_t* d_llr, void* stream);

/* ---- mixed-MODCOD (VCM/ACM) batches ------------------------------------------------------------ */
/* The reference's blocks are CCM: one MODCOD per block instance (lib/ldpc_decoder_bb_impl.cc:79-368), a VCM
 * receiver runs one chain per MODCOD.  A dvbs2b200_mixed holds one code handle per MODCOD of the stream on
 * one device and decodes a batch whose frames carry a per-frame code id (BASELINE config 5): the frames are
 * bucketed by code, each bucket runs through that code's LDPC + BCH kernels on its own stream (buckets
 * overlap on the device), results return to the frames' positions.
 *   code_id [frames]  index into the set given to dvbs2b200_mixed_create
 *   llr     the frames' soft input back to back, n_ldpc(code_id[f]) bytes each
 *   msg     the BBFRAMEs back to back, kbch(code_id[f])/8 bytes each
 * Per-frame termination (term_group 0).  Results are those of dvbs2b200_fec_decode on each code's frames. */
typedef struct dvbs2b200_mixed dvbs2b200_mixed;
int dvbs2b200_mixed_create(dvbs2b200_mixed** m, int device, int n_codes, const int* standard, const int* framesize,
                           const int* rate);
void dvbs2b200_mixed_destroy(dvbs2b200_mixed* m);
int dvbs2b200_mixed_code_info(const dvbs2b200_mixed* m, int code, dvbs2b200_code_info* info);
int dvbs2b200_mixed_fec_decode(dvbs2b200_mixed* m, int frames, const uint8_t* code_id, const int8_t* llr,
                               int max_trials, uint8_t* msg, int32_t* trials_left, int32_t* corrections);
/* device variant: d_llr / d_msg / status arrays in device memory (code_id stays a host array: it is control
 * data, read before the call returns); asynchronous on `stream`, which waits for every bucket. */
int dvbs2b200_mixed_fec_decode_dev(dvbs2b200_mixed* m, int frames, const uint8_t* code_id, const int8_t* d_llr,
                                   int max_trials, uint8_t* d_msg, int32_t* d_trials_left, int32_t* d_corrections,
                                   void* stream);
/* the set from packed tables (dvbs2b200_tables_build on rank 0, ONE broadcast of the concatenated blobs) */
int dvbs2b200_mixed_create_from_tables(dvbs2b200_mixed** m, int device, int n_codes, const void* const* blobs,
                                       const size_t* sizes);
uint64_t dvbs2b200_mixed_launch_count(const dvbs2b200_mixed* m);

/* ---- one code on several devices of ONE process ------------------------------------------------ */
/* A GNU Radio flowgraph is one process: a block that wants all GPUs of the node cannot use one rank per GPU.
 * dvbs2b200_multi holds one handle per listed device (the tables are built once on the host and copied to each
 * device: plain host-to-device copies, no collective is needed inside one process; a device may be listed more
 * than once).  dvbs2b200_multi_fec_decode is dvbs2b200_fec_decode with the batch cut into contiguous frame
 * ranges whose boundaries fall on multiples of 32 frames (a reference SIMD batch never straddles devices,
 * lib/ldpc_decoder/layered_decoder.hh:153); the ranges are staged and decoded concurrently (one library-owned
 * worker per device), the call returns when all are done.  Results are identical to the single-device call.
 * dvbs2b200_multi_code(m, i) exposes handle i for the other entry points (demap, bch, ...). */
typedef struct dvbs2b200_multi dvbs2b200_multi;
int dvbs2b200_multi_create(dvbs2b200_multi** m, const int* devices, int n_devices, int standard, int framesize,
                           int rate);
void dvbs2b200_multi_destroy(dvbs2b200_multi* m);
int dvbs2b200_multi_device_count(const dvbs2b200_multi* m);
dvbs2b200_code* dvbs2b200_multi_code(dvbs2b200_multi* m, int index);
int dvbs2b200_multi_shard(const dvbs2b200_multi* m, int frames, int index, int* first, int* count);
int dvbs2b200_multi_fec_decode(dvbs2b200_multi* m, int constellation, const float* iq, const float* n0,
                               const int8_t* llr, int frames, int max_trials, int term_group, uint8_t* msg,
                               int32_t* trials_left, int32_t* corrections);

/* ---- SNR estimate of the demapper block ------------------------------------------------------- */
/* dvbs2b200_estimate_snr <- the initial estimate in general_work  lib/xfecframe_demapper_cb_impl.cc:123-146
 *                           (QpskConstellation::estimate_snr      lib/qpsk.h:240-244, PSK hard/map lib/psk.hh:135-157)
 *                           and the refinement in handle_llr_pdu   lib/xfecframe_demapper_cb_impl.cc:252-318
 *                           (QpskConstellation::estimate_snr(llr)  lib/qpsk.h:267-281)
 * snr_lin[f] = sum |ref|^2 / sum |x - ref|^2 of frame f (linear Es/N0; the block keeps 10 log10 of it as
 * get_snr() and N0 = 1 / snr_lin).  llr_post == NULL: reference points by slicing the symbols; else from
 * the signs of the posterior LLRs [frames][n_ldpc] that the LDPC decoder publishes on "llr_pdu".
 * Floating-point sums: agreement with the reference is to tolerance (its VOLK sums have no fixed order). */
int dvbs2b200_estimate_snr(dvbs2b200_code* h, int constellation, const float* iq, const int8_t* llr_post,
                           int frames, float* snr_lin);
int dvbs2b200_estimate_snr_dev(dvbs2b200_code* h, int constellation, const float* d_iq,
                               const int8_t* d_llr_post, int frames, float* d_snr_lin, void* stream);

/* ---- fused chain: LLRs (or symbols) in, BBFRAME bytes out ---------------------------------- */
/* Runs [demap ->] LDPC (OM_MESSAGE) -> BCH with intermediates kept in device memory.
 * iq == NULL: start from llr.  msg [frames][kbch/8]; status arrays may be NULL. */
int dvbs2b200_fec_decode(dvbs2b200_code* h, int constellation, const float* iq, const float* n0,
                         const int8_t* llr, int frames, int max_trials, int term_group,
                         uint8_t* msg, int32_t* trials_left, int32_t* corrections);
int dvbs2b200_fec_decode_dev(dvbs2b200_code* h, int constellation, const float* d_iq,
                             const float* d_n0, const int8_t* d_llr, int frames, int max_trials,
                             int term_group, uint8_t* d_msg, int32_t* d_trials_left,
                             int32_t* d_corrections, void* stream);

/* ---- BB layer: descrambler + deheader, BBFRAMEs -> MPEG-TS bytes ------------------------------ */
/* dvbs2b200_bb_descramble  <- bbdescrambler_bb_impl::work             lib/bbdescrambler_bb_impl.cc:67-82
 * dvbs2b200_bb_deheader    <- bbdeheader_bb_impl::general_work        lib/bbdeheader_bb_impl.cc:144-261
 *                             (+ parse_bbheader / check_crc8          lib/bbdeheader_bb_impl.cc:76-142)
 * dvbs2b200_bb_counters_get<- get_packet_count ... get_bbframe_gap_count  lib/bbdeheader_bb_impl.h:89-93
 * dvbs2b200_fec_decode_ts  <- ldpc_decoder_bb -> bch_decoder_bb -> bbdescrambler_bb -> bbdeheader_bb chained
 *
 * in/out [frames][kbch/8] BBFRAMEs.  The deheader is a stream: whether it is synchronised and the bytes of
 * a TS packet cut by a BBFRAME boundary carry over from call to call in the handle (device memory), exactly
 * the reference block's members; dvbs2b200_bb_reset() puts it back to the just-constructed state.
 * ts receives whole 188-byte packets (sync byte restored, transport-error indicator set on a CRC-8 failure);
 * dvbs2b200_bb_ts_capacity(frames) bytes always suffice and are required: a smaller ts_cap is DVBS2B200_EINVAL and
 * nothing is consumed (a clipped batch would lose packets while the deheader state moves past them).  scrambled != 0: the input is BCH output and is
 * descrambled on the fly (bbdescrambler_bb fused in); 0: already descrambled, as the reference block expects.
 * One deviation: where the reference's unsigned arithmetic wraps and reads past the BBFRAME (re-synchronising
 * on a header with syncd == dfl, lib/bbdeheader_bb_impl.cc:203-209), that frame yields no packet here. */
typedef struct {
    uint64_t packets;  /* get_packet_count()       */
    uint64_t errors;   /* get_error_count(): packets whose CRC-8 failed */
    uint64_t bbframes; /* get_bbframe_count()      */
    uint64_t dropped;  /* get_bbframe_drop_count() */
    uint64_t gaps;     /* get_bbframe_gap_count()  */
} dvbs2b200_bb_counters;

int dvbs2b200_bb_descramble(dvbs2b200_code* h, const uint8_t* in, int frames, uint8_t* out);
int dvbs2b200_bb_descramble_dev(dvbs2b200_code* h, const uint8_t* d_in, int frames, uint8_t* d_out, void* stream);
size_t dvbs2b200_bb_ts_capacity(const dvbs2b200_code* h, int frames);
int dvbs2b200_bb_deheader(dvbs2b200_code* h, const uint8_t* bbframes, int frames, int scrambled, uint8_t* ts,
                          size_t ts_cap, size_t* ts_bytes);
/* device variant: the byte count of the call is read back with dvbs2b200_bb_produced_dev after the stream
 * has been synchronised (it lives in the handle's device state).  d_ts must be 4-byte aligned (packets are
 * written as 32-bit words); d_bbframes may have any alignment. */
int dvbs2b200_bb_deheader_dev(dvbs2b200_code* h, const uint8_t* d_bbframes, int frames, int scrambled,
                              uint8_t* d_ts, size_t ts_cap, void* stream);
int dvbs2b200_bb_produced_dev(dvbs2b200_code* h, void* stream, size_t* ts_bytes);
int dvbs2b200_bb_reset(dvbs2b200_code* h);
int dvbs2b200_bb_counters_get(dvbs2b200_code* h, dvbs2b200_bb_counters* c);

/* [demap ->] LDPC -> BCH -> descrambler -> deheader with every intermediate in device memory: soft input
 * in, TS packets out.  Arguments as dvbs2b200_fec_decode; only TS bytes and the status words travel back. */
int dvbs2b200_fec_decode_ts(dvbs2b200_code* h, int constellation, const float* iq, const float* n0,
                            const int8_t* llr, int frames, int max_trials, int term_group, uint8_t* ts,
                            size_t ts_cap, size_t* ts_bytes, int32_t* trials_left, int32_t* corrections);
int dvbs2b200_fec_decode_ts_dev(dvbs2b200_code* h, int constellation, const float* d_iq, const float* d_n0,
                                const int8_t* d_llr, int frames, int max_trials, int term_group,
                                uint8_t* d_ts, size_t ts_cap, int32_t* d_trials_left,
                                int32_t* d_corrections, void* stream);

/* ---- PL descrambler + pilot-segment de-rotation: PLFRAME payloads -> XFECFRAMEs ------------------------------ */
/* dvbs2b200_pl_descramble_derotate <- the output stage of plsync_cc_impl::handle_payload  lib/plsync_cc_impl.cc:639-802
 *                                     (pl_descrambler::descramble                          lib/pl_descrambler.cc:100-104,
 *                                      the per-16-slot-segment phase correction            lib/plsync_cc_impl.cc:732-776)
 * dvbs2b200_pl_create              <- pl_descrambler::pl_descrambler / compute_descrambling_sequence
 *                                                                                           lib/pl_descrambler.cc:15-98
 * the last data-parallel stage upstream of the demapper.  Frame detection, PLSC decoding and the
 * phase / frequency ESTIMATES stay where they are (feedback-loop PHY, out of scope); this entry point takes their
 * results per frame and does the per-symbol work:
 *   payload    [frames][payload_len][2] float, payload_len = n_slots * 90 + (has_pilots ? ((n_slots - 1) / 16) * 36 : 0):
 *              the symbols after the PLHEADER, still scrambled, pilot blocks included (dvbs2b200_pl_payload_len)
 *   info       [frames] dvbs2b200_pl_frame: PLHEADER phase estimate, fine frequency offset (cycles per symbol, applied
 *              only when coarse_corrected), the phase estimates of the frame's pilot blocks
 *   xfecframe  [frames][n_slots * 90][2] float: descrambled, pilots removed, de-rotated -- the demapper's input.
 * The reference de-rotates with VOLK's serial rotator (phase *= increment per sample, float); here each symbol's
 * phase is evaluated in closed form, so agreement is to float tolerance (about 1e-5 relative), not bit-exact.
 * Symbol buffers of the _dev variant must be 16-byte aligned. */
typedef struct dvbs2b200_pl dvbs2b200_pl;
typedef struct {
    float plheader_phase;  /* plframe_info_t::plheader_phase */
    float fine_foffset;    /* plframe_info_t::fine_foffset, normalised */
    int coarse_corrected;  /* plframe_info_t::coarse_corrected */
    int reserved;
    float pilot_phase[22]; /* freq_sync::get_pilot_phase(b), b < n_pilots (MAX_PILOT_BLKS = 22) */
} dvbs2b200_pl_frame;
int dvbs2b200_pl_create(dvbs2b200_pl** h, int device, int gold_code);
void dvbs2b200_pl_destroy(dvbs2b200_pl* h);
int dvbs2b200_pl_payload_len(int n_slots, int has_pilots);
/* the scrambling codes R_n (0..3) of the first n payload symbols of a Gold code (host only; descrambling factor
 * {1, -j, -1, +j}[R_n]) */
int dvbs2b200_pl_scrambling_codes(int gold_code, uint8_t* rn, int n);
int dvbs2b200_pl_descramble_derotate(dvbs2b200_pl* h, const float* payload, int frames, int n_slots, int has_pilots,
                                     const dvbs2b200_pl_frame* info, float* xfecframe);
int dvbs2b200_pl_descramble_derotate_dev(dvbs2b200_pl* h, const float* d_payload, int frames, int n_slots,
                                         int has_pilots, const dvbs2b200_pl_frame* d_info, float* d_xfecframe,
                                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DVBS2_B200_H */
