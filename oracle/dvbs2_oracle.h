/*
 * dvbs2_oracle.h -- CPU restatement of the gr-dvbs2rx FEC decode path (TEST INFRASTRUCTURE).
 *
 * This is the checker, not the product.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product
 * (gr-dvbs2rx_b200/) never links or calls anything in oracle/.
 *
 * Parity pinning: the LDPC and BCH restatements are checked byte-for-byte against the
 * reference's own translation units compiled into oracle/_ref (tests/test_oracle_vs_ref.py,
 * container only) and against the committed fixtures in tests/golden/ produced by that
 * compiled reference (tools/gen_golden.py).  BCH/GF are additionally checked against the
 * reference's known-answer tests (lib/qa_gf.cc:204-283, lib/qa_bch.cc:90-410).
 * The QPSK demap depends on VOLK (absent, unvendored; version unpinned by the reference):
 * its restatement is pinned only by lib/qa_qpsk.cc:67-79 -> "parity unpinned" beyond that.
 * The 8PSK demap is checked against lib/psk.hh compiled in oracle/_ref with
 * -ffp-contract=off (no reference test pins it).
 */
#ifndef DVBS2_ORACLE_H
#define DVBS2_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- code lookup --------------------------------------------------------------------- */
int orc_num_tables(void);
const char* orc_table_name(int table);
/* (standard, framesize, rate) ordinals of include/gnuradio/dvbs2rx/dvb_config.h:15-121.
 * Returns the table index or -1; fills kbch/nbch/t (lib/fec_params.cc:16-344) when non-NULL. */
int orc_lookup(int standard, int framesize, int rate, int* kbch, int* nbch, int* t);
int orc_table_n(int table);
int orc_table_k(int table);

/* ---- LDPC (lib/ldpc_decoder/layered_decoder.hh, algorithms.hh:151-207) ---------------- */
typedef struct orc_ldpc orc_ldpc;
orc_ldpc* orc_ldpc_create(int table);
void orc_ldpc_destroy(orc_ldpc*);
/* code: [lanes][N] int8 LLRs, overwritten with posteriors.  All lanes iterate together
 * (layered_decoder.hh:143-160).  Returns trials left (>= 0) or -1. */
int orc_ldpc_decode(orc_ldpc*, int8_t* code, int lanes, int trials);
/* syndrome check only (layered_decoder.hh:32-49): 1 if any lane has an unsatisfied check */
int orc_ldpc_bad(orc_ldpc*, const int8_t* code, int lanes);
/* systematic IRA encoder of the standard (not in the reference; validated by orc_ldpc_bad):
 * bits are one per byte (0/1); cw[0..K) = msg, cw[K..N) = parity. */
void orc_ldpc_encode(orc_ldpc*, const uint8_t* msg_bits, uint8_t* cw_bits);
/* hard decision + MSB-first packing (lib/ldpc_decoder_bb_impl.cc:432-442) */
void orc_pack_hard(const int8_t* llr, int nbits, uint8_t* out);

/* ---- BCH (lib/bch.cc, lib/gf.cc, lib/gf_util.h) --------------------------------------- */
typedef struct orc_bch orc_bch;
/* prim_poly includes the x^m term, e.g. 0x1002D for x^16+x^5+x^3+x^2+1
 * (lib/bch_decoder_bb_impl.cc:58-63).  n = 0 -> 2^m - 1. */
orc_bch* orc_bch_create(uint32_t prim_poly, int t, int n);
void orc_bch_destroy(orc_bch*);
int orc_bch_n(const orc_bch*);
int orc_bch_k(const orc_bch*);
/* generator polynomial coefficients, g[i] = coefficient of x^i, returns degree */
int orc_bch_genpoly(const orc_bch*, uint8_t* g, int cap);
/* minimal polynomial of alpha^i as a bit mask (lib/gf.cc get_min_poly) */
uint32_t orc_gf_min_poly(const orc_bch*, uint32_t i);
uint32_t orc_gf_alpha(const orc_bch*, uint32_t i);
void orc_bch_encode(const orc_bch*, const uint8_t* msg, uint8_t* cw);   /* lib/bch.cc:157-173 */
/* lib/bch.cc:467-487: returns #corrected, 0, or -1 (flips found so far still applied) */
int orc_bch_decode(const orc_bch*, const uint8_t* cw, uint8_t* msg);
/* pieces exposed for the reference's known-answer tests */
int orc_bch_syndrome(const orc_bch*, const uint8_t* cw, uint32_t* synd /*[2t]*/); /* 0 if clean */
int orc_bch_err_loc_poly(const orc_bch*, const uint32_t* synd, uint32_t* sigma /*[t+2]*/);
int orc_bch_err_loc_numbers(const orc_bch*, const uint32_t* sigma, int deg, uint32_t* numbers);

/* ---- soft demapper (lib/qpsk.h:208-214, lib/psk.hh:143-150, ----------------------------
 *      lib/xfecframe_demapper_cb_impl.cc:48-69,152-176) */
void orc_demap_qpsk(const float* iq, int n_syms, float n0, int8_t* llr);
/* soft demap + 3-column deinterleave; rate picks the column order */
void orc_demap_8psk(const float* iq, int n_syms, float n0, int rate, int8_t* llr);

/* ---- SNR estimates (lib/qpsk.h:41-65,240-281; lib/xfecframe_demapper_cb_impl.cc:128-142,267-302):
 * linear Es/N0 of one frame; llr == NULL slices the symbols, else the posterior LLR signs give the
 * reference points.  Tolerance-checked only (VOLK summation order is unspecified). */
float orc_snr_qpsk(const float* iq, int n_syms, const int8_t* llr);
float orc_snr_8psk(const float* iq, int n_syms, const int8_t* llr, int rate);

/* ---- BB layer (lib/bbdescrambler_bb_impl.cc:51-82, lib/bbdeheader_bb_impl.cc:76-261) -------
 * Pinned against the reference's two translation units compiled unmodified over a 60-line
 * gr::block shim (oracle/ref_bb_harness.cc, oracle/shim/gnuradio/block.h) and against the
 * reference's QA cases (python/dvbs2rx/qa_bbdeheader_bb.py) restated in tests/. */
void orc_bb_prbs(uint8_t* seq, int nbytes);
void orc_bb_descramble(const uint8_t* in, int frames, int kbch_bytes, uint8_t* out);
uint8_t orc_crc8(const uint8_t* in, int size); /* 0 <=> check passes */
typedef struct orc_bbdeheader orc_bbdeheader;
orc_bbdeheader* orc_bbdeheader_create(int kbch);
void orc_bbdeheader_destroy(orc_bbdeheader*);
/* `frames` descrambled BBFRAMEs of kbch/8 bytes -> TS bytes in out (capacity frames * (kbch/8 + 188));
 * state (sync, partial packet, counters) persists across calls.  Returns the bytes produced. */
long orc_bbdeheader_work(orc_bbdeheader*, const uint8_t* in, int frames, uint8_t* out);
/* packet, error, bbframe, bbframe_drop, bbframe_gap counts */
void orc_bbdeheader_counters(const orc_bbdeheader*, uint64_t* out5);

/* ---- PL descrambler + pilot-segment de-rotation (lib/pl_descrambler.cc:36-98, lib/plsync_cc_impl.cc:639-802) ----
 * rn[i] in 0..3: scrambling code of payload symbol i (pilot blocks included); descrambling factor {1, -j, -1, +j}[rn]. */
void orc_pl_rn(int gold_code, uint8_t* rn, int n);
/* payload [n_slots * 90 + n_pilots * 36][2] floats -> out [n_slots * 90][2]; n_pilots = (n_slots - 1) / 16 with pilots.
 * De-rotation restates VOLK's generic rotator (serial float recurrence): compare to tolerance. */
void orc_pl_payload(const float* payload, int n_slots, int has_pilots, const uint8_t* rn, float plheader_phase, float fine_foffset,
                    int coarse_corrected, const float* pilot_phase, float* out);

#ifdef __cplusplus
}
#endif
#endif
