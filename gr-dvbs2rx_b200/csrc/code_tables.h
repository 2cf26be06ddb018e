// code_tables.h -- host-side construction of the packed code tables ("blob") that the
// sm_100a kernels consume: LDPC circulants + serial-order schedule, GF(2^m) log/antilog.
//
// What it replaces in the reference: LDPCDecoder::init (lib/ldpc_decoder/layered_decoder.hh:
// 101-142), the LDPC<TABLE> iterator (lib/ldpc_decoder/ldpc.hh:27-88), get_fec_info
// (lib/fec_params.cc:16-344), galois_field / bch_codec construction (lib/gf.cc:19-67,
// lib/bch.cc:36-113).  The blob is position independent so rank 0 can build it once and
// broadcast it to the other GPUs.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace dvbs2b200 {

struct LdpcTableDef {
    const char* name;
    int N, K, q, n_circ, links_total, links_max_cn;
    const uint32_t* circ; // layer << 17 | group << 9 | shift
};
struct ModcodDef {
    int framesize, rate, standard, table, kbch, nbch, t;
};

int num_tables();
const LdpcTableDef* table_def(int table);
const ModcodDef* find_modcod(int standard, int framesize, int rate);

// ---- packed blob ---------------------------------------------------------------------------
constexpr uint32_t kBlobMagic = 0x32425344u; // "DSB2"
constexpr uint32_t kBlobVersion = 21;

// One per layer, 8 bytes, lives in shared memory.
struct LayerRec {
    uint32_t edge_begin; // first circulant of the layer in edges[]
    uint16_t cnt;        // data links per check node in this layer
    uint16_t conflict;   // number of shared links (circulants whose 360-bit group carries another one of this layer)
};
// One per circulant, 8 bytes, lives in shared memory.  Check-node pair p (nodes p and p + 180) reads, through
// circulant (group, shift = a' + 180 ra), the halfword at  hi + 2p - (p >= a' ? 360 : 0),  hi = group*360 + 360 - 2a';
// node p is the LOW byte of that halfword iff (ra ^ (p < a')) == 0.  With  t = w0 - p,  w0 = (a' - 1) + ra * 2^31:
//     bit 30 of t = (p >= a')              ("ge": the rotation wrapped)
//     bit 31 of t = ge ^ ra                ("g2": node p is the low byte)
// so two multiply-high instructions give both, and the PRMT selectors depend on g2 alone (ldpc_core.cuh).
struct alignas(8) EdgeRec { // one 64-bit shared-memory load
    uint32_t w0;
    int32_t hi;
};
// One per schedule step of an iteration, 8 bytes, lives in shared memory.  A conflict-free layer is one "pair"
// step (count == 0: check-node pair p = thread); a layer whose circulants share a 360-bit group is one
// "split" step (below).
struct StepRec {
    uint8_t layer;
    uint8_t run_len;   // chain-form split step: delta
    uint16_t count;    // split step: depth (levels of the serial order)
    uint32_t work_off; // low 24 bits: offset into work[]; high bits: kStep* flags
};
// A block barrier is needed before a step only if it touches a 360-bit group (or parity bits) that another
// thread wrote since the last barrier; consecutive conflict-free layers over disjoint groups run without one
// (parity links are thread private there).
constexpr uint32_t kStepBarrierBefore = 1u << 24;
constexpr uint32_t kStepChainOutLink1 = 1u << 29; // chain form: the second shared link carries the bit forward
constexpr uint32_t kStepSplit = 1u << 30;         // conflict layer (always set when count != 0)
constexpr uint32_t kStepChain = 1u << 31;         // split step whose shared links form independent chains (below)
constexpr uint32_t kStepOffMask = (1u << 24) - 1;
// Split step.  Only the links into a 360-bit group that carries two or more circulants of the layer ("shared"
// links, LayerRec::conflict of them, sorted last) are order sensitive: every other bit of the layer is touched
// by exactly one check node.  Thread p keeps the pair mapping (check nodes p and p+180): it evaluates the
// private links of both nodes in s16x2 like a conflict-free layer, then the shared links follow the serial order
// of the reference (lib/ldpc_decoder/layered_decoder.hh:50-79), then the private links are updated.
// A node's level = 1 + the highest level among the earlier nodes that touch one of its bits; nodes of one level are
// independent, and levels rise with j (the predecessors of node j are those of node j-1 moved up by one, plus
// possibly one more), so a level is a range of j.  StepRec::count = depth; work[] holds level[j] (1-based) for
// j = 0..359, then first_node[l] for l = 0..depth+1 (first_node[0] unused, first_node[depth+1] = 360).
// Level form: the nodes of a level are handed out to the threads of the CTA (the node's private minimum, sign parity
// and shared-link operands wait in shared memory, put there by the thread that owns the node), one block barrier per
// level; the v->c values the nodes saw go back through shared memory to the owners for phase 3.
// Chain form (one doubled group, the usual case): the serial order through the layer is `delta` independent
// chains of nodes j, j + delta, ...; one lane walks one chain and hands the updated bit to the next node in a
// register.  delta and the forwarding link ride in the step record.
constexpr int kMaxSharedLinks = 12; // largest over the 57 tables (DVB-S2 8/9 short)

struct BlobHeader {
    uint32_t magic, version, total_bytes, reserved0;
    int32_t table, standard, framesize, rate;
    int32_t N, K, R, q;
    int32_t n_circ, links_total, max_cn_deg, max_cnt; // max_cnt = max data links per check
    int32_t kbch, nbch, t, gf_m;
    int32_t kldpc_out;  // bits emitted in OM_MESSAGE (= nbch, lib/ldpc_decoder_bb_impl.cc:98)
    int32_t msg_words;  // 32-bit words of compressed check-node state per check-node PAIR: 1 + ceil((max_cnt + 2) / 8)
    int32_t n_steps_total, n_conflict_layers;
    int32_t steps_per_iter, max_depth;
    int32_t uniform_cnt; // 1 if every layer has max_cnt data links per check node
    int32_t tmem_cols;   // unused (0)
    // section offsets from the start of the blob, all 16-byte aligned
    uint32_t smem_off, smem_bytes; // [LayerRec q][EdgeRec n_circ][StepRec steps]: TMA-staged
    uint32_t layer_off, edge_off;  // (inside the smem section)
    uint32_t step_off, order_off;  // StepRec[] (inside the smem section), uint16 work[] (global)
    uint32_t antilog_off, log_off; // uint16[2^m] each: alpha^i (i < 2^m-1), log(x)
    uint32_t bch_shorten;          // s = 2^m - 1 - nbch
    uint32_t split_steps;          // 1 (conflict layers are split steps)
    uint32_t chain_scratch;        // bytes of shared-memory scratch the split steps need (chain form: 360 x 8 B node records,
                                   // level form: 720 B per shared link + 360 B + the first_node table, ldpc_steps.cuh: LevelScratch)
    uint32_t level_calls;          // unused (0)
};
static_assert(sizeof(BlobHeader) % 16 == 0, "header must keep sections 16-byte aligned");

// Shared-memory layout of a frame's posteriors ("pair-interleaved"): element m of 360-bit group g
// sits at byte g*360 + 2*(m % 180) + m/180, parity bit c at K + 2*(c % (R/2)) + c/(R/2), so the
// two check nodes p and p+180 of a layer always find their operands in ONE 16-bit word.
inline EdgeRec pack_edge(int group, int shift)
{
    const int ap = shift % 180, ra = shift / 180;
    EdgeRec e;
    e.w0 = (uint32_t)(ap - 1) + ((uint32_t)ra << 31);
    e.hi = group * 360 + 360 - 2 * ap;
    return e;
}
// (a', ra, group base) back from a record, for the code that walks single check nodes
#if defined(__CUDACC__)
__host__ __device__
#endif
inline void unpack_edge(const EdgeRec& e, int& ap, int& ra, int& gbase)
{
    ra = (int)((e.w0 + 1u) >> 31);
    ap = (int)((e.w0 + 1u) & 0x7fffffffu);
    gbase = e.hi - 360 + 2 * ap;
}

// Builds the blob for a (standard, framesize, rate).  Returns false and sets err on failure.
bool build_blob(int standard, int framesize, int rate, std::vector<uint8_t>& blob, std::string& err);
bool validate_blob(const void* blob, size_t size, std::string& err);

// serial-order wavefront schedule of one table (also used for dvbs2b200_schedule_stats)
struct Schedule {
    std::vector<LayerRec> layers;
    std::vector<EdgeRec> edges;
    std::vector<StepRec> steps; // one iteration, in execution order
    std::vector<uint16_t> order; // work[]: check-node indices of the conflict steps
    int max_cnt = 0, min_cnt = 1 << 30, steps_per_iter = 0, max_depth = 0, conflict_layers = 0, barriers_per_iter = 0;
    bool has_chain = false; // at least one of them in chain form
    int max_level_shared = 0; // most shared links in a level-form split step
};
void build_schedule(const LdpcTableDef& def, Schedule& s);

// ---- GF(2^m) / BCH host helpers --------------------------------------------------------------
uint32_t bch_prim_poly(int framesize); // lib/bch_decoder_bb_impl.cc:58-63
void gf_tables(uint32_t prim_poly, std::vector<uint16_t>& antilog, std::vector<uint16_t>& log);
// generator polynomial, g[i] = coefficient of x^i (lib/bch.cc:36-62)
std::vector<uint8_t> bch_genpoly(uint32_t prim_poly, int t);

} // namespace dvbs2b200
