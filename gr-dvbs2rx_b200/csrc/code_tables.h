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
constexpr uint32_t kBlobVersion = 13;

// One per layer, 8 bytes, lives in shared memory.
struct LayerRec {
    uint32_t edge_begin; // first circulant of the layer in edges[]
    uint16_t cnt;        // data links per check node in this layer
    uint16_t conflict;   // number of shared links (circulants whose 360-bit group carries another one of this layer)
};
// One per circulant, 8 bytes, lives in shared memory (see pack_edge).
struct EdgeRec {
    uint32_t e0; // hi | a' << 16   (hi <= 65160 fits 16 bits)
    uint32_t e1; // PRMT selectors for p >= a': unpack | pack << 16
};
// One per schedule step of an iteration, 8 bytes, lives in shared memory.  A conflict-free layer
// is one step (count == 0: check-node pair p = thread); a conflict layer is a sequence of wavefront
// steps, each a list of `count` check-node indices j in work[work_off ...].
struct StepRec {
    uint8_t layer;
    uint8_t run_len;   // > 0 on the first step of a link-parallel run: steps in the run
    uint16_t count;
    uint32_t work_off; // low 24 bits: offset into work[]; high bits: kStep* flags
};
// Barriers.  A block barrier is needed before a step only if it touches a 360-bit group (or parity
// bits) that another thread wrote since the last barrier; consecutive conflict-free layers over
// disjoint groups run without one (parity links are thread private there).
// Wavefront steps come in three classes, chosen by instruction cost (code_tables.cc):
//  * wide levels (> 32 check nodes): one check node per thread on all warps;
//  * narrow levels: one check node per lane on warp 0 alone;
//  * narrow levels of high-degree codes / long chains: "link parallel", one lane per (check node,
//    link) with G = degree rounded up to 8/16/32 lanes per node, minima by warp REDUX, on the first
//    nwarps warps.
// Consecutive narrow levels of a layer form a run: the warps that take part order the levels with
// __syncwarp() / a named barrier, the other warps skip the whole run.
constexpr uint32_t kStepBarrierBefore = 1u << 24;
constexpr uint32_t kStepRun = 1u << 25;          // part of a run executed by a subset of the warps
constexpr uint32_t kStepWarpsShift = 26;         // 3 bits: warps taking part in the run
constexpr uint32_t kStepLinkParallel = 1u << 29; // run class: link parallel (else one node per lane)
constexpr uint32_t kStepSplit = 1u << 30;        // whole conflict layer in one step, see below
constexpr uint32_t kStepChain = 1u << 31;        // split step whose shared links form independent chains (below);
                                                 // StepRec::run_len = delta, kStepChainOutLink1 = forwarding link
constexpr uint32_t kStepChainOutLink1 = 1u << 29; // (shares its bit with kStepLinkParallel: never both)
constexpr uint32_t kStepOffMask = (1u << 24) - 1;
// Split step (the default for conflict layers).  Only the links into a 360-bit group that carries two
// or more circulants of the layer ("shared" links, LayerRec::conflict of them, sorted last) are order
// sensitive: every other bit of the layer is touched by exactly one check node.  So thread p keeps the
// pair mapping (check nodes p and p+180): it evaluates the private links of both nodes in s16x2 like a
// conflict-free layer, then the block walks the levels of the serial order and a node merges only its
// shared links into its partial minima / sign product and updates those bits, then the private links
// are updated in s16x2 again.  The dependent chain through a layer shrinks from a whole check-node
// update per level to a 2-link merge; StepRec::count = depth, work[] holds level[j] (1-based).
// Chain form of a split step (one doubled group, the usual case): the serial order through the layer is
// `delta` independent chains of nodes j, j + delta, ...; one lane walks one chain and hands the updated bit
// to the next node in a register, so the dependent path per node is a 2-link merge with no barrier and no
// shared-memory round trip.  The nodes' partial results travel through a 360 x 8-byte scratch in shared
// memory.  delta and the forwarding link ride in the step record.
constexpr int kMaxSharedLinks = 12; // largest over the 57 tables (DVB-S2 8/9 short)

// Tensor memory as scratch for the check-node state of order-sensitive layers.  Those layers are a
// dependent chain of short steps, and the ~700-cycle L2 round trip for the state word sits on that
// chain; TMEM (256 KB per SM, otherwise idle here) answers in ~30 cycles.  Each CTA owns kTmemCols
// columns (3 CTAs x 128 of the SM's 512); a step gets one column per 4 warps and pass, the host
// assigns them (narrow runs first) until the budget is spent -- the rest stays in L2.
constexpr int kTmemCols = 128;
constexpr uint8_t kNoTmem = 0xff;

struct BlobHeader {
    uint32_t magic, version, total_bytes, reserved0;
    int32_t table, standard, framesize, rate;
    int32_t N, K, R, q;
    int32_t n_circ, links_total, max_cn_deg, max_cnt; // max_cnt = max data links per check
    int32_t kbch, nbch, t, gf_m;
    int32_t kldpc_out;  // bits emitted in OM_MESSAGE (= nbch, lib/ldpc_decoder_bb_impl.cc:98)
    int32_t msg_words;  // 32-bit words of compressed message state per check node (1 or 2)
    int32_t n_steps_total, n_conflict_layers;
    int32_t steps_per_iter, max_depth;
    int32_t uniform_cnt; // 1 if every layer has max_cnt data links per check node
    int32_t tmem_cols;   // TMEM columns a CTA allocates (0 or kTmemCols)
    // section offsets from the start of the blob, all 16-byte aligned
    uint32_t smem_off, smem_bytes; // [LayerRec q][EdgeRec n_circ][StepRec steps][uint8 tmem column steps]: TMA-staged
    uint32_t layer_off, edge_off;  // (inside the smem section)
    uint32_t step_off, order_off;  // StepRec[] (inside the smem section), uint16 work[] (global)
    uint32_t antilog_off, log_off; // uint16[2^m] each: alpha^i (i < 2^m-1), log(x)
    uint32_t bch_shorten;          // s = 2^m - 1 - nbch
    uint32_t split_steps;          // 1: conflict layers are split steps (kernels of the _split build), 0: wavefront steps
    uint32_t chain_scratch;        // 1: some split step is in chain form: the kernel needs the 360 x 8 B node scratch
    uint32_t level_calls;          // 1: split build, level-form steps through the out-of-line compile-time-count copies
};
static_assert(sizeof(BlobHeader) % 16 == 0, "header must keep sections 16-byte aligned");

// Shared-memory layout of a frame's posteriors ("pair-interleaved"): element m of 360-bit group g
// sits at byte g*360 + 2*(m % 180) + m/180, parity bit c at K + 2*(c % (R/2)) + c/(R/2), so the
// two check nodes p and p+180 of a layer always find their operands in ONE 16-bit word.
// Check-node pair p reads, through circulant (group, shift = a' + 180*ra), the halfword at
//   hi + 2p - (p >= a' ? 360 : 0),   hi = group*360 + 360 - 2a';
// node p is the low byte iff (ra ^ (p < a')) == 0; ra is bit 0 of the unpack selector.
inline EdgeRec pack_edge(int group, int shift)
{
    const int ap = shift % 180, ra = shift / 180;
    EdgeRec e;
    e.e0 = (uint32_t)(group * 360 + 360 - 2 * ap) | ((uint32_t)ap << 16);
    // PRMT selectors when p >= a' (r = ra); the kernel XORs 0x1111 / 0x0022 when p < a'
    const uint32_t unpack = ra ? 0x8091u : 0x9180u; // bytes -> sign-extended s16x2 [node p | node p+180]
    const uint32_t pack = ra ? 0x4402u : 0x4420u;   // s16x2 -> two bytes in memory order
    e.e1 = unpack | (pack << 16);
    return e;
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
    std::vector<uint8_t> tcol;   // per step: first TMEM column of its state, or kNoTmem
    int max_cnt = 0, min_cnt = 1 << 30, steps_per_iter = 0, max_depth = 0, conflict_layers = 0, barriers_per_iter = 0;
    bool split = false;     // conflict layers emitted as split steps
    bool has_chain = false; // at least one of them in chain form
};
// split: 1 / 0 force the form of the conflict layers, -1 lets choose_split() decide (env DVBS2B200_SPLIT overrides)
void build_schedule(const LdpcTableDef& def, Schedule& s, bool use_tmem = true, int split = -1);
bool choose_split(const LdpcTableDef& def);
bool choose_level_calls(const LdpcTableDef& def);

// ---- GF(2^m) / BCH host helpers --------------------------------------------------------------
uint32_t bch_prim_poly(int framesize); // lib/bch_decoder_bb_impl.cc:58-63
void gf_tables(uint32_t prim_poly, std::vector<uint16_t>& antilog, std::vector<uint16_t>& log);
// generator polynomial, g[i] = coefficient of x^i (lib/bch.cc:36-62)
std::vector<uint8_t> bch_genpoly(uint32_t prim_poly, int t);

} // namespace dvbs2b200
