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
constexpr uint32_t kBlobVersion = 3;

// One per layer, 16 bytes, lives in shared memory.
struct LayerRec {
    uint32_t edge_begin; // first circulant of the layer in edges[]
    uint16_t cnt;        // data links per check node in this layer
    uint16_t n_steps;    // 1 = no intra-layer conflict (identity order), else wavefront steps
    uint32_t step_begin; // first entry in steps[] (conflict layers only)
    uint32_t order_begin; // first entry in order[] (conflict layers only)
};
// One per wavefront step of a conflict layer (global memory).
struct StepRec {
    uint16_t begin; // offset into the layer's 360-entry order[] slice
    uint16_t count; // check nodes in this step
};

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
    // section offsets from the start of the blob, all 16-byte aligned
    uint32_t smem_off, smem_bytes; // [LayerRec q][edges n_circ]: TMA-staged into shared memory
    uint32_t layer_off, edge_off;  // (inside the smem section)
    uint32_t step_off, order_off;  // StepRec[], uint16 order[]
    uint32_t antilog_off, log_off; // uint16[2^m] each: alpha^i (i < 2^m-1), log(x)
    uint32_t bch_shorten;          // s = 2^m - 1 - nbch
    uint32_t reserved1[5];
};
static_assert(sizeof(BlobHeader) % 16 == 0, "header must keep sections 16-byte aligned");

// edge word in edges[]: hi | shift << 17, hi = group*360 + 360 - shift.
// Check node j of the layer reads data bit  hi + j - (j >= shift ? 360 : 0).
inline uint32_t pack_edge(int group, int shift) { return (uint32_t)(group * 360 + 360 - shift) | ((uint32_t)shift << 17); }

// Builds the blob for a (standard, framesize, rate).  Returns false and sets err on failure.
bool build_blob(int standard, int framesize, int rate, std::vector<uint8_t>& blob, std::string& err);
bool validate_blob(const void* blob, size_t size, std::string& err);

// serial-order wavefront schedule of one table (also used for dvbs2b200_schedule_stats)
struct Schedule {
    std::vector<LayerRec> layers;
    std::vector<uint32_t> edges;
    std::vector<StepRec> steps;
    std::vector<uint16_t> order;
    int max_cnt = 0, steps_per_iter = 0, max_depth = 0, conflict_layers = 0;
};
void build_schedule(const LdpcTableDef& def, Schedule& s);

// ---- GF(2^m) / BCH host helpers --------------------------------------------------------------
uint32_t bch_prim_poly(int framesize); // lib/bch_decoder_bb_impl.cc:58-63
void gf_tables(uint32_t prim_poly, std::vector<uint16_t>& antilog, std::vector<uint16_t>& log);
// generator polynomial, g[i] = coefficient of x^i (lib/bch.cc:36-62)
std::vector<uint8_t> bch_genpoly(uint32_t prim_poly, int t);

} // namespace dvbs2b200
