// code_tables.cc -- see code_tables.h.  Host only, no CUDA.
#include "code_tables.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace dvbs2b200 {

typedef LdpcTableDef Dvbs2LdpcTableDef;
typedef ModcodDef Dvbs2ModcodDef;
#include "dvbs2_code_tables.inc"

int num_tables() { return (int)(sizeof(kLdpcTables) / sizeof(kLdpcTables[0])); }
const LdpcTableDef* table_def(int table) { return (table >= 0 && table < num_tables()) ? &kLdpcTables[table] : nullptr; }

// Table choice of lib/ldpc_decoder_bb_impl.cc:104-307: rows with a standard are picked by
// `standard == STANDARD_DVBS2`, anything else takes the DVB-T2 table.
const ModcodDef* find_modcod(int standard, int framesize, int rate)
{
    const int n = (int)(sizeof(kModcods) / sizeof(kModcods[0]));
    for (int i = 0; i < n; ++i) {
        const ModcodDef& m = kModcods[i];
        if (m.framesize != framesize || m.rate != rate)
            continue;
        if (m.standard >= 0 && (m.standard == 0) != (standard == 0))
            continue;
        return &m;
    }
    return nullptr;
}

// ---- serial-order schedule -------------------------------------------------------------------
// The reference visits the 360 check nodes of a layer in ascending j
// (lib/ldpc_decoder/layered_decoder.hh:50-79).  Two circulants of the same 360-bit group in one
// layer make check nodes share data bits, and then that order is visible in the result.  For
// such layers we compute level[j] = 1 + max(level[j'] : j' < j touches one of j's bits); check
// nodes of equal level are independent, and running the levels in order reproduces the serial
// result bit for bit.
// Which form of the conflict layers is faster was measured per code on B200 (profiles/r01b_sweep_configs.jsonl):
// with the call-free split kernel the split steps win on all five BASELINE codes (1/2 normal 82 k -> 97 k,
// 3/4 normal 77 k -> 133 k, 3/5 normal 70 k -> 91 k, 2/3 short 100 k -> 155 k, 9/10 normal 70 k -> 80 k frames/s),
// so they are the default; the wavefront build stays in the library for A/B runs (DVBS2B200_SPLIT=0).
bool choose_split(const LdpcTableDef& def)
{
    (void)def;
    if (const char* env = getenv("DVBS2B200_SPLIT"))
        return atoi(env) != 0;
    return true;
}

// Split build, which kernel variant: with or without the out-of-line level-form copies (ldpc_kernel.cu).
// Measured on B200: the call-free variant wins wherever most conflict layers take the chain form (1/2 normal
// 84 k -> 97 k, 3/4 normal 118 k -> 132 k frames/s: no local-memory frame, DRAM traffic back to the compulsory
// bytes); short frames with a good share of three/four-link level-form layers keep the calls (2/3 short:
// 119 k without, 154 k with).  DVBS2B200_LEVEL_CALLS=0/1 overrides.
bool choose_level_calls(const LdpcTableDef& def)
{
    if (const char* env = getenv("DVBS2B200_LEVEL_CALLS"))
        return atoi(env) != 0;
    std::vector<std::vector<int>> groups(def.q);
    for (int c = 0; c < def.n_circ; ++c)
        groups[def.circ[c] >> 17].push_back((int)((def.circ[c] >> 9) & 0xff));
    int conflict = 0, level_form = 0, max_cnt = 0;
    for (auto& g : groups) {
        std::sort(g.begin(), g.end());
        max_cnt = std::max(max_cnt, (int)g.size());
        int shared = 0;
        for (size_t a = 0; a < g.size(); ++a)
            if ((a && g[a] == g[a - 1]) || (a + 1 < g.size() && g[a] == g[a + 1]))
                ++shared;
        conflict += shared > 0;
        level_form += shared > 2;
    }
    return def.N <= 16200 && max_cnt <= 13 && 4 * level_form >= conflict && level_form > 0;
}

void build_schedule(const LdpcTableDef& def, Schedule& s, bool use_tmem, int split_arg)
{
    const int q = def.q;
    s = Schedule();
    s.layers.resize(q);
    std::vector<std::vector<std::pair<int, int>>> per_layer(q);
    for (int c = 0; c < def.n_circ; ++c) {
        uint32_t w = def.circ[c];
        per_layer[w >> 17].push_back({ (int)((w >> 9) & 0xff), (int)(w & 0x1ff) });
    }
    const int ngroups = def.K / 360;
    std::vector<int> last(ngroups * 360);
    const bool split = split_arg < 0 ? choose_split(def) : split_arg != 0;
    s.split = split;
    const char* env_chain = getenv("DVBS2B200_CHAIN"); // diagnostics: 0 = level-by-level form for every split step
    const bool chain = !(env_chain && atoi(env_chain) == 0);
    for (int i = 0; i < q; ++i) {
        auto& circ = per_layer[i];
        LayerRec& L = s.layers[i];
        L.edge_begin = (uint32_t)s.edges.size();
        L.cnt = (uint16_t)circ.size();
        s.max_cnt = std::max(s.max_cnt, (int)circ.size());
        s.min_cnt = std::min(s.min_cnt, (int)circ.size());
        // Links whose 360-bit group carries another circulant of this layer are "shared": their bits are
        // touched by two or more check nodes of the layer.  They go last in the layer's link order (the
        // order of the links of a check node is immaterial to the result: minima, sign products and the
        // per-link saturating updates are all order independent), so the private links are 0 .. cnt-n_shared-1.
        std::vector<int> mult(ngroups, 0);
        for (auto& ga : circ)
            mult[ga.first]++;
        std::stable_partition(circ.begin(), circ.end(), [&](const std::pair<int, int>& ga) { return mult[ga.first] == 1; });
        int n_shared = 0;
        for (size_t a = 0; a < circ.size(); ++a) {
            s.edges.push_back(pack_edge(circ[a].first, circ[a].second));
            n_shared += mult[circ[a].first] > 1;
        }
        const bool conflict = n_shared > 0;
        L.conflict = (uint16_t)n_shared;
        if (!conflict) {
            StepRec st;
            st.layer = (uint8_t)i;
            st.run_len = 0;
            st.count = 0;
            st.work_off = 0;
            s.steps.push_back(st);
            s.steps_per_iter += 1;
            continue;
        }
        s.conflict_layers++;
        std::fill(last.begin(), last.end(), 0);
        std::vector<int> level(360);
        int depth = 0;
        for (int j = 0; j < 360; ++j) {
            int lv = 0;
            for (auto& ga : circ)
                lv = std::max(lv, last[ga.first * 360 + (j - ga.second + 360) % 360]);
            lv += 1;
            for (auto& ga : circ)
                last[ga.first * 360 + (j - ga.second + 360) % 360] = lv;
            level[j] = lv;
            depth = std::max(depth, lv);
        }
        s.steps_per_iter += depth;
        s.max_depth = std::max(s.max_depth, depth);
        if (split) { // n_shared <= kMaxSharedLinks holds for all 57 tables (asserted by tests/test_host_cpu.py)
            // One "split" step: private links of all 360 check nodes in parallel, the shared links level by
            // level (code_tables.h).  work[] holds level[j] for j = 0..359.
            StepRec st;
            st.layer = (uint8_t)i;
            st.run_len = 0;
            st.count = (uint16_t)depth;
            st.work_off = (uint32_t)s.order.size() | kStepSplit;
            for (int j = 0; j < 360; ++j)
                s.order.push_back((uint16_t)level[j]);
            // thread count of the named barrier in front of level l: the warps with nodes in level l-1 or l
            // (thread p = j % 180 owns node j; levels rise with j)
            std::vector<uint32_t> warps(depth + 2, 0);
            for (int j = 0; j < 360; ++j)
                warps[level[j]] |= 1u << ((j % 180) / 32);
            s.order.push_back(0);
            for (int lv = 1; lv <= depth; ++lv)
                s.order.push_back((uint16_t)(32 * __builtin_popcount(warps[lv - 1] | warps[lv])));
            // Chain form (code_tables.h): exactly two circulants (shifts a0 < a1) in one group.  With d = a1 - a0,
            // the bit node j reaches through link 0 is the bit node (j + d) mod 360 reaches through link 1, so
            // with delta = min(d, 360 - d) the layer is `delta` independent chains j, j + delta, j + 2 delta, ...
            // along which the updated bit is handed from a node's "out" link to the next node's "in" link.
            if (n_shared == 2 && chain) {
                const int a0 = circ[circ.size() - 2].second, a1 = circ[circ.size() - 1].second;
                const int d = ((a1 - a0) % 360 + 360) % 360;
                const int delta = std::min(d, 360 - d);
                const int out_link = d <= 180 ? 0 : 1; // which of the two shared links carries the bit forward
                if (delta >= 1 && (360 + delta - 1) / delta == depth) {
                    st.work_off |= kStepChain | (out_link ? kStepChainOutLink1 : 0u);
                    st.run_len = (uint8_t)delta; // <= 180
                    s.has_chain = true;
                }
            }
            s.steps.push_back(st);
            continue;
        }
        for (int lv = 1; lv <= depth; ++lv) {
            StepRec st;
            st.layer = (uint8_t)i;
            st.run_len = 0;
            st.work_off = (uint32_t)s.order.size();
            for (int j = 0; j < 360; ++j)
                if (level[j] == lv)
                    s.order.push_back((uint16_t)j);
            st.count = (uint16_t)(s.order.size() - st.work_off);
            s.steps.push_back(st);
        }
    }

    // ---- barrier placement ---------------------------------------------------------------------
    // dirty_all / dirty_sub: groups written since the last block barrier by steps every warp runs /
    // by link-parallel runs (a subset of the warps).  Group index ngroups stands for "parity bits touched across threads".
    std::vector<char> dirty_all(ngroups + 1, 0), dirty_sub(ngroups + 1, 0);
    auto clear = [&]() {
        std::fill(dirty_all.begin(), dirty_all.end(), 0);
        std::fill(dirty_sub.begin(), dirty_sub.end(), 0);
    };
    auto group_lanes = [&](int layer) { // lanes per check node in the link-parallel path
        const int deg = (int)per_layer[layer].size() + 2;
        return deg <= 8 ? 8 : deg <= 16 ? 16 : 32;
    };
    // class of a wavefront step: 0 wide (all warps), 1 narrow scalar on warp 0, 2 link parallel.
    // Instruction cost model (measured on B200): a scalar check-node update is ~55 warp instructions
    // per link for up to 32 nodes, a link-parallel pass ~150 per warp of (node, link) lanes.
    // tuning knobs (diagnostics): DVBS2B200_LP_LANES = link-parallel up to that many (node, link) lanes
    // per level instead of the cost model; DVBS2B200_W0_MAX = largest level run on warp 0 alone
    const char* env_lp = getenv("DVBS2B200_LP_LANES");
    const char* env_w0 = getenv("DVBS2B200_W0_MAX");
    const int lp_lanes = env_lp ? atoi(env_lp) : -1;
    const int w0_max = env_w0 ? atoi(env_w0) : 32;
    auto step_class = [&](const StepRec& st) {
        if (st.count == 0 || (st.work_off & kStepSplit))
            return 0;
        const int G = group_lanes(st.layer), deg = (int)per_layer[st.layer].size() + 2;
        const int lp_warps = ((int)st.count * G + 31) / 32;
        if (lp_lanes >= 0) {
            if ((int)st.count * G <= lp_lanes)
                return 2;
        } else if (st.count <= 32) {
            if (lp_warps <= 6 && 150 * lp_warps < 55 * deg)
                return 2;
            // high-degree codes (> 16 links): a scalar level is a ~1600-instruction dependent chain on one
            // warp; spreading it over the links pays up to 14 nodes per level even with several passes
            if (G == 32 && st.count <= 14)
                return 2;
        }
        return (int)st.count <= w0_max ? 1 : 0;
    };
    // runs of consecutive narrow steps of one layer and one class
    for (size_t k = 0; k < s.steps.size();) {
        const int cls = step_class(s.steps[k]);
        if (cls == 0) {
            ++k;
            continue;
        }
        size_t e = k;
        int lanes = 0;
        while (e < s.steps.size() && e - k < 255 && s.steps[e].layer == s.steps[k].layer && step_class(s.steps[e]) == cls) {
            lanes = std::max(lanes, (int)s.steps[e].count * group_lanes(s.steps[e].layer));
            ++e;
        }
        const int warps = cls == 2 ? std::min(6, (lanes + 31) / 32) : 1;
        s.steps[k].run_len = (uint8_t)(e - k);
        for (size_t t = k; t < e; ++t)
            s.steps[t].work_off |= kStepRun | ((uint32_t)warps << kStepWarpsShift) | (cls == 2 ? kStepLinkParallel : 0u);
        k = e;
    }
    for (size_t k = 0; k < s.steps.size(); ++k) {
        StepRec& st = s.steps[k];
        const bool is_split = (st.work_off & kStepSplit) != 0;
        // a split step keeps the pair mapping (thread p owns check nodes p and p+180): parity links stay private
        const bool conflict_layer = st.count != 0 && !is_split;
        const bool sub = (st.work_off & kStepRun) != 0; // runs on a subset of the warps
        const bool inside_run = sub && st.run_len == 0;          // ordered by the run's own barrier
        std::vector<int> groups;
        for (auto& ga : per_layer[st.layer])
            groups.push_back(ga.first);
        // conflict layers run single check nodes on arbitrary threads: their parity links are not
        // thread private, and the neighbouring layers' parity links collide with them
        const bool prev_conflict = k > 0 && s.steps[k - 1].count != 0 && !(s.steps[k - 1].work_off & kStepSplit);
        bool need = false;
        for (int g : groups)
            need |= dirty_all[g] || (!inside_run && dirty_sub[g]);
        if (conflict_layer || prev_conflict) // cross-thread parity access: every earlier step wrote parity
            need |= dirty_all[ngroups] || (!inside_run && dirty_sub[ngroups]);
        groups.push_back(ngroups); // every step writes parity bits
        if (k == 0)
            need = false; // the iteration starts behind a barrier
        if (inside_run)
            need = false; // (dirty_all was cleared by the barrier in front of the run)
        if (need) {
            st.work_off |= kStepBarrierBefore;
            s.barriers_per_iter++;
            clear();
        }
        if (is_split && !need) {
            // the named barriers of the levels are reused from one split step to the next: keep a block
            // barrier between them and whatever ran before
            st.work_off |= kStepBarrierBefore;
            s.barriers_per_iter++;
            clear();
        }
        for (int g : groups)
            (sub ? dirty_sub : dirty_all)[g] = 1;
    }

    // ---- tensor-memory columns for the state of wavefront steps (code_tables.h) --------------------
    s.tcol.assign(s.steps.size(), kNoTmem);
    if (use_tmem && s.max_cnt <= 13) { // one-word state only
        int next_col = 0;
        auto cols_needed = [&](const StepRec& st) {
            if (st.count == 0 || (st.work_off & (kStepLinkParallel | kStepSplit)))
                return 0;
            if (st.work_off & kStepRun)
                return 1;                          // warp 0, one pass
            return 2 * (((int)st.count + 191) / 192); // warps 0-3 and 4-5, per pass
        };
        // narrow runs first (they are the deepest chains); a run is placed whole or not at all
        for (size_t k = 0; k < s.steps.size(); ++k) {
            const StepRec& st = s.steps[k];
            if (!(st.work_off & kStepRun) || (st.work_off & kStepLinkParallel) || st.run_len == 0)
                continue;
            if (next_col + (int)st.run_len > kTmemCols)
                continue;
            for (int t = 0; t < (int)st.run_len; ++t)
                s.tcol[k + t] = (uint8_t)next_col++;
        }
        for (size_t k = 0; k < s.steps.size(); ++k) {
            const StepRec& st = s.steps[k];
            const int need = cols_needed(st);
            if (need == 0 || (st.work_off & kStepRun) || next_col + need > kTmemCols)
                continue;
            s.tcol[k] = (uint8_t)next_col;
            next_col += need;
        }
    }
}

// ---- GF(2^m) ---------------------------------------------------------------------------------
uint32_t bch_prim_poly(int framesize)
{
    // lib/bch_decoder_bb_impl.cc:58-63; framesize ordinals of dvb_config.h: SHORT 0, NORMAL 1, MEDIUM 2
    if (framesize == 1)
        return 0x1002D; // x^16 + x^5 + x^3 + x^2 + 1
    if (framesize == 0)
        return 0x402B; // x^14 + x^5 + x^3 + x + 1
    return 0x802D;     // x^15 + x^5 + x^3 + x^2 + 1
}

static int poly_m(uint32_t p)
{
    int m = 31;
    while (m > 0 && !(p >> m))
        --m;
    return m;
}

// lib/gf.cc:46-62: alpha^(i+1) = alpha^i * x mod p(x)
void gf_tables(uint32_t prim_poly, std::vector<uint16_t>& antilog, std::vector<uint16_t>& log)
{
    const int m = poly_m(prim_poly);
    const uint32_t nz = (1u << m) - 1;
    antilog.assign((size_t)1 << m, 0);
    log.assign((size_t)1 << m, 0);
    uint32_t v = 1;
    const uint32_t low = prim_poly ^ (1u << m);
    for (uint32_t i = 0; i < nz; ++i) {
        antilog[i] = (uint16_t)v;
        log[v] = (uint16_t)i;
        v = ((v << 1) & nz) ^ ((v >> (m - 1)) * low);
    }
    antilog[nz] = 1; // alpha^(2^m-1) = 1, lets kernels skip one wrap test
}

// lib/bch.cc:36-62: g(x) = lcm of the minimal polynomials of alpha^1, alpha^3, ..., alpha^(2t-1)
std::vector<uint8_t> bch_genpoly(uint32_t prim_poly, int t)
{
    std::vector<uint16_t> al, lg;
    gf_tables(prim_poly, al, lg);
    const int m = poly_m(prim_poly);
    const uint32_t nz = (1u << m) - 1;
    auto mul = [&](uint32_t a, uint32_t b) -> uint32_t {
        if (!a || !b)
            return 0;
        return al[(lg[a] + lg[b]) % nz];
    };
    std::vector<uint8_t> g(1, 1);
    std::vector<uint32_t> seen_exp;
    for (int i = 0; i < t; ++i) {
        uint32_t e = (uint32_t)(2 * i + 1) % nz;
        if (std::find(seen_exp.begin(), seen_exp.end(), e) != seen_exp.end())
            continue;
        // conjugacy class of alpha^e
        std::vector<uint32_t> cls;
        uint32_t x = e;
        do {
            cls.push_back(x);
            seen_exp.push_back(x);
            x = (uint32_t)(((uint64_t)x * 2) % nz);
        } while (x != e);
        // phi(x) = prod (x + alpha^c)
        std::vector<uint32_t> phi(1, 1);
        for (uint32_t c : cls) {
            uint32_t beta = al[c];
            std::vector<uint32_t> nx(phi.size() + 1, 0);
            for (size_t d = 0; d < phi.size(); ++d) {
                nx[d + 1] ^= phi[d];
                nx[d] ^= mul(phi[d], beta);
            }
            phi.swap(nx);
        }
        std::vector<uint8_t> prod(g.size() + phi.size() - 1, 0);
        for (size_t a = 0; a < g.size(); ++a)
            if (g[a])
                for (size_t b = 0; b < phi.size(); ++b)
                    if (phi[b]) // phi has 0/1 coefficients
                        prod[a + b] ^= 1;
        g.swap(prod);
    }
    return g;
}

// ---- blob ------------------------------------------------------------------------------------
static size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

bool build_blob(int standard, int framesize, int rate, std::vector<uint8_t>& blob, std::string& err)
{
    const ModcodDef* mc = find_modcod(standard, framesize, rate);
    if (!mc) {
        err = "no LDPC table for this (standard, framesize, rate)";
        return false;
    }
    const LdpcTableDef& def = kLdpcTables[mc->table];
    Schedule s;
    build_schedule(def, s);

    BlobHeader h;
    memset(&h, 0, sizeof(h));
    h.magic = kBlobMagic;
    h.version = kBlobVersion;
    h.table = mc->table;
    h.standard = standard;
    h.framesize = framesize;
    h.rate = rate;
    h.N = def.N;
    h.K = def.K;
    h.R = def.N - def.K;
    h.q = def.q;
    h.n_circ = def.n_circ;
    h.links_total = def.links_total;
    h.max_cn_deg = s.max_cnt + 2;
    h.max_cnt = s.max_cnt;
    h.kbch = mc->kbch;
    h.nbch = mc->nbch;
    h.t = mc->t;
    const uint32_t pp = bch_prim_poly(framesize);
    h.gf_m = poly_m(pp);
    h.kldpc_out = mc->nbch;
    // compressed check-node state: 6+6 bits of clamped minima, 5 bits argmin, 1 sign bit/link
    h.msg_words = (s.max_cnt <= 13) ? 1 : 2; // must agree with ldpc_wide_state()
    h.n_steps_total = (int32_t)s.steps.size();
    h.n_conflict_layers = s.conflict_layers;
    h.steps_per_iter = s.steps_per_iter;
    h.max_depth = s.max_depth;
    h.uniform_cnt = (s.min_cnt == s.max_cnt) ? 1 : 0;
    h.bch_shorten = ((1u << h.gf_m) - 1) - (uint32_t)mc->nbch;
    h.split_steps = s.split ? 1u : 0u;
    h.chain_scratch = s.has_chain ? 1u : 0u;
    h.level_calls = (s.split && choose_level_calls(def)) ? 1u : 0u;

    size_t off = sizeof(BlobHeader);
    h.smem_off = (uint32_t)off;
    h.layer_off = (uint32_t)off;
    off += sizeof(LayerRec) * s.layers.size();
    h.edge_off = (uint32_t)off;
    off += sizeof(EdgeRec) * s.edges.size();
    h.step_off = (uint32_t)off;
    off += sizeof(StepRec) * s.steps.size();
    const size_t tcol_off = off;
    off += s.tcol.size();
    off = align16(off);
    h.smem_bytes = (uint32_t)(off - h.smem_off);
    // Use the tensor-memory kernel variant only when most of the scalar wavefront steps got columns:
    // measured on B200, codes where under ~3/4 of them fit (DVB-S2 3/4, 3/5 normal) ran slower with a
    // partial placement than with all state in L2, codes where they fit (1/2 normal, 2/3 short) faster.
    h.tmem_cols = 0;
    {
        int placed = 0, scalar_steps = 0;
        for (size_t k = 0; k < s.steps.size(); ++k) {
            if (s.steps[k].count == 0 || (s.steps[k].work_off & (kStepLinkParallel | kStepSplit)))
                continue;
            ++scalar_steps;
            placed += s.tcol[k] != kNoTmem;
        }
        if (placed > 0 && 4 * placed >= 3 * scalar_steps)
            h.tmem_cols = kTmemCols;
    }
    h.order_off = (uint32_t)off;
    off = align16(off + sizeof(uint16_t) * s.order.size());
    std::vector<uint16_t> al, lg;
    gf_tables(pp, al, lg);
    h.antilog_off = (uint32_t)off;
    off = align16(off + sizeof(uint16_t) * al.size());
    h.log_off = (uint32_t)off;
    off = align16(off + sizeof(uint16_t) * lg.size());
    h.total_bytes = (uint32_t)off;

    blob.assign(off, 0);
    memcpy(blob.data(), &h, sizeof(h));
    memcpy(blob.data() + h.layer_off, s.layers.data(), sizeof(LayerRec) * s.layers.size());
    memcpy(blob.data() + h.edge_off, s.edges.data(), sizeof(EdgeRec) * s.edges.size());
    if (!s.steps.empty())
        memcpy(blob.data() + h.step_off, s.steps.data(), sizeof(StepRec) * s.steps.size());
    memcpy(blob.data() + tcol_off, s.tcol.data(), s.tcol.size());
    if (!s.order.empty())
        memcpy(blob.data() + h.order_off, s.order.data(), sizeof(uint16_t) * s.order.size());
    memcpy(blob.data() + h.antilog_off, al.data(), sizeof(uint16_t) * al.size());
    memcpy(blob.data() + h.log_off, lg.data(), sizeof(uint16_t) * lg.size());
    return true;
}

bool validate_blob(const void* blob, size_t size, std::string& err)
{
    if (!blob || size < sizeof(BlobHeader)) {
        err = "table blob too small";
        return false;
    }
    BlobHeader h;
    memcpy(&h, blob, sizeof(h));
    if (h.magic != kBlobMagic || h.version != kBlobVersion) {
        err = "table blob has wrong magic/version";
        return false;
    }
    if (h.total_bytes != size) {
        err = "table blob size mismatch";
        return false;
    }
    const size_t gfn = (size_t)2 << h.gf_m;
    bool ok = h.q > 0 && h.q <= 180 && h.R == h.q * 360 && h.N == h.K + h.R && h.n_circ > 0 &&
              h.smem_off == sizeof(BlobHeader) && h.layer_off == h.smem_off &&
              h.edge_off == h.layer_off + sizeof(LayerRec) * (size_t)h.q &&
              (size_t)h.smem_off + h.smem_bytes <= size && h.step_off <= h.smem_off + h.smem_bytes && h.order_off <= size &&
              h.n_steps_total > 0 && h.n_steps_total <= h.steps_per_iter &&
              (size_t)h.antilog_off + gfn <= size && (size_t)h.log_off + gfn <= size &&
              (h.msg_words == 1 || h.msg_words == 2) && h.gf_m >= 14 && h.gf_m <= 16;
    if (!ok) {
        err = "table blob failed consistency checks";
        return false;
    }
    // The kernels trust every index in the blob (step list, circulants, work lists).  The blob is a pure
    // function of (standard, framesize, rate): rebuild it (a few ms) and compare -- anything corrupted,
    // truncated or built by another version of the library is rejected here instead of faulting on the device.
    std::vector<uint8_t> again;
    std::string berr;
    if (!build_blob(h.standard, h.framesize, h.rate, again, berr) || again.size() != size ||
        memcmp(again.data(), blob, size) != 0) {
        err = "table blob does not match the tables this library builds for its (standard, framesize, rate)";
        return false;
    }
    return true;
}

} // namespace dvbs2b200
