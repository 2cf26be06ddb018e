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
void build_schedule(const LdpcTableDef& def, Schedule& s)
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
        { // n_shared <= kMaxSharedLinks holds for all 57 tables (asserted by tests/test_host_cpu.py)
            // One "split" step: private links of all 360 check nodes in parallel, the shared links level by
            // level (code_tables.h).  work[] holds level[j] for j = 0..359.
            StepRec st;
            st.layer = (uint8_t)i;
            st.run_len = 0;
            st.count = (uint16_t)depth;
            st.work_off = (uint32_t)s.order.size() | kStepSplit;
            for (int j = 0; j < 360; ++j)
                s.order.push_back((uint16_t)level[j]);
            // first node of each level (levels rise with j)
            s.order.push_back(0);
            for (int lv = 1, j = 0; lv <= depth + 1; ++lv) {
                while (j < 360 && level[j] < lv)
                    ++j;
                s.order.push_back((uint16_t)j);
            }
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
            if (!(st.work_off & kStepChain))
                s.max_level_shared = std::max(s.max_level_shared, n_shared);
            s.steps.push_back(st);
            continue;
        }
    }

    // ---- barrier placement ---------------------------------------------------------------------
    // dirty: groups written since the last block barrier.  Parity links are thread private in every step (pair
    // and split steps keep the pair mapping), except between the last layer and layer 0 of the next iteration,
    // which starts behind a barrier anyway.
    std::vector<char> dirty(ngroups, 0);
    bool parity_dirty = false; // layer 0 wrote the parity bits q*(p-1) + q-1 that thread p-1 owns in layer q-1
    for (size_t k = 0; k < s.steps.size(); ++k) {
        StepRec& st = s.steps[k];
        const bool is_split = (st.work_off & kStepSplit) != 0;
        bool need = false;
        for (auto& ga : per_layer[st.layer])
            need |= dirty[ga.first] != 0;
        if (st.layer == q - 1 && parity_dirty)
            need = true;
        if (k == 0)
            need = false; // the iteration starts behind a barrier
        // a split step always starts behind a block barrier: its scratch and named barriers are reused from
        // one split step to the next, and its serial phase reads what any thread wrote before
        if (need || (is_split && k > 0)) {
            st.work_off |= kStepBarrierBefore;
            s.barriers_per_iter++;
            std::fill(dirty.begin(), dirty.end(), 0);
            parity_dirty = false;
        }
        for (auto& ga : per_layer[st.layer])
            dirty[ga.first] = 1;
        if (st.layer == 0)
            parity_dirty = true;
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
    // compressed check-node state per check-node pair: one word of clamped minima (4 x 6 bits in bytes) and two
    // bits (sign, argmin) per link and node, 8 links per word (ldpc_core.cuh)
    h.msg_words = 1 + (s.max_cnt + 2 + 7) / 8;
    h.n_steps_total = (int32_t)s.steps.size();
    h.n_conflict_layers = s.conflict_layers;
    h.steps_per_iter = s.steps_per_iter;
    h.max_depth = s.max_depth;
    h.uniform_cnt = (s.min_cnt == s.max_cnt) ? 1 : 0;
    h.bch_shorten = ((1u << h.gf_m) - 1) - (uint32_t)mc->nbch;
    h.split_steps = 1u;
    h.chain_scratch = (uint32_t)std::max(s.has_chain ? 360 * 8 : 0,
                                         s.max_level_shared ? 720 * s.max_level_shared + 360 + (int)align16(2 * (s.max_depth + 2)) : 0);
    h.level_calls = 0u;

    size_t off = sizeof(BlobHeader);
    h.smem_off = (uint32_t)off;
    h.layer_off = (uint32_t)off;
    off += sizeof(LayerRec) * s.layers.size();
    h.edge_off = (uint32_t)off;
    off += sizeof(EdgeRec) * s.edges.size();
    h.step_off = (uint32_t)off;
    off += sizeof(StepRec) * s.steps.size();
    off = align16(off);
    h.smem_bytes = (uint32_t)(off - h.smem_off);
    h.tmem_cols = 0;
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
              h.msg_words >= 2 && h.msg_words <= 5 && h.gf_m >= 14 && h.gf_m <= 16;
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
