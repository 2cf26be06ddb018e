// ldpc_steps.cuh -- what one thread does in a schedule step (host/device; see ldpc_core.cuh).
//
//   pair_step    conflict-free layer: nodes p and p+180, every link in s16x2.
//   split_*      layer whose circulants share a 360-bit group (order sensitive in the reference, which visits the
//                check nodes serially: lib/ldpc_decoder/layered_decoder.hh:50-79):
//                  split_p1     private links of both nodes in s16x2 -> partial minima / signs
//                  level form   level_prep() leaves what the serial phase needs of nodes p, p+180 in shared memory;
//                               level_link_load() / level_link_store(): one shared link of one node, run by any
//                               lane at the node's level of the serial order; level_p3_links() merges the values
//                               the nodes saw into the pair's accumulators
//                  chain form   chain_p1() writes a record per node, chain_node() walks a chain and hands the
//                               updated bit on in a register, chain_p3_links() redoes the shared links in s16x2
//                               with the inputs the walk saw
//                  split_p3     final minima / signs -> new state, private (and, chain form, shared) links updated
// The kernel puts the barriers between these calls; tools/ldpc_emul.cc runs them thread by thread.
#pragma once
#include "ldpc_core.cuh"

namespace dvbs2b200 {
namespace core {

template <int NW>
struct RawState { // as stored: word of clamped minima + field words
    uint32_t Cw;
    uint32_t W[NW];
};

struct FrameCtx {
    uint8_t* L;            // biased posteriors, pair-interleaved
    const LayerRec* layers;
    const EdgeRec* edges;
    int K, q;
};

template <int NW>
LDPC_HD void decode_state(const RawState<NW>& in, PairState<NW>& s)
{
    decode_minima(in.Cw, s.candA, s.candB);
#pragma unroll
    for (int w = 0; w < NW; ++w)
        s.W[w] = in.W[w];
}

LDPC_HD uint32_t lds16(const uint8_t* L, uint32_t adr) { return (uint32_t) * reinterpret_cast<const uint16_t*>(L + adr); }
LDPC_HD void sts16(uint8_t* L, uint32_t adr, uint32_t v) { *reinterpret_cast<uint16_t*>(L + adr) = (uint16_t)v; }

// operands of link d of the pair (0 = own parity bit, 1 = previous parity bit, 2.. = data links)
LDPC_HD LinkOp link_operand(const FrameCtx& c, const ThreadConst& tc, uint32_t edge_begin, int layer, int d, bool first)
{
    LinkOp o;
    if (d == 0) {
        o.adr = (uint32_t)c.K + 2u * ((uint32_t)c.q * tc.p + (uint32_t)layer);
        o.g2 = 1u;
    } else if (d == 1) {
        // check 0 (layer 0, pair 0, node A) has no previous parity bit; node B's (parity bit 180q - 1) is the
        // LOW byte of the last parity halfword
        o.adr = first ? (uint32_t)c.K + 2u * (uint32_t)(kPairs * c.q - 1) : (uint32_t)c.K + 2u * ((uint32_t)c.q * tc.p + (uint32_t)layer) - 2u;
        o.g2 = 1u;
    } else {
        o = data_link(c.edges[edge_begin + d - 2], tc);
    }
    return o;
}
LDPC_HD uint32_t link_load(const FrameCtx& c, const LinkOp& o, int d, bool first)
{
    const uint32_t raw = lds16(c.L, o.adr);
    if (d == 1 && first)
        return prmt(raw, 0u, 0x4044u) | 0x000000ffu; // node B from the low byte; node A: a link that never matters
    return prmt(raw, 0u, sel_unpack(o.g2));
}
LDPC_HD void link_store(const FrameCtx& c, const LinkOp& o, int d, bool first, uint32_t ub)
{
    if (d == 1 && first)
        c.L[o.adr] = (uint8_t)(ub >> 16); // only node B's link exists
    else
        sts16(c.L, o.adr, prmt(ub, 0u, sel_pack(o.g2)));
}
// x of the non-existent link of check 0: +127, the largest magnitude there is (never changes a minimum that matters)
LDPC_HD uint32_t first_fix(uint32_t xb, int d, bool first) { return (d == 1 && first) ? ((xb & 0xffff0000u) | 0x00ffu) : xb; }

// unsatisfied-check test from the new posteriors of all links of the pair (layered_decoder.hh:32-49)
LDPC_HD int syndrome_bad(uint32_t syn, uint32_t zer, int deg)
{
    const uint32_t odd_neg = ((syn >> 7) ^ ((deg & 1) ? 0x00010001u : 0u)) & 0x00010001u; // bit 7: parity of the links with x >= 0
    return (int)((odd_neg | (zer & 0x80008000u)) != 0u);
}

// ---- conflict-free layer -----------------------------------------------------------------------------
// For check nodes of up to 9 data links the operand addresses of all links are computed first (a dense block of
// independent multiply-adds on the FMA pipe) and the loads of the posteriors follow back to back: measured +3 % on
// 1/2 normal against computing each address next to its load; wider check nodes lose by it (registers), and so do
// the phases of the split steps (their kernels spill).  A split-phase mbarrier in front of the step, with this table
// work between arrive and wait, was measured too: -7 % (the polling wait costs more issue slots than it frees).
template <int CNT_MAX, bool UNIFORM, bool SELF_CHECK, int NW>
LDPC_HD int pair_step(const FrameCtx& c, const ThreadConst& tc, int layer, const RawState<NW>& in, RawState<NW>& out)
{
    constexpr int DEG_MAX = CNT_MAX + 2;
    constexpr bool ADDRESSES_FIRST = CNT_MAX <= 9;
    const LayerRec lr = c.layers[layer];
    const int deg = UNIFORM ? DEG_MAX : (int)lr.cnt + 2;
    const bool first = (layer == 0 && tc.p == 0u);
    PairState<NW> st;
    decode_state(in, st);
    Acc<NW> acc;
    acc_init(acc);
    LinkOp op[DEG_MAX];
    uint32_t xb[DEG_MAX];
    if (ADDRESSES_FIRST) {
#pragma unroll
        for (int d = 0; d < DEG_MAX; ++d)
            if (UNIFORM || d < deg)
                op[d] = link_operand(c, tc, lr.edge_begin, layer, d, first);
    }
    uint32_t wsh = 0u;
#pragma unroll
    for (int d = 0; d < DEG_MAX; ++d) {
        wsh = field_word(st.W[d >> 3], wsh, d, tc);
        if (UNIFORM || d < deg) {
            if (!ADDRESSES_FIRST)
                op[d] = link_operand(c, tc, lr.edge_begin, layer, d, first);
            const uint32_t u = link_load(c, op[d], d, first);
            xb[d] = first_fix(link_x(u, wsh, st.candA, st.candB), d, first);
            link_merge(acc, xb[d], d, tc);
        }
    }
    Final<NW> f;
    finalize(acc, deg, f);
    uint32_t syn = 0u, zer = 0u;
#pragma unroll
    for (int d = 0; d < DEG_MAX; ++d) {
        wsh = field_word(f.W[d >> 3], wsh, d, tc);
        if (UNIFORM || d < deg) {
            const uint32_t ub = link_new(xb[d], wsh, f);
            link_store(c, op[d], d, first, ub);
            if (SELF_CHECK)
                syndrome_acc(ub, syn, zer);
        }
    }
    out.Cw = f.Cw;
#pragma unroll
    for (int w = 0; w < NW; ++w)
        out.W[w] = f.W[w];
    return SELF_CHECK ? syndrome_bad(syn, zer, deg) : 0;
}

// syndrome test of the pair on the posteriors as they are (layered_decoder.hh:32-49)
template <int CNT_MAX, bool UNIFORM>
LDPC_HD int check_pair(const FrameCtx& c, const ThreadConst& tc, int layer)
{
    constexpr int DEG_MAX = CNT_MAX + 2;
    const LayerRec lr = c.layers[layer];
    const int deg = UNIFORM ? DEG_MAX : (int)lr.cnt + 2;
    const bool first = (layer == 0 && tc.p == 0u);
    uint32_t syn = 0u, zer = 0u;
#pragma unroll
    for (int d = 0; d < DEG_MAX; ++d) {
        if (UNIFORM || d < deg) {
            const LinkOp o = link_operand(c, tc, lr.edge_begin, layer, d, first);
            syndrome_acc(link_load(c, o, d, first), syn, zer);
        }
    }
    return syndrome_bad(syn, zer, deg);
}

// ---- conflict layer ------------------------------------------------------------------------------------
template <int CNT_MAX, int NW>
struct SplitRegs {
    PairState<NW> st;
    Acc<NW> acc;
    uint32_t xb[CNT_MAX + 2]; // private links
    uint32_t xb_in, xb_out;   // chain form: the two shared links in phase 3
    uint32_t u_out;           // chain form: unpacked posteriors of the forwarding link as phase 1 saw them
    LinkOp op_in, op_out;     // chain form: the two shared links
    int npriv;                // private data links
    int deg;
    bool first;
};

template <int CNT_MAX, bool UNIFORM, int NW>
LDPC_HD void split_p1(const FrameCtx& c, const ThreadConst& tc, int layer, const RawState<NW>& in, SplitRegs<CNT_MAX, NW>& r)
{
    constexpr int DEG_MAX = CNT_MAX + 2;
    const LayerRec lr = c.layers[layer];
    r.deg = UNIFORM ? DEG_MAX : (int)lr.cnt + 2;
    r.npriv = (int)lr.cnt - (int)lr.conflict;
    r.first = (layer == 0 && tc.p == 0u);
    decode_state(in, r.st);
    acc_init(r.acc);
    uint32_t wsh = 0u;
#pragma unroll
    for (int d = 0; d < DEG_MAX; ++d) {
        wsh = field_word(r.st.W[d >> 3], wsh, d, tc);
        if (d < 2 + r.npriv) {
            const LinkOp o = link_operand(c, tc, lr.edge_begin, layer, d, r.first);
            const uint32_t u = link_load(c, o, d, r.first);
            r.xb[d] = first_fix(link_x(u, wsh, r.st.candA, r.st.candB), d, r.first);
            link_merge(r.acc, r.xb[d], d, tc);
        }
    }
}

// ---- level form: the nodes of a level, handed out to the threads of the CTA ---------------------------------
// scalar view of one half of the accumulators
template <int NW>
LDPC_HD int half_nonneg(const Acc<NW>& a, int hs)
{
    int n = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w)
        n += popc(sgn_word(a, w) & (hs ? 0x55550000u : 0x00005555u));
    return n;
}
// What the serial phase needs of a node does not depend on what its predecessors do, except the shared bits
// themselves.  Phase 1 (thread p, nodes p and p+180) leaves it in the step's shared-memory scratch (LevelScratch):
//   adr[s * 180 + p], s < nshared: byte address of node A's operand of shared link s (node B's is that ^ 1: the two
//       bytes of one halfword)
//   msg[s * 360 + 2p + hs]: -(old message) of the node through shared link s; the serial phase replaces it by the v->c
//       value + 128 that the node saw, for phase 3
//   nq[2p + hs]: the smallest magnitude of the node's private links, capped at 126, and in bit 7 the parity of its
//       negative private v->c values
//   first_node[l]: copy of the step's level table (code_tables.h)
struct LevelScratch {
    uint16_t* adr;
    uint8_t* msg;
    uint8_t* nq;
    uint16_t* first_node;
};
LDPC_HD LevelScratch level_scratch(void* base, int nshared)
{
    LevelScratch ls;
    ls.adr = reinterpret_cast<uint16_t*>(base);
    ls.msg = reinterpret_cast<uint8_t*>(base) + 2 * kPairs * nshared;
    ls.nq = ls.msg + 2 * kPairs * nshared;
    ls.first_node = reinterpret_cast<uint16_t*>(ls.nq + 2 * kPairs);
    return ls;
}
template <int CNT_MAX, int NW>
LDPC_HD void level_prep(const FrameCtx& c, const ThreadConst& tc, int layer, const LevelScratch& ls, SplitRegs<CNT_MAX, NW>& r)
{
    const LayerRec lr = c.layers[layer];
    const int nshared = (int)lr.conflict, d0 = 2 + r.npriv;
    for (int s = 0; s < nshared; ++s) {
        const int d = d0 + s;
        const LinkOp o = data_link(c.edges[lr.edge_begin + r.npriv + s], tc);
        const uint32_t wsh = pick_word(r.st.W, d >> 3) >> field_shift(d);
        const uint32_t no = prmt(r.st.candA, r.st.candB, imad(wsh & 0x0303u, 0x11u, 0xc480u));
        ls.adr[s * kPairs + (int)tc.p] = (uint16_t)(o.adr + 1u - o.g2);
        reinterpret_cast<uint16_t*>(ls.msg)[s * kPairs + (int)tc.p] = (uint16_t)prmt(no, 0u, 0x4420u);
    }
    const uint32_t qb = vmin2((r.acc.k0 >> 5) & 0x07ff07ffu, h2(126));
    const uint32_t pnA = (uint32_t)((d0 - half_nonneg(r.acc, 0)) & 1), pnB = (uint32_t)((d0 - half_nonneg(r.acc, 1)) & 1);
    reinterpret_cast<uint16_t*>(ls.nq)[(int)tc.p] = (uint16_t)((qb & 0x7fu) | (pnA << 7) | (((qb >> 16) & 0x7fu) << 8) | (pnB << 15));
}

// One shared link s of one node (pair p, half hs) of the serial order: a lane of its own.  The dependent chain of a
// level is what the serial phase costs (at the ~7 cycles per dependent instruction a lone warp gets on a busy SM), so
// the links of a node sit in neighbouring lanes and meet in three warp reductions (ldpc_kernel.cu: level_phase; the
// CPU emulation loops): the two smallest keys and the number of negative values over the node's shared links.
struct LevelLink {
    uint32_t adr; // byte address of the operand
    int xb;       // v->c value + 128
    uint32_t key; // max(|x| - 1, 0) * 32 + s
    uint32_t nq;  // the node's private minimum (bits 0-6) and sign parity (bit 7)
};
constexpr uint32_t kLevelNoKey = 0x7fffu; // a lane without a link (keys are below 128 * 32)
LDPC_HD void level_link_load(const uint8_t* L, const LevelScratch& ls, int p, int hs, int s, LevelLink& k)
{
    k.nq = ls.nq[2 * p + hs];
    k.adr = (uint32_t)ls.adr[s * kPairs + p] ^ (uint32_t)hs;
    const int no = (int)(int8_t)ls.msg[2 * kPairs * s + 2 * p + hs];
    k.xb = clamp255((int)L[k.adr] + no);
    k.key = (uint32_t)(mag_scalar(k.xb) * 32 + s);
}
// k0, k1: the two smallest keys over the node's shared links; negs: how many of them are negative (its parity counts).
// The message to this link is the minimum over the node's OTHER links -- the private minimum and the other shared
// links -- with the sign parity of the other links.
LDPC_HD void level_link_store(uint8_t* L, const LevelScratch& ls, int p, int hs, int s, const LevelLink& k, uint32_t k0, uint32_t k1, uint32_t negs)
{
    const uint32_t q = k.nq & 0x7fu;
    const uint32_t other = (k.key == k0 ? k1 : k0) >> 5;
    const int m = (int)(other < q ? other : q);
    const uint32_t isneg = k.xb < 128 ? 1u : 0u;
    const uint32_t neg = ((k.nq >> 7) ^ negs ^ isneg) & 1u;
    L[k.adr] = (uint8_t)clamp255(k.xb + (neg ? -m : m));
    ls.msg[2 * kPairs * s + 2 * p + hs] = (uint8_t)k.xb;
}

// phase 3 of the level form: the shared links join the accumulators of the pair with the values the nodes saw
template <int CNT_MAX, int NW>
LDPC_HD void level_p3_links(const ThreadConst& tc, int nshared, const LevelScratch& ls, SplitRegs<CNT_MAX, NW>& r)
{
    const int d0 = 2 + r.npriv;
    for (int s = 0; s < nshared; ++s) {
        const uint32_t v = reinterpret_cast<const uint16_t*>(ls.msg)[s * kPairs + (int)tc.p];
        link_merge(r.acc, prmt(v, 0u, 0x4140u), d0 + s, tc);
    }
}

// ---- chain form ---------------------------------------------------------------------------------------------
// Node record, written by phase 1, read by the walk:
//   x: byte address of the in-link operand | byte address of the out-link operand << 16
//   y: -(old message) of the in link (byte 0) and of the out link (byte 1), bits 16..23 smallest magnitude of the
//      private links (capped at 126), bit 24 parity of the negative private v->c values
struct ChainRec {
    uint32_t x, y;
};
// The walk leaves the in-link posterior (biased) that node j saw in the first byte of its record (the record has
// been consumed by then): phase 3 redoes the node with it.
constexpr int kLinStride = (int)sizeof(ChainRec);

template <int CNT_MAX, bool UNIFORM, int NW>
LDPC_HD void chain_p1(const FrameCtx& c, const ThreadConst& tc, int layer, int out_link1, ChainRec* rec, SplitRegs<CNT_MAX, NW>& r)
{
    const LayerRec lr = c.layers[layer];
    const int d0 = 2 + r.npriv;
    const int d_in = out_link1 ? d0 : d0 + 1, d_out = out_link1 ? d0 + 1 : d0;
    r.op_in = data_link(c.edges[lr.edge_begin + d_in - 2], tc);
    r.op_out = data_link(c.edges[lr.edge_begin + d_out - 2], tc);
    r.u_out = prmt(lds16(c.L, r.op_out.adr), 0u, sel_unpack(r.op_out.g2));
    const uint32_t f_in = (pick_word(r.st.W, d_in >> 3) >> field_shift(d_in)) & 0x0303u;
    const uint32_t f_out = (pick_word(r.st.W, d_out >> 3) >> field_shift(d_out)) & 0x0303u;
    const uint32_t no_in = prmt(r.st.candA, r.st.candB, imad(f_in, 0x11u, 0xc480u));
    const uint32_t no_out = prmt(r.st.candA, r.st.candB, imad(f_out, 0x11u, 0xc480u));
    const uint32_t qb = vmin2((r.acc.k0 >> 5) & 0x07ff07ffu, h2(126));
    const int npl = 2 + r.npriv; // private links, the missing one of check 0 counted as positive
#pragma unroll
    for (int hs = 0; hs < 2; ++hs) {
        const uint32_t a_in = r.op_in.adr + (hs ? r.op_in.g2 : 1u - r.op_in.g2), a_out = r.op_out.adr + (hs ? r.op_out.g2 : 1u - r.op_out.g2);
        const uint32_t pn = (uint32_t)((npl - half_nonneg(r.acc, hs)) & 1);
        ChainRec cr;
        cr.x = a_in | (a_out << 16);
        cr.y = ((no_in >> (16 * hs)) & 0xffu) | (((no_out >> (16 * hs)) & 0xffu) << 8) | (((qb >> (16 * hs)) & 0xffu) << 16) | (pn << 24);
        rec[(int)tc.p + kPairs * hs] = cr;
    }
}

// one node of the walk: the in link holds l_in (biased).  Returns the updated bit of the out link (biased); for the
// first node of a chain (with_in) also updates the in-link bit in L -- the last node of some chain meets it.
LDPC_HD int chain_node(uint8_t* L, const ChainRec cr, int l_in, int l_out, bool with_in)
{
    const int no_in = (int)(int8_t)(cr.y & 0xffu), no_out = (int)(int8_t)((cr.y >> 8) & 0xffu);
    const int q = (int)((cr.y >> 16) & 0xffu), pn = (int)((cr.y >> 24) & 1u);
    const int xi = clamp255(l_in + no_in), xo = clamp255(l_out + no_out);
    const int mi = mag_scalar(xi), mo = mag_scalar(xo); // q <= 126 caps them
    const int m_out = mi < q ? mi : q; // smallest magnitude over the links other than the out link
    const int neg_out = pn ^ (xi < 128 ? 1 : 0);
    if (with_in) {
        const int m_in = mo < q ? mo : q;
        const int neg_in = pn ^ (xo < 128 ? 1 : 0);
        L[cr.x & 0xffffu] = (uint8_t)clamp255(xi + (neg_in ? -m_in : m_in));
    }
    return clamp255(xo + (neg_out ? -m_out : m_out));
}

// phase 3 of the chain form: the two shared links join the accumulators with the inputs the walk saw
template <int CNT_MAX, int NW>
LDPC_HD void chain_p3_links(const FrameCtx& c, const ThreadConst& tc, int out_link1, int delta, const uint8_t* lin, SplitRegs<CNT_MAX, NW>& r)
{
    const int d0 = 2 + r.npriv;
    const int d_in = out_link1 ? d0 : d0 + 1, d_out = out_link1 ? d0 + 1 : d0;
    const uint32_t u_in = (uint32_t)lin[kLinStride * tc.p] | ((uint32_t)lin[kLinStride * (tc.p + kPairs)] << 16);
    uint32_t u_out = r.u_out;
    if ((int)tc.p + kPairs + delta >= 360) { // node B is the last of its chain: its out-link bit was updated by a first node
        const uint32_t a_out = r.op_out.adr + r.op_out.g2;
        u_out = (u_out & 0x0000ffffu) | ((uint32_t)c.L[a_out] << 16);
    }
    r.xb_in = link_x(u_in, pick_word(r.st.W, d_in >> 3) >> field_shift(d_in), r.st.candA, r.st.candB);
    r.xb_out = link_x(u_out, pick_word(r.st.W, d_out >> 3) >> field_shift(d_out), r.st.candA, r.st.candB);
    link_merge(r.acc, r.xb_in, d_in, tc);
    link_merge(r.acc, r.xb_out, d_out, tc);
}
// ... and their update: every shared bit is written by the LAST node of the serial order that touches it
template <int CNT_MAX, int NW>
LDPC_HD void chain_p3_store(const FrameCtx& c, const ThreadConst& tc, int out_link1, int delta, const Final<NW>& f, SplitRegs<CNT_MAX, NW>& r,
                            uint32_t& syn, uint32_t& zer, bool self_check)
{
    const int d0 = 2 + r.npriv;
    const int d_in = out_link1 ? d0 : d0 + 1, d_out = out_link1 ? d0 + 1 : d0;
    const uint32_t ub_in = link_new(r.xb_in, pick_word(f.W, d_in >> 3) >> field_shift(d_in), f);
    const uint32_t ub_out = link_new(r.xb_out, pick_word(f.W, d_out >> 3) >> field_shift(d_out), f);
    // in link: node B always; node A unless it is the first node of its chain (its bit is rewritten by a last node)
    c.L[r.op_in.adr + r.op_in.g2] = (uint8_t)(ub_in >> 16);
    if ((int)tc.p >= delta)
        c.L[r.op_in.adr + 1u - r.op_in.g2] = (uint8_t)ub_in;
    // out link: only the last node of a chain (nobody takes the bit over); node A is never last
    if ((int)tc.p + kPairs + delta >= 360)
        c.L[r.op_out.adr + r.op_out.g2] = (uint8_t)(ub_out >> 16);
    (void)syn, (void)zer, (void)self_check;
}

// ---- phase 3: new state, private links updated ---------------------------------------------------------------
template <int CNT_MAX, bool UNIFORM, int NW>
LDPC_HD void split_p3(const FrameCtx& c, const ThreadConst& tc, int layer, SplitRegs<CNT_MAX, NW>& r, Final<NW>& f, RawState<NW>& out)
{
    constexpr int DEG_MAX = CNT_MAX + 2;
    const LayerRec lr = c.layers[layer];
    finalize(r.acc, r.deg, f);
    uint32_t wsh = 0u;
#pragma unroll
    for (int d = 0; d < DEG_MAX; ++d) {
        wsh = field_word(f.W[d >> 3], wsh, d, tc);
        if (d < 2 + r.npriv) {
            const LinkOp o = link_operand(c, tc, lr.edge_begin, layer, d, r.first);
            link_store(c, o, d, r.first, link_new(r.xb[d], wsh, f));
        }
    }
    out.Cw = f.Cw;
#pragma unroll
    for (int w = 0; w < NW; ++w)
        out.W[w] = f.W[w];
}

} // namespace core
} // namespace dvbs2b200
