// ldpc_emul.cc -- CPU emulation of the LDPC kernel's thread functions (gr-dvbs2rx_b200/csrc/ldpc_steps.cuh),
// checked against the oracle.  Development tool: the arithmetic and the step protocol of the kernel can be
// verified here, without a GPU, on every code table; what it cannot see are races -- the phases of a step run
// one thread after the other.  Test infrastructure (it links oracle/liboracle.so), never part of the product.
//
//   g++ -std=c++17 -O2 -I gr-dvbs2rx_b200/csrc tools/ldpc_emul.cc gr-dvbs2rx_b200/csrc/code_tables.cc \
//       -Loracle -loracle -Wl,-rpath,$PWD/oracle -o /tmp/ldpc_emul && /tmp/ldpc_emul [table ...]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../oracle/dvbs2_oracle.h"
#include "ldpc_steps.cuh"

using namespace dvbs2b200;
using namespace dvbs2b200::core;

static int data_addr(int group_base, int s) { return group_base + 2 * (s >= kPairs ? s - kPairs : s) + (s >= kPairs); }
static int parity_addr(int K, int half, int c) { return K + 2 * (c >= half ? c - half : c) + (c >= half); }
static int pos_of_bit(int n, int K, int R)
{
    if (n < K)
        return data_addr((n / 360) * 360, n % 360);
    return parity_addr(K, R / 2, n - K);
}

template <int CNT_MAX, bool UNIFORM>
static int run_frame(const LdpcTableDef& def, const Schedule& s, const int8_t* llr, int trials, int8_t* post)
{
    constexpr int NW = (CNT_MAX + 2 + 7) / 8;
    const int N = def.N, K = def.K, R = N - K, q = def.q;
    std::vector<uint8_t> L(N + 16);
    for (int n = 0; n < N; ++n)
        L[pos_of_bit(n, K, R)] = (uint8_t)(llr[n] ^ 0x80);
    FrameCtx c{ L.data(), s.layers.data(), s.edges.data(), K, q };
    std::vector<RawState<NW>> state((size_t)q * kPairs);
    std::vector<ChainRec> rec(4096); // the step's shared-memory scratch
    auto tconst = [](int p) {
        ThreadConst tc;
        tc.p = (uint32_t)p;
        tc.two = 2u;
        tc.four = 4u;
        tc.c30 = 1u << 30;
        tc.c16 = 1u << 16;
        tc.c32 = 32u;
        tc.neg1 = 0xffffffffu;
        return tc;
    };
    int left = trials;
    bool zero_state = true;
    for (;;) {
        int bad = 0;
        for (int i = 0; i < q && !bad; ++i)
            for (int p = 0; p < kPairs; ++p)
                bad |= check_pair<CNT_MAX, UNIFORM>(c, tconst(p), i);
        if (!bad || --left < 0)
            break;
        for (const StepRec& st : s.steps) {
            const int layer = st.layer;
            auto state_in = [&](int p) {
                RawState<NW> in;
                if (zero_state) {
                    in.Cw = 0;
                    for (int w = 0; w < NW; ++w)
                        in.W[w] = 0;
                } else
                    in = state[(size_t)layer * kPairs + p];
                return in;
            };
            if (st.count == 0) {
                for (int p = 0; p < kPairs; ++p) {
                    RawState<NW> out;
                    pair_step<CNT_MAX, UNIFORM, false, NW>(c, tconst(p), layer, state_in(p), out);
                    state[(size_t)layer * kPairs + p] = out;
                }
                continue;
            }
            const bool chain = (st.work_off & kStepChain) != 0;
            const int out_link1 = (st.work_off & kStepChainOutLink1) ? 1 : 0;
            const int delta = st.run_len, depth = st.count;
            const uint16_t* level = s.order.data() + (st.work_off & kStepOffMask);
            std::vector<SplitRegs<CNT_MAX, NW>> regs(kPairs);
            for (int p = 0; p < kPairs; ++p)
                split_p1<CNT_MAX, UNIFORM, NW>(c, tconst(p), layer, state_in(p), regs[p]);
            if (chain) {
                for (int p = 0; p < kPairs; ++p)
                    chain_p1<CNT_MAX, UNIFORM, NW>(c, tconst(p), layer, out_link1, rec.data(), regs[p]);
                std::vector<int> carried(delta);
                uint8_t* lin = reinterpret_cast<uint8_t*>(rec.data());
                for (int l = 0; l < delta; ++l) { // first nodes
                    const ChainRec cr = rec[l];
                    const int l_in = L[cr.x & 0xffffu], l_out = L[cr.x >> 16];
                    lin[kLinStride * l] = (uint8_t)l_in;
                    carried[l] = chain_node(L.data(), cr, l_in, l_out, true);
                }
                for (int l = 0; l < delta; ++l)
                    for (int j = l + delta; j < 360; j += delta) {
                        const ChainRec cr = rec[j];
                        lin[kLinStride * j] = (uint8_t)carried[l];
                        carried[l] = chain_node(L.data(), cr, carried[l], L[cr.x >> 16], false);
                    }
                for (int p = 0; p < kPairs; ++p)
                    chain_p3_links<CNT_MAX, NW>(c, tconst(p), out_link1, delta, reinterpret_cast<const uint8_t*>(rec.data()), regs[p]);
            } else {
                const int nshared = (int)c.layers[layer].conflict;
                LevelScratch ls = level_scratch(rec.data(), nshared);
                const uint16_t* first_node = level + 360;
                for (int p = 0; p < kPairs; ++p)
                    level_prep<CNT_MAX, NW>(c, tconst(p), layer, ls, regs[p]);
                for (int lvl = 1; lvl <= depth; ++lvl)
                    for (int j = first_node[lvl]; j < first_node[lvl + 1]; ++j) {
                        if (level[j] != lvl) {
                            printf("first_node table inconsistent\n");
                            exit(1);
                        }
                        LevelLink k[kMaxSharedLinks];
                        uint32_t k0 = kLevelNoKey, k1 = kLevelNoKey, negs = 0;
                        for (int sl = 0; sl < nshared; ++sl) {
                            level_link_load(L.data(), ls, j % kPairs, j / kPairs, sl, k[sl]);
                            if (k[sl].key < k0) {
                                k1 = k0;
                                k0 = k[sl].key;
                            } else if (k[sl].key < k1)
                                k1 = k[sl].key;
                            negs ^= k[sl].xb < 128 ? 1u : 0u;
                        }
                        for (int sl = 0; sl < nshared; ++sl)
                            level_link_store(L.data(), ls, j % kPairs, j / kPairs, sl, k[sl], k0, k1, negs);
                    }
                for (int p = 0; p < kPairs; ++p)
                    level_p3_links<CNT_MAX, NW>(tconst(p), nshared, ls, regs[p]);
            }
            // phase 3: every thread finalizes first (reads), then stores -- the kernel has no barrier in between,
            // and needs none: see ldpc_kernel.cu
            for (int p = 0; p < kPairs; ++p) {
                Final<NW> f;
                RawState<NW> out;
                split_p3<CNT_MAX, UNIFORM, NW>(c, tconst(p), layer, regs[p], f, out);
                if (chain) {
                    uint32_t syn = 0, zer = 0;
                    chain_p3_store<CNT_MAX, NW>(c, tconst(p), out_link1, delta, f, regs[p], syn, zer, false);
                }
                state[(size_t)layer * kPairs + p] = out;
            }
        }
        zero_state = false;
    }
    for (int n = 0; n < N; ++n)
        post[n] = (int8_t)(L[pos_of_bit(n, K, R)] ^ 0x80);
    return left;
}

template <int C, bool U>
static int run_one(const LdpcTableDef& def, const Schedule& s, const int8_t* llr, int trials, int8_t* post)
{
    return run_frame<C, U>(def, s, llr, trials, post);
}

static int dispatch(const LdpcTableDef& def, const Schedule& s, const int8_t* llr, int trials, int8_t* post)
{
    const bool uniform = s.min_cnt == s.max_cnt;
    const int max_cnt = s.max_cnt;
#define CALL(C, U) return run_one<C, U>(def, s, llr, trials, post)
    if (uniform) {
        switch (max_cnt) {
        case 2: CALL(2, true);
        case 3: CALL(3, true);
        case 4: CALL(4, true);
        case 5: CALL(5, true);
        case 7: CALL(7, true);
        case 8: CALL(8, true);
        case 9: CALL(9, true);
        case 11: CALL(11, true);
        case 12: CALL(12, true);
        case 16: CALL(16, true);
        case 20: CALL(20, true);
        case 25: CALL(25, true);
        case 28: CALL(28, true);
        default: break;
        }
    }
    if (max_cnt <= 5) CALL(5, false);
    if (max_cnt <= 9) CALL(9, false);
    if (max_cnt <= 13) CALL(13, false);
    if (max_cnt <= 18) CALL(18, false);
    CALL(28, false);
#undef CALL
}

int main(int argc, char** argv)
{
    std::vector<int> tables;
    for (int i = 1; i < argc; ++i)
        tables.push_back(atoi(argv[i]));
    if (tables.empty())
        for (int t = 0; t < num_tables(); ++t)
            tables.push_back(t);
    const int trials = getenv("EMUL_TRIALS") ? atoi(getenv("EMUL_TRIALS")) : 6;
    const int frames = getenv("EMUL_FRAMES") ? atoi(getenv("EMUL_FRAMES")) : 2;
    int failures = 0;
    for (int t : tables) {
        const LdpcTableDef& def = *table_def(t);
        Schedule s;
        build_schedule(def, s);
        orc_ldpc* ol = orc_ldpc_create(t);
        std::mt19937 rng(1234 + t);
        const int N = def.N, K = def.K;
        const double rate = (double)K / N;
        int bad_frames = 0, conv = 0;
        for (int f = 0; f < frames; ++f) {
            // a noisy codeword near the code's waterfall (frame 0 a little below, the rest a little above), saturating LLRs included
            std::vector<uint8_t> msg(K), cw(N);
            for (int i = 0; i < K; ++i)
                msg[i] = rng() & 1;
            orc_ldpc_encode(ol, msg.data(), cw.data());
            const double ebn0_db = (f == 0 ? 0.3 : 1.6) + (rate > 0.7 ? 1.5 : 0.0);
            const double esn0 = pow(10.0, ebn0_db / 10.0) * rate * 2.0, n0 = 1.0 / esn0;
            std::normal_distribution<float> noise(0.f, (float)sqrt(n0 / 2));
            std::vector<int8_t> llr(N), post(N), want(N);
            const float scale = (f == frames - 1 ? 4.0f : 1.0f) * (float)(2.0 * sqrt(2.0) / n0); // last frame: heavy saturation
            for (int i = 0; i < N; ++i) {
                float v = rintf(((cw[i] ? -0.70710678f : 0.70710678f) + noise(rng)) * scale);
                llr[i] = (int8_t)(v > 127 ? 127 : v < -128 ? -128 : v);
            }
            want = llr;
            const int want_ret = orc_ldpc_decode(ol, want.data(), 1, trials);
            const int got_ret = dispatch(def, s, llr.data(), trials, post.data());
            int diff = 0, first_diff = -1;
            for (int i = 0; i < N; ++i)
                if (post[i] != want[i]) {
                    if (first_diff < 0)
                        first_diff = i;
                    ++diff;
                }
            if (diff || want_ret != got_ret) {
                ++bad_frames;
                printf("  table %d (%s) frame %d: %d posteriors differ (first at %d: got %d want %d), ret %d vs %d\n", t, def.name, f, diff,
                       first_diff, first_diff >= 0 ? post[first_diff] : 0, first_diff >= 0 ? want[first_diff] : 0, got_ret, want_ret);
            }
            conv += want_ret >= 0;
        }
        printf("table %2d %-22s N=%5d q=%3d cnt %d..%d conflict layers %2d: %s (%d/%d frames converged)\n", t, def.name, N, def.q, s.min_cnt,
               s.max_cnt, s.conflict_layers, bad_frames ? "FAIL" : "ok", conv, frames);
        fflush(stdout);
        failures += bad_frames;
        orc_ldpc_destroy(ol);
    }
    printf("%s\n", failures ? "FAILED" : "all ok");
    return failures ? 1 : 0;
}
