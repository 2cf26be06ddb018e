/*
 * dvbs2_oracle.c -- CPU restatement of the gr-dvbs2rx FEC decode path.
 *
 * TEST INFRASTRUCTURE ONLY (see dvbs2_oracle.h).  Plain scalar C, one FECFRAME lane at a
 * time; every function cites the reference lines it restates.  Paths are relative to the
 * reference checkout (igorauad/gr-dvbs2rx v1.4.0).
 */
#include "dvbs2_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------- */
/* code tables (generated data, shared with the product: tools/gen_code_tables.py)        */
/* ------------------------------------------------------------------------------------- */
typedef struct {
    const char* name;
    int N, K, q, n_circ, links_total, links_max_cn;
    const uint32_t* circ; /* layer << 17 | group << 9 | shift */
} Dvbs2LdpcTableDef;
typedef struct {
    int framesize, rate, standard, table, kbch, nbch, t;
} Dvbs2ModcodDef;
#include "../gr-dvbs2rx_b200/csrc/dvbs2_code_tables.inc"

#define NTABLES ((int)(sizeof(kLdpcTables) / sizeof(kLdpcTables[0])))
#define NMODCODS ((int)(sizeof(kModcods) / sizeof(kModcods[0])))

int orc_num_tables(void) { return NTABLES; }
const char* orc_table_name(int t) { return (t >= 0 && t < NTABLES) ? kLdpcTables[t].name : 0; }
int orc_table_n(int t) { return kLdpcTables[t].N; }
int orc_table_k(int t) { return kLdpcTables[t].K; }

/* lib/ldpc_decoder_bb_impl.cc:104-307 (table choice), lib/fec_params.cc:16-344 */
int orc_lookup(int standard, int framesize, int rate, int* kbch, int* nbch, int* t)
{
    for (int i = 0; i < NMODCODS; ++i) {
        const Dvbs2ModcodDef* m = &kModcods[i];
        if (m->framesize != framesize || m->rate != rate)
            continue;
        /* the reference tests `standard == STANDARD_DVBS2`, else takes the T2 table */
        if (m->standard >= 0 && (m->standard == 0) != (standard == 0))
            continue;
        if (kbch) *kbch = m->kbch;
        if (nbch) *nbch = m->nbch;
        if (t) *t = m->t;
        return m->table;
    }
    return -1;
}

/* ------------------------------------------------------------------------------------- */
/* LDPC                                                                                   */
/* ------------------------------------------------------------------------------------- */
struct orc_ldpc {
    int N, K, R, q, CNL, LT;
    uint16_t* pos; /* [R*CNL] data-bit index per link, check order (layer, j) */
    uint8_t* cnc;  /* [R] data links per check (original check order, as the reference) */
    int8_t* bnl;   /* [LT] stored check->variable messages of ONE lane */
    int8_t* pty;   /* [R] permuted parity posteriors of ONE lane */
    const Dvbs2LdpcTableDef* def;
};

/* lib/ldpc_decoder/layered_decoder.hh:101-142 (init) with the bit iteration of
 * lib/ldpc_decoder/ldpc.hh:67-78: data bit g*360+m accumulates into checks (x + q*m) mod R. */
orc_ldpc* orc_ldpc_create(int table)
{
    if (table < 0 || table >= NTABLES)
        return 0;
    const Dvbs2LdpcTableDef* d = &kLdpcTables[table];
    orc_ldpc* o = (orc_ldpc*)calloc(1, sizeof(*o));
    o->def = d;
    o->N = d->N;
    o->K = d->K;
    o->R = d->N - d->K;
    o->q = d->q;
    o->CNL = d->links_max_cn - 2;
    o->LT = d->links_total;
    const int M = 360, R = o->R, q = o->q, CNL = o->CNL;
    uint16_t* pos = (uint16_t*)calloc((size_t)R * CNL, sizeof(uint16_t));
    o->cnc = (uint8_t*)calloc(R, 1);
    /* visit data bits in ascending order so links of a check come out sorted by bit index */
    int ngroups = o->K / M;
    /* bucket circulants by group (they are sorted by layer in the table) */
    for (int g = 0; g < ngroups; ++g) {
        for (int m = 0; m < M; ++m) {
            int j = g * M + m;
            for (int c = 0; c < d->n_circ; ++c) {
                uint32_t w = d->circ[c];
                if ((int)((w >> 9) & 0xff) != g)
                    continue;
                int x = q * (int)(w & 0x1ff) + (int)(w >> 17);
                int i = (x + q * m) % R;
                pos[CNL * i + o->cnc[i]++] = (uint16_t)j;
            }
        }
    }
    /* layered_decoder.hh:135-141: reorder checks so (i, j) <- original check q*j + i */
    o->pos = (uint16_t*)calloc((size_t)R * CNL, sizeof(uint16_t));
    for (int i = 0; i < q; ++i)
        for (int j = 0; j < M; ++j)
            for (int c = 0; c < CNL; ++c)
                o->pos[CNL * (M * i + j) + c] = pos[CNL * (q * j + i) + c];
    free(pos);
    o->bnl = (int8_t*)malloc(o->LT);
    o->pty = (int8_t*)malloc(R);
    return o;
}

void orc_ldpc_destroy(orc_ldpc* o)
{
    if (!o)
        return;
    free(o->pos);
    free(o->cnc);
    free(o->bnl);
    free(o->pty);
    free(o);
}

/* int8 lane primitives: lib/ldpc_decoder/simd.hh (generic) :287-294 vqabs, :1011-1020 vqadd,
 * :1085-1094 vqsub (signed), :1105-1113 vqsub (unsigned), :1142-1150 vsign */
static inline int8_t qadd(int8_t a, int8_t b)
{
    int x = (int)a + (int)b;
    return (int8_t)(x < -128 ? -128 : x > 127 ? 127 : x);
}
static inline int8_t qsub(int8_t a, int8_t b)
{
    int x = (int)a - (int)b;
    return (int8_t)(x < -128 ? -128 : x > 127 ? 127 : x);
}
static inline int8_t qabs(int8_t a)
{
    int x = a < -127 ? -127 : a;
    return (int8_t)(x < 0 ? -x : x);
}
static inline uint8_t usub(uint8_t a, uint8_t b) { return (uint8_t)(a > b ? a - b : 0); }
static inline int8_t vsign(int8_t a, int8_t b) { return (int8_t)(b > 0 ? a : b < 0 ? -a : 0); }

/* lib/ldpc_decoder/algorithms.hh:170-192 finalp, offset min-sum, beta = nearbyint(0.5*2) = 1 */
static void finalp(int8_t* links, int cnt)
{
    int8_t mags[64];
    for (int i = 0; i < cnt; ++i)
        mags[i] = (int8_t)usub((uint8_t)qabs(links[i]), 1);
    int8_t m0 = mags[0] < mags[1] ? mags[0] : mags[1];
    int8_t m1 = mags[0] < mags[1] ? mags[1] : mags[0];
    for (int i = 2; i < cnt; ++i) {
        int8_t mx = m0 > mags[i] ? m0 : mags[i];
        m1 = m1 < mx ? m1 : mx;
        m0 = m0 < mags[i] ? m0 : mags[i];
    }
    uint8_t signs = (uint8_t)links[0];
    for (int i = 1; i < cnt; ++i)
        signs ^= (uint8_t)links[i];
    for (int i = 0; i < cnt; ++i) {
        int8_t other = (mags[i] == m0) ? m1 : m0;                       /* :166-169 */
        int8_t s = (int8_t)((uint8_t)(signs ^ (uint8_t)links[i]) | 127); /* +127 or -1 */
        links[i] = vsign(other, s);
    }
}

/* lib/ldpc_decoder/layered_decoder.hh:32-49 for one lane: 1 if some check is unsatisfied */
static int bad_lane(const orc_ldpc* o, const int8_t* data, const int8_t* parity)
{
    const int M = 360, q = o->q, CNL = o->CNL;
    for (int i = 0; i < q; ++i) {
        int cnt = o->cnc[i];
        for (int j = 0; j < M; ++j) {
            int8_t cnv = vsign(1, parity[M * i + j]);
            if (i)
                cnv = vsign(cnv, parity[M * (i - 1) + j]);
            else if (j)
                cnv = vsign(cnv, parity[j + (q - 1) * M - 1]);
            for (int c = 0; c < cnt; ++c)
                cnv = vsign(cnv, data[o->pos[CNL * (M * i + j) + c]]);
            if (!(cnv > 0)) /* algorithms.hh:195-202: lane is bad unless cnv > 0 */
                return 1;
        }
    }
    return 0;
}

/* lib/ldpc_decoder/layered_decoder.hh:50-79 for one lane */
static void update_lane(const orc_ldpc* o, int8_t* data, int8_t* parity, int8_t* bnl)
{
    const int M = 360, q = o->q, CNL = o->CNL;
    int8_t* bl = bnl;
    for (int i = 0; i < q; ++i) {
        int cnt = o->cnc[i];
        for (int j = 0; j < M; ++j) {
            int deg = cnt + 2 - !(i | j);
            int8_t inp[64], out[64];
            const uint16_t* p = &o->pos[CNL * (M * i + j)];
            for (int c = 0; c < cnt; ++c)
                inp[c] = out[c] = qsub(data[p[c]], bl[c]);
            inp[cnt] = out[cnt] = qsub(parity[M * i + j], bl[cnt]);
            if (i)
                inp[cnt + 1] = out[cnt + 1] = qsub(parity[M * (i - 1) + j], bl[cnt + 1]);
            else if (j)
                inp[cnt + 1] = out[cnt + 1] = qsub(parity[j + (q - 1) * M - 1], bl[cnt + 1]);
            finalp(out, deg);
            for (int c = 0; c < cnt; ++c)
                data[p[c]] = qadd(inp[c], out[c]);
            parity[M * i + j] = qadd(inp[cnt], out[cnt]);
            if (i)
                parity[M * (i - 1) + j] = qadd(inp[cnt + 1], out[cnt + 1]);
            else if (j)
                parity[j + (q - 1) * M - 1] = qadd(inp[cnt + 1], out[cnt + 1]);
            for (int d = 0; d < deg; ++d) { /* algorithms.hh:203-206: clamp to [-32, 31] */
                int8_t v = out[d];
                *bl++ = v < -32 ? -32 : v > 31 ? 31 : v;
            }
        }
    }
}

int orc_ldpc_bad(orc_ldpc* o, const int8_t* code, int lanes)
{
    const int M = 360, q = o->q, K = o->K, N = o->N;
    for (int l = 0; l < lanes; ++l) {
        const int8_t* data = code + (size_t)l * N;
        for (int i = 0; i < q; ++i)
            for (int j = 0; j < M; ++j)
                o->pty[M * i + j] = data[K + q * j + i];
        if (bad_lane(o, data, o->pty))
            return 1;
    }
    return 0;
}

/* lib/ldpc_decoder/layered_decoder.hh:143-160.  Lanes are independent inside an iteration;
 * only the loop condition couples them ("any lane bad"), so lanes are stepped one at a time
 * with per-lane message/parity state. */
int orc_ldpc_decode(orc_ldpc* o, int8_t* code, int lanes, int trials)
{
    const int M = 360, q = o->q, K = o->K, N = o->N, R = o->R;
    int8_t* bnl = (int8_t*)calloc((size_t)lanes * o->LT, 1); /* reset(): :27-31 */
    int8_t* pty = (int8_t*)malloc((size_t)lanes * R);
    for (int l = 0; l < lanes; ++l)
        for (int i = 0; i < q; ++i)
            for (int j = 0; j < M; ++j) /* :150-152 */
                pty[(size_t)l * R + M * i + j] = code[(size_t)l * N + K + q * j + i];
    for (;;) {
        int bad = 0;
        for (int l = 0; l < lanes && !bad; ++l)
            bad = bad_lane(o, code + (size_t)l * N, pty + (size_t)l * R);
        if (!(bad && --trials >= 0)) /* :153 */
            break;
        for (int l = 0; l < lanes; ++l)
            update_lane(o, code + (size_t)l * N, pty + (size_t)l * R, bnl + (size_t)l * o->LT);
    }
    for (int l = 0; l < lanes; ++l)
        for (int i = 0; i < q; ++i)
            for (int j = 0; j < M; ++j) /* :155-157 */
                code[(size_t)l * N + K + q * j + i] = pty[(size_t)l * R + M * i + j];
    free(bnl);
    free(pty);
    return trials;
}

/* IRA encoder of EN 302 307-1 clause 5.3.2 (the reference has no encoder; gr-dtv does). */
void orc_ldpc_encode(orc_ldpc* o, const uint8_t* msg, uint8_t* cw)
{
    const int M = 360, q = o->q, K = o->K, R = o->R;
    memcpy(cw, msg, K);
    uint8_t* p = cw + K;
    memset(p, 0, R);
    for (int c = 0; c < o->def->n_circ; ++c) {
        uint32_t w = o->def->circ[c];
        int g = (w >> 9) & 0xff;
        int x = q * (int)(w & 0x1ff) + (int)(w >> 17);
        for (int m = 0; m < M; ++m)
            p[(x + q * m) % R] ^= msg[g * M + m];
    }
    for (int i = 1; i < R; ++i)
        p[i] ^= p[i - 1];
}

/* lib/ldpc_decoder_bb_impl.cc:432-442 */
void orc_pack_hard(const int8_t* llr, int nbits, uint8_t* out)
{
    for (int j = 0; j < nbits / 8; ++j) {
        uint8_t b = 0;
        for (int k = 0; k < 8; ++k)
            if (llr[j * 8 + k] < 0)
                b |= (uint8_t)(1 << (7 - k));
        out[j] = b;
    }
}

/* ------------------------------------------------------------------------------------- */
/* GF(2^m) and BCH                                                                        */
/* ------------------------------------------------------------------------------------- */
#define BCH_MAX_T 12
#define BCH_MAX_DEG 200

struct orc_bch {
    int m, t, n, k, s, gdeg;
    uint32_t nz;        /* 2^m - 1 */
    uint32_t* antilog;  /* alpha^i, i in [0, 2^m-1)       lib/gf.cc:46-58 */
    uint32_t* log;      /* exponent of a non-zero element lib/gf.cc:60-62 */
    uint32_t* quad_lut; /* lib/bch.cc:107-112 */
    uint8_t g[BCH_MAX_DEG + 1];
};

static inline uint32_t gf_alpha(const orc_bch* b, uint32_t i) { return b->antilog[i % b->nz]; } /* gf.h:99 */
static inline uint32_t gf_mul(const orc_bch* b, uint32_t x, uint32_t y)                        /* gf.cc:69-75 */
{
    if (!x || !y)
        return 0;
    return gf_alpha(b, b->log[x] + b->log[y]);
}
static inline uint32_t gf_inv(const orc_bch* b, uint32_t x) { return gf_alpha(b, b->nz - b->log[x]); } /* gf.cc:77-85 */
static inline uint32_t gf_div(const orc_bch* b, uint32_t x, uint32_t y) { return gf_mul(b, x, gf_inv(b, y)); }

uint32_t orc_gf_alpha(const orc_bch* b, uint32_t i) { return gf_alpha(b, i); }

/* lib/gf.cc get_conjugates / get_min_poly: phi(x) = prod over distinct conjugates (x + beta^(2^l)) */
uint32_t orc_gf_min_poly(const orc_bch* b, uint32_t i)
{
    uint32_t conj[32];
    int nc = 0;
    uint32_t e = i % b->nz;
    for (int l = 0; l < b->m; ++l) {
        int seen = 0;
        for (int c = 0; c < nc; ++c)
            seen |= (conj[c] == e);
        if (seen)
            break;
        conj[nc++] = e;
        e = (uint32_t)(((uint64_t)e * 2) % b->nz);
    }
    uint32_t poly[34];
    memset(poly, 0, sizeof(poly));
    poly[0] = 1;
    int deg = 0;
    for (int c = 0; c < nc; ++c) {
        uint32_t beta = gf_alpha(b, conj[c]);
        for (int d = deg + 1; d >= 1; --d)
            poly[d] = poly[d - 1] ^ gf_mul(b, poly[d], beta);
        poly[0] = gf_mul(b, poly[0], beta);
        ++deg;
    }
    uint32_t mask = 0;
    for (int d = 0; d <= deg; ++d)
        if (poly[d])
            mask |= 1u << d; /* coefficients are 0/1 for a minimal polynomial */
    return mask;
}

/* lib/gf.cc:19-67 field tables; lib/bch.cc:36-62 generator; :64-113 codec parameters */
orc_bch* orc_bch_create(uint32_t prim_poly, int t, int n)
{
    int m = 31;
    while (m > 0 && !(prim_poly >> m))
        --m;
    if (m < 2 || m > 16 || t < 1 || t > BCH_MAX_T)
        return 0;
    orc_bch* b = (orc_bch*)calloc(1, sizeof(*b));
    b->m = m;
    b->t = t;
    b->nz = (1u << m) - 1;
    b->antilog = (uint32_t*)malloc(sizeof(uint32_t) * b->nz);
    b->log = (uint32_t*)calloc((size_t)1 << m, sizeof(uint32_t));
    uint32_t low = prim_poly ^ (1u << m);
    uint32_t v = 1;
    for (uint32_t i = 0; i < b->nz; ++i) {
        b->antilog[i] = v;
        b->log[v] = i;
        v = ((v << 1) & b->nz) ^ ((v >> (m - 1)) * low);
    }
    /* g(x) = product of the distinct minimal polynomials of alpha^(2i+1), i < t */
    memset(b->g, 0, sizeof(b->g));
    b->g[0] = 1;
    b->gdeg = 0;
    uint32_t done[BCH_MAX_T];
    int ndone = 0;
    for (int i = 0; i < t; ++i) {
        uint32_t mp = orc_gf_min_poly(b, (uint32_t)(2 * i + 1));
        int dup = 0;
        for (int d = 0; d < ndone; ++d)
            dup |= (done[d] == mp);
        if (dup)
            continue;
        done[ndone++] = mp;
        uint8_t prod[BCH_MAX_DEG + 1];
        memset(prod, 0, sizeof(prod));
        for (int a = 0; a <= b->gdeg; ++a)
            if (b->g[a])
                for (int c = 0; c <= m; ++c)
                    if ((mp >> c) & 1)
                        prod[a + c] ^= 1;
        int md = m;
        while (!((mp >> md) & 1))
            --md;
        b->gdeg += md;
        memcpy(b->g, prod, sizeof(prod));
    }
    b->n = n ? n : (int)b->nz;
    b->s = (int)b->nz - b->n;
    b->k = b->n - b->gdeg;
    b->quad_lut = (uint32_t*)calloc((size_t)1 << m, sizeof(uint32_t));
    for (uint32_t r = 0; r < (1u << m); ++r)
        b->quad_lut[gf_mul(b, r, r) ^ r] = r;
    return b;
}

void orc_bch_destroy(orc_bch* b)
{
    if (!b)
        return;
    free(b->antilog);
    free(b->log);
    free(b->quad_lut);
    free(b);
}
int orc_bch_n(const orc_bch* b) { return b->n; }
int orc_bch_k(const orc_bch* b) { return b->k; }
int orc_bch_genpoly(const orc_bch* b, uint8_t* g, int cap)
{
    for (int i = 0; i <= b->gdeg && i < cap; ++i)
        g[i] = b->g[i];
    return b->gdeg;
}

/* remainder of bits[0..nbits) (first bit = highest power) by g(x); rem[i] = coef of x^i.
 * Same value as the byte-LUT of lib/gf_util.h:219-262, computed bit-serially. */
static void poly_rem(const orc_bch* b, const uint8_t* bytes, int nbits, uint8_t* rem)
{
    int gd = b->gdeg;
    memset(rem, 0, gd);
    for (int i = 0; i < nbits; ++i) {
        uint8_t in = (bytes[i >> 3] >> (7 - (i & 7))) & 1;
        uint8_t fb = rem[gd - 1]; /* coefficient leaving at x^gd */
        for (int d = gd - 1; d > 0; --d)
            rem[d] = rem[d - 1] ^ (fb & b->g[d]);
        rem[0] = in ^ (fb & b->g[0]);
    }
}

/* lib/bch.cc:157-173 */
void orc_bch_encode(const orc_bch* b, const uint8_t* msg, uint8_t* cw)
{
    int kb = b->k / 8, nb = b->n / 8;
    memcpy(cw, msg, kb);
    memset(cw + kb, 0, nb - kb);
    uint8_t rem[BCH_MAX_DEG];
    poly_rem(b, cw, b->n, rem);
    for (int i = 0; i < b->gdeg; ++i)
        if (rem[i]) {
            int bit = b->n - 1 - i;
            cw[bit >> 3] |= (uint8_t)(1 << (7 - (bit & 7)));
        }
}

/* lib/bch.cc:216-222, 175-189: S_i = s(alpha^i), s = r mod g; empty (return 0) when s == 0 */
int orc_bch_syndrome(const orc_bch* b, const uint8_t* cw, uint32_t* synd)
{
    uint8_t rem[BCH_MAX_DEG];
    poly_rem(b, cw, b->n, rem);
    int any = 0;
    for (int i = 0; i < b->gdeg; ++i)
        any |= rem[i];
    if (!any)
        return 0;
    for (int i = 1; i <= 2 * b->t; ++i) { /* gf.cc:278-287 eval_by_exp */
        uint32_t acc = 0;
        for (int j = 0; j < b->gdeg; ++j)
            if (rem[j])
                acc ^= gf_alpha(b, (uint32_t)i * (uint32_t)j);
        synd[i - 1] = acc;
    }
    return 1;
}

static int poly_degree(const uint32_t* p, int len)
{
    int d = len - 1;
    while (d >= 0 && p[d] == 0)
        --d;
    return d; /* -1 for the zero polynomial, as gf2m_poly after stripping (gf.cc:178-186) */
}

/* lib/bch.cc:224-304: simplified Berlekamp, Lin & Costello table form.
 * sigma out: coefficients [0..t+1]; returns the degree. */
int orc_bch_err_loc_poly(const orc_bch* b, const uint32_t* S, uint32_t* sigma_out)
{
    enum { W = 2 * BCH_MAX_T + 4 };
    const int t = b->t;
    uint32_t sig[BCH_MAX_T + 3][W];
    int deg[BCH_MAX_T + 3];
    int two_mu[BCH_MAX_T + 3]; /* 2*mu per row: -1, 0, 2, 4, ... */
    uint32_t d[BCH_MAX_T + 3];
    memset(sig, 0, sizeof(sig));
    two_mu[0] = -1;
    for (int i = 0; i < t + 1; ++i)
        two_mu[i + 1] = 2 * i;
    sig[0][0] = 1;
    sig[1][0] = 1;
    sig[2][0] = 1;
    sig[2][1] = S[0];
    deg[0] = 0;
    deg[1] = 0;
    deg[2] = poly_degree(sig[2], W);
    d[0] = 1;
    d[1] = S[0];
    int row = 2;
    while (row <= t) {
        int tm = two_mu[row]; /* 2*mu */
        d[row] = S[tm];
        for (int j = 1; j <= deg[row]; ++j)
            if (sig[row][j])
                d[row] ^= gf_mul(b, sig[row][j], S[tm - j]);
        if (d[row] == 0) {
            memcpy(sig[row + 1], sig[row], sizeof(sig[row]));
        } else {
            int row_rho = 0, max_diff = -2;
            for (int j = row - 1; j >= 0; --j) /* latest row wins ties (strict >) */
                if (d[j] != 0) {
                    int diff = two_mu[j] - deg[j];
                    if (diff > max_diff) {
                        max_diff = diff;
                        row_rho = j;
                    }
                }
            uint32_t coef = gf_div(b, d[row], d[row_rho]);
            int shift = tm - two_mu[row_rho]; /* int(2*(mu - rho)) */
            memcpy(sig[row + 1], sig[row], sizeof(sig[row]));
            for (int j = 0; j <= deg[row_rho]; ++j)
                if (j + shift < W)
                    sig[row + 1][j + shift] ^= gf_mul(b, coef, sig[row_rho][j]);
        }
        deg[row + 1] = poly_degree(sig[row + 1], W);
        ++row;
    }
    for (int j = 0; j < t + 2; ++j)
        sigma_out[j] = sig[row][j];
    return deg[row];
}

/* lib/bch.cc:306-385.  Returns the number of locators found, or -2 where the reference
 * would throw out of gf.h:110 (unsolvable quadratic -> inverse(0)). */
int orc_bch_err_loc_numbers(const orc_bch* b, const uint32_t* sigma, int deg, uint32_t* numbers)
{
    if (deg > b->t)
        return 0;
    if (deg == 1) {
        numbers[0] = gf_div(b, sigma[1], sigma[0]);
        return 1;
    }
    if (deg == 2) {
        if (sigma[1] == 0 || sigma[0] == 0)
            return 0;
        uint32_t b_over_a = gf_div(b, sigma[1], sigma[2]);
        uint32_t rr = gf_div(b, gf_mul(b, sigma[0], sigma[2]), gf_mul(b, sigma[1], sigma[1]));
        uint32_t r = b->quad_lut[rr];
        uint32_t x0 = gf_mul(b, r, b_over_a);
        uint32_t x1 = gf_mul(b, b_over_a, r ^ 1);
        if (x0 == 0 || x1 == 0)
            return -2;
        numbers[0] = gf_inv(b, x0);
        numbers[1] = gf_inv(b, x1);
        return 2;
    }
    /* lib/gf.cc:289-404: Chien search over exponents s+1 .. n+s, stop after `deg` roots */
    int nfound = 0;
    for (uint32_t e = (uint32_t)b->s + 1; e <= (uint32_t)(b->n + b->s); ++e) {
        uint32_t res = 0;
        for (int j = 0; j <= deg; ++j)
            if (sigma[j])
                res ^= gf_alpha(b, b->log[sigma[j]] + e * (uint32_t)j);
        if (res == 0) {
            numbers[nfound++] = gf_alpha(b, b->nz - e); /* inverse_by_exp */
            if (nfound == deg)
                break;
        }
    }
    return nfound;
}

/* lib/bch.cc:467-487 with correct_errors(u8) :428-452 */
int orc_bch_decode(const orc_bch* b, const uint8_t* cw, uint8_t* msg)
{
    memcpy(msg, cw, b->k / 8);
    uint32_t S[2 * BCH_MAX_T];
    if (!orc_bch_syndrome(b, cw, S))
        return 0;
    uint32_t sigma[BCH_MAX_T + 2];
    int deg = orc_bch_err_loc_poly(b, S, sigma);
    uint32_t numbers[BCH_MAX_T + 2];
    int nn = orc_bch_err_loc_numbers(b, sigma, deg, numbers);
    /* Two places where the reference throws out of decode() instead of returning (both need
     * > t errors that imitate a 1- or 2-error syndrome): an unsolvable quadratic reaches
     * inverse(0) (lib/gf.h:110), and a closed-form locator outside the shortened code hits
     * "Error location number out of range" (lib/bch.cc:441).  The oracle defines both as
     * "-1, nothing flipped". */
    if (nn < 0)
        return -1;
    for (int i = 0; i < nn; ++i)
        if (b->log[numbers[i]] >= (uint32_t)b->n)
            return -1;
    for (int i = 0; i < nn; ++i) {
        uint32_t bit_idx = b->log[numbers[i]];
        if (bit_idx < (uint32_t)(b->n - b->k))
            continue;
        uint32_t net = (uint32_t)b->n - 1 - bit_idx;
        msg[net >> 3] ^= (uint8_t)(1u << (7 - (net & 7)));
    }
    return deg == nn ? nn : -1;
}

/* ------------------------------------------------------------------------------------- */
/* soft demapper                                                                          */
/* ------------------------------------------------------------------------------------- */
/* VOLK generic volk_32f_s32f_convert_8i (third-party, not vendored; gnuradio/volk 3.x):
 * r = in*scalar; r > 127 -> 127; r < -128 -> -128; else (int8) rintf(r). */
static inline int8_t convert_8i(float r)
{
    if (r > 127.0f)
        return 127;
    if (r < -128.0f)
        return -128;
    return (int8_t)rintf(r);
}

/* lib/qpsk.h:208-214 */
void orc_demap_qpsk(const float* iq, int n_syms, float n0, int8_t* llr)
{
    float scalar = (float)(2 * M_SQRT2 / n0);
    for (int i = 0; i < 2 * n_syms; ++i)
        llr[i] = convert_8i(iq[i] * scalar);
}

/* lib/psk.hh:123-131 */
static inline int8_t quantize8(float dist, float precision, float value)
{
    value *= dist * precision;
    value = nearbyintf(value);
    value = fminf(fmaxf(value, -128.0f), 127.0f);
    return (int8_t)value;
}

/* lib/psk.hh:143-150 + lib/xfecframe_demapper_cb_impl.cc:48-69,148,155-176.
 * Build with -ffp-contract=off: the complex rotation is 4 mul + 2 add, unfused. */
void orc_demap_8psk(const float* iq, int n_syms, float n0, int rate, int8_t* llr)
{
    const float rcp_sqrt_2 = 0.70710678118654752440f;
    const float dist = 2 * 0.38268343236508977173f;
    const float rot_re = (float)cos(-M_PI / 8), rot_im = (float)sin(-M_PI / 8);
    float precision = (float)(4.0 / n0);
    int rows = n_syms, r0, r1, r2;
    if (rate == 4) { /* C3_5: 2-1-0 */
        r0 = rows * 2; r1 = rows; r2 = 0;
    } else if (rate == 26 || rate == 28 || rate == 38 || rate == 39 || rate == 19) {
        r0 = rows; r1 = 0; r2 = rows * 2; /* C25_36 C13_18 C7_15 C8_15 C26_45: 1-0-2 */
    } else {
        r0 = 0; r1 = rows; r2 = rows * 2;
    }
    for (int j = 0; j < n_syms; ++j) {
        float a = iq[2 * j], b = iq[2 * j + 1];
        float re = a * rot_re - b * rot_im;
        float im = a * rot_im + b * rot_re;
        int8_t b1 = quantize8(dist, precision, re);
        int8_t b2 = quantize8(dist, precision, im);
        int8_t b0 = quantize8(dist, precision, rcp_sqrt_2 * (fabsf(re) - fabsf(im)));
        llr[r0 + j] = b0;
        llr[r1 + j] = b1;
        llr[r2 + j] = b2;
    }
}

/* ------------------------------------------------------------------------------------- */
/* SNR estimates of the demapper block -- tolerance-checked, not bit-exact:
 * the reference sums with VOLK dot products whose order is unspecified (lib/qpsk.h:41-65).   */
/* ------------------------------------------------------------------------------------- */
/* QPSK: reference points by slicing the symbols (llr == NULL, lib/qpsk.h:171-181,240-244: a
 * component >= 0 maps to +sqrt(2)/2) or the posterior LLRs (lib/qpsk.h:267-281).  Es/N0 linear. */
float orc_snr_qpsk(const float* iq, int n_syms, const int8_t* llr)
{
    const float a = 0.70710678118654752440f;
    double sp = 0, np = 0;
    for (int i = 0; i < 2 * n_syms; ++i) {
        const int pos = llr ? (llr[i] >= 0) : (iq[i] >= 0.0f);
        const float r = pos ? a : -a;
        const float e = iq[i] - r;
        sp += (double)r * r;
        np += (double)e * e;
    }
    if (np == 0)
        np = 1e-12;
    return (float)(sp / np);
}

/* 8PSK: lib/xfecframe_demapper_cb_impl.cc:128-142 (hard slice, lib/psk.hh:135-141) and :267-302
 * (posterior LLR signs through the deinterleaver rows); constellation lib/psk.hh:114-121,152-157. */
float orc_snr_8psk(const float* iq, int n_syms, const int8_t* llr, int rate)
{
    static const float m8[8][2] = { { 0.70710678118654752440f, 0.70710678118654752440f }, { 1, 0 }, { -1, 0 },
                                    { -0.70710678118654752440f, -0.70710678118654752440f }, { 0, 1 },
                                    { 0.70710678118654752440f, -0.70710678118654752440f },
                                    { -0.70710678118654752440f, 0.70710678118654752440f }, { 0, -1 } };
    const float rot_re = (float)cos(-M_PI / 8), rot_im = (float)sin(-M_PI / 8);
    int rows = n_syms, r0, r1, r2;
    if (rate == 4) {
        r0 = rows * 2; r1 = rows; r2 = 0;
    } else if (rate == 26 || rate == 28 || rate == 38 || rate == 39 || rate == 19) {
        r0 = rows; r1 = 0; r2 = rows * 2;
    } else {
        r0 = 0; r1 = rows; r2 = rows * 2;
    }
    float sp = 0, np = 0;
    for (int j = 0; j < n_syms; ++j) {
        const float a = iq[2 * j], b = iq[2 * j + 1];
        int b0, b1, b2; /* 1 = negative */
        if (llr) {
            b0 = llr[r0 + j] < 0; b1 = llr[r1 + j] < 0; b2 = llr[r2 + j] < 0;
        } else {
            const float re = a * rot_re - b * rot_im, im = a * rot_im + b * rot_re;
            b1 = re < 0; b2 = im < 0; b0 = fabsf(re) < fabsf(im);
        }
        const float* s = m8[4 * b0 + 2 * b1 + b2];
        const float er = a - s[0], ei = b - s[1];
        sp += s[0] * s[0] + s[1] * s[1];
        np += er * er + ei * ei;
    }
    if (!(np > 0))
        np = 1e-12f;
    return sp / np;
}

/* ------------------------------------------------------------------------------------- */
/* BB layer: descrambler and deheader                                   */
/* ------------------------------------------------------------------------------------- */
/* lib/bbdescrambler_bb_impl.cc:51-65: the PRBS 1 + x^14 + x^15, register loaded with
 * 100101010000000, one bit per BBFRAME bit, MSB first inside a byte. */
void orc_bb_prbs(uint8_t* seq, int nbytes)
{
    memset(seq, 0, (size_t)nbytes);
    int sr = 0x4A80;
    for (int i = 0; i < nbytes * 8; i++) {
        int b = ((sr) ^ (sr >> 1)) & 1;
        seq[i / 8] |= (uint8_t)(b << (7 - (i % 8)));
        sr >>= 1;
        if (b)
            sr |= 0x4000;
    }
}

/* lib/bbdescrambler_bb_impl.cc:67-82: every BBFRAME of kbch/8 bytes XORed with the same sequence */
void orc_bb_descramble(const uint8_t* in, int frames, int kbch_bytes, uint8_t* out)
{
    uint8_t* seq = (uint8_t*)malloc((size_t)kbch_bytes);
    orc_bb_prbs(seq, kbch_bytes);
    for (int f = 0; f < frames; ++f)
        for (int j = 0; j < kbch_bytes; ++j)
            out[(size_t)f * kbch_bytes + j] = in[(size_t)f * kbch_bytes + j] ^ seq[j];
    free(seq);
}

/* CRC-8 with g(x) = x^8+x^7+x^6+x^4+x^2+1 (lib/bbdeheader_bb_impl.cc:55).  The reference takes the
 * remainder of the byte string itself (lib/gf_util.h:219-262, no x^8 padding) and tests it for zero
 * (lib/bbdeheader_bb_impl.cc:138-142); since g(0) = 1, rem(y) = 0 <=> rem(y * x^8) = 0, which is
 * what the byte-wise table recursion below yields. */
static uint8_t crc8_tab[256];
static int crc8_ready = 0;
static void crc8_init(void)
{
    for (int i = 0; i < 256; ++i) {
        uint32_t r = (uint32_t)i << 8;
        for (int b = 15; b >= 8; --b)
            if (r & (1u << b))
                r ^= 0x1D5u << (b - 8);
        crc8_tab[i] = (uint8_t)r;
    }
    crc8_ready = 1;
}
uint8_t orc_crc8(const uint8_t* in, int size)
{
    if (!crc8_ready)
        crc8_init();
    uint8_t c = 0;
    for (int i = 0; i < size; ++i)
        c = crc8_tab[c ^ in[i]];
    return c;
}

struct orc_bbdeheader {
    unsigned kbch_bytes, max_dfl; /* max_dfl in bits */
    int synched;
    unsigned partial_ts_bytes;
    uint8_t partial_pkt[188];
    uint64_t packet_cnt, error_cnt, bbframe_cnt, bbframe_drop_cnt, bbframe_gap_cnt;
};

/* lib/bbdeheader_bb_impl.cc:40-61 */
orc_bbdeheader* orc_bbdeheader_create(int kbch)
{
    orc_bbdeheader* h = (orc_bbdeheader*)calloc(1, sizeof(*h));
    h->kbch_bytes = (unsigned)kbch / 8;
    h->max_dfl = (unsigned)kbch - 80;
    return h;
}
void orc_bbdeheader_destroy(orc_bbdeheader* h) { free(h); }
void orc_bbdeheader_counters(const orc_bbdeheader* h, uint64_t* out5)
{
    out5[0] = h->packet_cnt;
    out5[1] = h->error_cnt;
    out5[2] = h->bbframe_cnt;
    out5[3] = h->bbframe_drop_cnt;
    out5[4] = h->bbframe_gap_cnt;
}

/* lib/bbdeheader_bb_impl.cc:76-136: integrity check, field extraction, validation */
static int bb_parse_header(const orc_bbdeheader* h, const uint8_t* in, unsigned* dfl, unsigned* syncd)
{
    if (orc_crc8(in, 10) != 0)
        return 0;
    unsigned upl = ((unsigned)in[2] << 8) | in[3];
    *dfl = ((unsigned)in[4] << 8) | in[5];
    *syncd = ((unsigned)in[7] << 8) | in[8];
    if (*dfl > h->max_dfl)
        return 0;
    if (*dfl % 8 != 0)
        return 0;
    if (*syncd > *dfl)
        return 0;
    if (upl != 188 * 8)
        return 0;
    if (*syncd % 8 != 0)
        return 0;
    return 1;
}

/* lib/bbdeheader_bb_impl.cc:144-261 for `frames` whole BBFRAMEs (descrambled); returns bytes produced.
 * One deviation, where the reference has undefined behaviour: on a re-synchronisation with
 * syncd/8 + 1 > dfl/8 its unsigned `df_remaining` wraps and it reads far past the BBFRAME
 * (:203-209); here such a frame yields nothing (df_remaining = 0). */
long orc_bbdeheader_work(orc_bbdeheader* h, const uint8_t* in, int frames, uint8_t* out)
{
    long produced = 0;
    for (int i = 0; i < frames; ++i) {
        const uint8_t* frame = in + (size_t)i * h->kbch_bytes;
        unsigned dfl = 0, syncd = 0;
        const int valid = bb_parse_header(h, frame, &dfl, &syncd);
        h->bbframe_cnt++;
        if (!valid) {
            h->synched = 0;
            h->bbframe_drop_cnt++;
            continue;
        }
        const uint8_t* p = frame + 10;
        unsigned df_remaining = dfl / 8;
        if (h->partial_ts_bytes > 0 && (syncd / 8) != 188 - 1 - h->partial_ts_bytes) {
            h->synched = 0;
            h->bbframe_gap_cnt++;
        }
        if (!h->synched) {
            unsigned skip = syncd / 8 + 1;
            if (skip > df_remaining)
                skip = df_remaining; /* reference: unsigned wrap, out-of-bounds reads */
            p += skip;
            df_remaining -= skip;
            h->synched = 1;
            h->partial_ts_bytes = 0;
        }
        while (df_remaining >= 188) {
            const uint8_t* packet;
            if (h->partial_ts_bytes > 0) {
                unsigned remaining = 188 - h->partial_ts_bytes;
                memcpy(h->partial_pkt + h->partial_ts_bytes, p, remaining);
                h->partial_ts_bytes = 0;
                p += remaining;
                df_remaining -= remaining;
                packet = h->partial_pkt;
            } else {
                packet = p;
                p += 188;
                df_remaining -= 188;
            }
            const int crc_valid = orc_crc8(packet, 188) == 0;
            out[0] = 0x47;
            memcpy(out + 1, packet, 187);
            if (!crc_valid) {
                out[1] |= 0x80;
                h->error_cnt++;
            }
            out += 188;
            produced += 188;
            h->packet_cnt++;
        }
        if (df_remaining > 0) {
            h->partial_ts_bytes = df_remaining;
            memcpy(h->partial_pkt, p, df_remaining);
        }
    }
    return produced;
}

/* ---- PL descrambler + pilot-segment de-rotation ----------------------------------------
 * lib/pl_descrambler.cc:36-98   complex scrambling sequence of a Gold code: R_n in 0..3, descrambling factor
 *                               conj(exp(j R_n pi/2)) = {1, -j, -1, +j}[R_n]
 * lib/plsync_cc_impl.cc:639-802 handle_payload: descramble the whole payload (pilot blocks included in the index),
 *                               then de-rotate the data slots: phase starts at -plheader_phase, advances by
 *                               -2 pi fine_foffset per symbol (only when coarse corrected), and is reset to
 *                               -pilot_phase[b - 1] at the start of every 16-slot segment b >= 1 (when coarse
 *                               corrected); pilot blocks (36 symbols after every 16 slots) are dropped.
 * The de-rotation is VOLK's volk_32fc_s32fc_x2_rotator_32fc (third party, absent, version unpinned by the reference);
 * restated in its generic form: out = in * phase; phase *= inc; the phase re-normalised every 512 samples.  The
 * serial float recurrence is not bit-reproducible by a parallel evaluation: parity of the de-rotation is to tolerance.
 * The R_n sequence is pinned against lib/pl_descrambler.cc compiled unmodified (oracle/ref_harness.cc: ref_pl_rn). */
static int pl_parity18(long a, long b)
{
    int c = 0;
    a &= b;
    for (int i = 0; i < 18; i++)
        c += (int)((a >> i) & 1);
    return c & 1;
}
void orc_pl_rn(int gold_code, uint8_t* rn, int n)
{
    long x = 0x00001, y = 0x3FFFF;
    for (int k = 0; k < gold_code; k++) { /* the x register advanced by the Gold code */
        const int xb = pl_parity18(x, 0x0081);
        x >>= 1;
        if (xb)
            x |= 0x20000;
    }
    for (int i = 0; i < n; i++) {
        const int xa = pl_parity18(x, 0x8050), xb = pl_parity18(x, 0x0081), xc = (int)(x & 1);
        x >>= 1;
        if (xb)
            x |= 0x20000;
        const int ya = pl_parity18(y, 0x04A1), yb = pl_parity18(y, 0xFF60), yc = (int)(y & 1);
        y >>= 1;
        if (ya)
            y |= 0x20000;
        rn[i] = (uint8_t)((((xa ^ yb) & 1) << 1) | ((xc ^ yc) & 1));
    }
}
/* one PLFRAME payload [n_slots * 90 + n_pilots * 36][2] -> XFECFRAME symbols [n_slots * 90][2] */
void orc_pl_payload(const float* payload, int n_slots, int has_pilots, const uint8_t* rn, float plheader_phase, float fine_foffset,
                    int coarse_corrected, const float* pilot_phase, float* out)
{
    const float phase_inc = coarse_corrected ? (float)(2.0 * 3.14159265358979323846 * fine_foffset) : 0.0f;
    const float inc_re = cosf(-phase_inc), inc_im = sinf(-phase_inc);
    float ph_re = cosf(-plheader_phase), ph_im = sinf(-plheader_phase);
    int in_idx = 0, produced = 0, blk = 0;
    for (int slot = 0; slot < n_slots;) {
        int slots = n_slots - slot;
        if (has_pilots && slots > 16 - (slot % 16))
            slots = 16 - (slot % 16);
        if (has_pilots && coarse_corrected && blk > 0 && (slot % 16) == 0) {
            ph_re = cosf(-pilot_phase[blk - 1]);
            ph_im = sinf(-pilot_phase[blk - 1]);
        }
        const int len = slots * 90;
        for (int k = 0; k < len; k++) { /* the rotator call on this slot sequence */
            static const float lut[4][2] = { { 1, 0 }, { 0, -1 }, { -1, 0 }, { 0, 1 } };
            const float ar = payload[2 * (in_idx + k)], ai = payload[2 * (in_idx + k) + 1];
            const float* d = lut[rn[in_idx + k]];
            const float yr = ar * d[0] - ai * d[1], yi = ar * d[1] + ai * d[0];
            out[2 * (produced + k)] = yr * ph_re - yi * ph_im;
            out[2 * (produced + k) + 1] = yr * ph_im + yi * ph_re;
            const float nr = ph_re * inc_re - ph_im * inc_im, ni = ph_re * inc_im + ph_im * inc_re;
            ph_re = nr, ph_im = ni;
            if ((k + 1) % 512 == 0 || k + 1 == len) {
                const float mag = hypotf(ph_re, ph_im);
                ph_re /= mag, ph_im /= mag;
            }
        }
        in_idx += len;
        produced += len;
        slot += slots;
        if (has_pilots && (slot % 16) == 0 && slot < n_slots) { /* a pilot block follows every 16 slots (not the last) */
            in_idx += 36;
            blk++;
        }
    }
}
