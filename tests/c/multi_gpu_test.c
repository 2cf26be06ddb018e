/* multi_gpu_test.c -- dvbs2b200_multi_fec_decode from plain C (no Python, no torch): one batch cut over the
 * listed devices must equal the single-device call on the same frames byte for byte, and a sample of frames
 * must equal the oracle (checker: oracle/liboracle.so).  Also exercises pageable host buffers (malloc).
 *   usage: multi_gpu_test <n_devices> [frames]     (a device index is reused when the box has fewer GPUs)
 * Built and run by tests/test_gpu_multi.py. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/dvbs2_b200.h"
#include "../../oracle/dvbs2_oracle.h"

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd(void)
{
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return (uint32_t)(rng_state >> 32);
}
static float gauss(void)
{
    const float u1 = ((float)(rnd() >> 8) + 1.0f) / 16777217.0f, u2 = (float)(rnd() >> 8) / 16777216.0f;
    return sqrtf(-2.0f * logf(u1)) * cosf(6.2831853f * u2);
}

#define CHECK(call)                                                                         \
    do {                                                                                    \
        int rc__ = (call);                                                                  \
        if (rc__) {                                                                         \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc__, dvbs2b200_last_error());         \
            return 2;                                                                       \
        }                                                                                   \
    } while (0)

int main(int argc, char** argv)
{
    const int want_dev = argc > 1 ? atoi(argv[1]) : 2;
    const int frames = argc > 2 ? atoi(argv[2]) : 200; /* not a multiple of 32 x devices: ragged last range */
    const int standard = 0, framesize = 0 /* short */, rate = 3 /* C1_2 */;
    const float esn0_db = 1.6f; /* some frames converge, some do not */
    const int ndev_box = dvbs2b200_device_count();
    if (ndev_box <= 0) {
        fprintf(stderr, "no CUDA device\n");
        return 3;
    }
    int devices[64];
    for (int i = 0; i < want_dev; ++i)
        devices[i] = i % ndev_box;

    dvbs2b200_code_info info;
    CHECK(dvbs2b200_lookup(standard, framesize, rate, &info));
    int kbch, nbch, t;
    const int table = orc_lookup(standard, framesize, rate, &kbch, &nbch, &t);
    if (table < 0 || kbch != info.kbch || nbch != info.nbch) {
        fprintf(stderr, "oracle / library parameter mismatch\n");
        return 2;
    }
    const int N = info.n_ldpc, K = info.k_ldpc, kb = kbch / 8, nb = nbch / 8;
    orc_ldpc* ol = orc_ldpc_create(table);
    orc_bch* ob = orc_bch_create(0x402B, t, nbch); /* GF(2^14), short frames */

    /* synthetic frames: random BBFRAME -> BCH -> LDPC -> QPSK + AWGN -> int8 LLR (lib/qpsk.h:208-214) */
    int8_t* llr = (int8_t*)malloc((size_t)frames * N); /* pageable on purpose */
    uint8_t* sent = (uint8_t*)malloc((size_t)frames * kb);
    uint8_t* cwb = (uint8_t*)malloc(nb), *bits = (uint8_t*)calloc(K, 1), *cw = (uint8_t*)malloc(N);
    const float n0 = powf(10.0f, -esn0_db / 10.0f), sigma = sqrtf(n0 / 2), scale = 2.0f * sqrtf(2.0f) / n0;
    for (int f = 0; f < frames; ++f) {
        for (int i = 0; i < kb; ++i)
            sent[(size_t)f * kb + i] = (uint8_t)rnd();
        orc_bch_encode(ob, sent + (size_t)f * kb, cwb);
        for (int i = 0; i < nbch; ++i)
            bits[i] = (cwb[i >> 3] >> (7 - (i & 7))) & 1;
        orc_ldpc_encode(ol, bits, cw);
        for (int i = 0; i < N; ++i) {
            const float x = (cw[i] ? -0.70710678f : 0.70710678f) + sigma * gauss();
            float v = rintf(x * scale);
            v = v > 127 ? 127 : v < -128 ? -128 : v;
            llr[(size_t)f * N + i] = (int8_t)v;
        }
    }

    uint8_t* msg1 = (uint8_t*)malloc((size_t)frames * kb), *msgN = (uint8_t*)malloc((size_t)frames * kb);
    int32_t* tr1 = (int32_t*)malloc(frames * 4), *trN = (int32_t*)malloc(frames * 4);
    int32_t* co1 = (int32_t*)malloc(frames * 4), *coN = (int32_t*)malloc(frames * 4);
    dvbs2b200_code* one = NULL;
    CHECK(dvbs2b200_code_create(&one, 0, standard, framesize, rate));
    CHECK(dvbs2b200_fec_decode(one, 0, NULL, NULL, llr, frames, 25, 0, msg1, tr1, co1));
    dvbs2b200_multi* multi = NULL;
    CHECK(dvbs2b200_multi_create(&multi, devices, want_dev, standard, framesize, rate));
    memset(msgN, 0xEE, (size_t)frames * kb);
    CHECK(dvbs2b200_multi_fec_decode(multi, 0, NULL, NULL, llr, frames, 25, 0, msgN, trN, coN));
    int bad = 0, covered = 0;
    for (int i = 0; i < want_dev; ++i) {
        int f0, nf;
        CHECK(dvbs2b200_multi_shard(multi, frames, i, &f0, &nf));
        if (f0 != covered || (f0 % 32) != 0) {
            fprintf(stderr, "device %d: range [%d, %d) does not continue at %d on a multiple of 32\n", i, f0, f0 + nf, covered);
            bad++;
        }
        covered += nf;
    }
    if (covered != frames)
        bad++;
    if (memcmp(msg1, msgN, (size_t)frames * kb) || memcmp(tr1, trN, frames * 4) || memcmp(co1, coN, frames * 4)) {
        fprintf(stderr, "multi-device result differs from the single-device result\n");
        bad++;
    }
    /* group-coupled termination (the reference's SIMD batch of 32) through the same split */
    const int fg = frames / 32 * 32;
    if (fg) {
        CHECK(dvbs2b200_fec_decode(one, 0, NULL, NULL, llr, fg, 25, 32, msg1, tr1, co1));
        CHECK(dvbs2b200_multi_fec_decode(multi, 0, NULL, NULL, llr, fg, 25, 32, msgN, trN, coN));
        if (memcmp(msg1, msgN, (size_t)fg * kb) || memcmp(tr1, trN, fg * 4) || memcmp(co1, coN, fg * 4)) {
            fprintf(stderr, "multi-device result differs from the single-device result (term_group 32)\n");
            bad++;
        }
        CHECK(dvbs2b200_multi_fec_decode(multi, 0, NULL, NULL, llr, frames, 25, 0, msgN, trN, coN)); /* back to per-frame results */
    }
    /* oracle on a sample of frames, one from every device's range */
    int8_t* post = (int8_t*)malloc(N);
    uint8_t* hard = (uint8_t*)malloc(nb), *omsg = (uint8_t*)malloc(kb);
    int converged = 0, checked = 0;
    for (int f = 0; f < frames; f += 17) {
        memcpy(post, llr + (size_t)f * N, N);
        const int ret = orc_ldpc_decode(ol, post, 1, 25);
        orc_pack_hard(post, nbch, hard);
        const int corr = orc_bch_decode(ob, hard, omsg);
        if (ret != trN[f] || corr != coN[f] || memcmp(omsg, msgN + (size_t)f * kb, kb)) {
            fprintf(stderr, "frame %d differs from the oracle (trials %d vs %d, corrections %d vs %d)\n", f, trN[f], ret, coN[f], corr);
            bad++;
        }
        if (ret >= 0 && memcmp(omsg, sent + (size_t)f * kb, kb) == 0)
            converged++;
        checked++;
    }
    printf("multi_gpu_test: %d frames over %d device slots (%d GPUs in the box), %d frames checked against the oracle (%d decoded to what was sent): %s\n",
           frames, want_dev, ndev_box, checked, converged, bad ? "FAIL" : "ok");
    dvbs2b200_multi_destroy(multi);
    dvbs2b200_code_destroy(one);
    return bad ? 1 : 0;
}
