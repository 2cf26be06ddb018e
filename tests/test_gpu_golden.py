"""GPU tests against the committed reference fixtures (tests/golden/, produced by the compiled
reference) and size-independent properties at BASELINE.json's full batch sizes."""
import numpy as np
import pytest

import golden_inputs as gi
from test_oracle_golden import LDPC_CASES, ldpc_case_inputs, load, sha

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", LDPC_CASES, ids=lambda c: "%s-s%d" % (c["name"], c["sigma_q8"]))
def test_ldpc_matches_reference_fixture(gpu, case):
    """term_group 32 / 16 = the reference's AVX2 / SSE4.1-generic batch semantics: posterior bytes,
    packed hard decisions and the per-batch return value equal the compiled reference's."""
    d = gpu
    info, cw, llr = ldpc_case_inputs(case)
    code = d.Code(d.STANDARD_DVBS2, case["framesize"], d.RATE[case["rate"]])
    for group, key in ((32, "group32"), (16, "group16")):
        hard, post, left = code.ldpc_decode(llr, case["trials"], group, d.OM_CODEWORD, want_post=True)
        exp = case[key]
        assert left.tolist() == exp["ret"]
        assert sha(post) == exp["post_sha256"]
        assert sha(hard) == exp["hard_sha256"]
    # per-frame termination: hard decisions of converged frames equal the transmitted bits
    hard, _, left = code.ldpc_decode(llr, case["trials"], d.TERM_PER_FRAME, d.OM_CODEWORD)
    from dvbs2rx_b200 import vectors
    ok = left >= 0
    assert np.array_equal(vectors.unpack_bits(hard)[ok], cw[ok])
    code.close()


@pytest.mark.parametrize("case", load("bch.json")["cases"], ids=lambda c: c["name"])
def test_bch_matches_reference_fixture(gpu, oracle, case):
    d = gpu
    fs, n, k, t = case["framesize"], case["n"], case["k"], case["t"]
    F = len(case["nerr"])
    msg = gi.random_bytes(case["seed"], (F, k // 8))
    cw = oracle.bch_encode(oracle.bch(fs, t, n), msg)  # input construction only
    pos = gi.lcg_stream(case["seed"] + 1, F * 256).reshape(F, 256) % np.uint32(n)
    for f in range(F):
        for p in list(dict.fromkeys(pos[f].tolist()))[:case["nerr"][f]]:
            cw[f, p >> 3] ^= 0x80 >> (p & 7)
    allcw = np.concatenate([cw, gi.random_bytes(case["seed"] + 2, (case["garbage_frames"], n // 8))])
    assert sha(allcw) == case["cw_sha256"]
    code = d.Code(d.STANDARD_DVBS2, fs, d.RATE[case["rate"]])
    out, ret = code.bch_decode(allcw)
    assert ret.tolist() == case["ret"]
    assert sha(out) == case["out_sha256"]
    code.close()


@pytest.mark.parametrize("case", load("demap.json")["cases"], ids=lambda c: "%s-n0_%s" % (c["rate"], c["n0"]))
def test_8psk_demap_matches_reference_fixture(gpu, case):
    d = gpu
    iq = gi.complex_symbols(case["seed"], (case["frames"], case["n_syms"]))
    code = d.Code(d.STANDARD_DVBS2, d.FECFRAME_NORMAL, d.RATE[case["rate"]])
    out = code.demap(d.MOD_8PSK, iq, case["n0"])
    assert out[0, :24].tolist() == case["llr_head"]
    assert sha(out) == case["llr_sha256"]
    code.close()


def test_full_batch_roundtrip_config1_sizes(gpu):
    """BASELINE-size batch (2368 normal frames = one bench step): encode -> AWGN at 2.0 dB -> decode
    gives back every BBFRAME; the same batch with the all-zero... linearity: decoding is invariant
    under adding a codeword (sign-flipping the LLRs of its one bits)."""
    d = gpu
    from dvbs2rx_b200 import vectors
    F = 2368
    rng = np.random.default_rng(77)
    msg, cw, info = vectors.encode_frames(0, 1, d.C1_2, 64, rng)
    reps = F // 64
    cw_all = np.tile(cw, (reps, 1))
    iq, n0 = vectors.awgn(vectors.map_symbols(cw_all, d.MOD_QPSK, d.C1_2), 2.0, rng)
    llr = vectors.qpsk_llr(iq, n0)
    code = d.Code(0, 1, d.C1_2)
    out, left, corr = code.fec_decode(llr=llr, max_trials=25)
    assert (left >= 0).all() and (corr >= 0).all()
    assert np.array_equal(out, np.tile(msg, (reps, 1)))
    # linearity / symmetry of min-sum: flip the LLR signs where another codeword has ones
    other = np.roll(cw_all, 1, axis=0)
    llr2 = np.where(other == 1, -llr.astype(np.int16), llr.astype(np.int16))
    llr2 = np.clip(llr2, -128, 127).astype(np.int8)  # -(-128) saturates exactly like the channel would
    hard1, _, left1 = code.ldpc_decode(llr, 25, 0, d.OM_CODEWORD)
    hard2, _, left2 = code.ldpc_decode(llr2, 25, 0, d.OM_CODEWORD)
    sat = (llr == -128).any(axis=1)
    assert np.array_equal(left1[~sat], left2[~sat])
    assert np.array_equal(vectors.unpack_bits(hard2)[~sat], (vectors.unpack_bits(hard1) ^ other)[~sat])
    code.close()


def test_ragged_and_empty_batches(gpu, oracle):
    d = gpu
    from dvbs2rx_b200 import vectors
    code = d.Code(0, 0, d.C2_3)
    hard, post, left = code.ldpc_decode(np.zeros((0, 16200), np.int8), 25, 0, d.OM_MESSAGE, want_post=True)
    assert hard.shape == (0, 1350) and left.shape == (0,)
    msg, cw, llr, info = vectors.make_llr_frames(0, 0, d.C2_3, 445, 3.2, seed=3)  # not a multiple of the grid
    hard, _, left = code.ldpc_decode(llr, 25, 0, d.OM_MESSAGE)
    pick = [0, 1, 443, 444, 222]
    o_post, o_left = oracle.ldpc_decode(info.table, llr[pick], 25)
    assert np.array_equal(left[pick], o_left)
    assert np.array_equal(hard[pick], oracle.pack_hard(o_post, info.nbch))
    with pytest.raises(d.Dvbs2Error):
        code.ldpc_decode(llr[:33], 25, 32, d.OM_MESSAGE)  # group mode needs whole groups
    # all-zero LLRs: every check counts as unsatisfied (vsign(., 0) = 0), nothing ever changes
    z = np.zeros((2, 16200), np.int8)
    hard, post, left = code.ldpc_decode(z, 5, 0, d.OM_CODEWORD, want_post=True)
    o_post, o_left = oracle.ldpc_decode(info.table, z, 5)
    assert np.array_equal(post, o_post) and np.array_equal(left, o_left)
    # saturated inputs
    s = np.where(cw[:2] == 1, -128, 127).astype(np.int8)
    hard, post, left = code.ldpc_decode(s, 25, 0, d.OM_CODEWORD, want_post=True)
    o_post, o_left = oracle.ldpc_decode(info.table, s, 25)
    assert np.array_equal(post, o_post) and np.array_equal(left, o_left)
    code.close()


@pytest.mark.parametrize("rate_name,fs", [("C1_4", 1), ("C1_3", 1), ("C2_5", 1), ("C4_5", 1), ("C5_6", 1), ("C8_9", 1),
                                          ("C1_4", 0), ("C1_2", 0), ("C3_4", 0), ("C5_6", 0), ("C8_9", 0),
                                          ("C1_3_MEDIUM", 2), ("C13_45", 1), ("C154_180", 1), ("C32_45", 0)])
def test_other_modcods_match_oracle(gpu, oracle, rate_name, fs):
    """Every degree class / kernel instantiation: 4 frames, 6 iterations, vs the oracle."""
    d = gpu
    from dvbs2rx_b200 import vectors
    rate = d.RATE[rate_name]
    info = d.lookup(0, fs, rate)
    cw = vectors.ldpc_encode_bits(info.table, gi.random_bits(rate, (4, info.k_ldpc)))
    llr = gi.noisy_llr(cw, 4, 1150, 50 + rate)
    code = d.Code(0, fs, rate)
    hard, post, left = code.ldpc_decode(llr, 6, 0, d.OM_CODEWORD, want_post=True)
    o_post, o_left = oracle.ldpc_decode(info.table, llr, 6)
    assert np.array_equal(left, o_left)
    assert np.array_equal(post, o_post)
    assert np.array_equal(hard, oracle.pack_hard(o_post, info.n_ldpc))
    code.close()


def test_every_table_four_iterations(gpu, oracle):
    """All 57 LDPC tables (DVB-S2, S2X, T2; short, medium, normal) through the default schedule: posteriors,
    packed hard decisions and return values equal the oracle after a fixed number of iterations at low SNR
    (and with early termination where a low-rate code converges).  Guards the split / chain / level forms of
    every table's conflict layers, including the ones no BASELINE config touches."""
    d = gpu
    from dvbs2rx_b200 import vectors
    seen = {}
    for std in (0, 1):
        for fs in (0, 1, 2):
            for name, rate in d.RATE.items():
                try:
                    info = d.lookup(std, fs, rate)
                except d.Dvbs2Error:
                    continue
                seen.setdefault(info.table, (std, fs, rate))
    assert len(seen) == 57
    rng = np.random.default_rng(77)
    for table, (std, fs, rate) in sorted(seen.items()):
        info = d.lookup(std, fs, rate)
        bits = rng.integers(0, 2, size=(2, info.k_ldpc), dtype=np.uint8)
        cw = vectors.ldpc_encode_bits(info.table, bits)
        iq, n0 = vectors.awgn(vectors.map_symbols(cw, d.MOD_QPSK, rate), 0.5, rng)
        llr = vectors.qpsk_llr(iq, n0)
        code = d.Code(std, fs, rate)
        hard, post, trials = code.ldpc_decode(llr, 4, d.TERM_PER_FRAME, d.OM_CODEWORD, want_post=True)
        o_post, o_ret = oracle.ldpc_decode(info.table, llr, 4)
        assert np.array_equal(trials, o_ret), (table, trials, o_ret)
        assert np.array_equal(post, o_post), table
        assert np.array_equal(hard, oracle.pack_hard(o_post, info.n_ldpc)), table
        code.close()
