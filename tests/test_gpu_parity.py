"""GPU parity tests: every call goes through the C ABI (libdvbs2_b200.so) and is compared
bit-for-bit with the oracle on the same seeded inputs.  Run on the B200 box: pytest -m gpu."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

S2, NORMAL, SHORT = 0, 1, 0


def _bch_of(orc, d, framesize, info):
    return orc.bch(framesize, info.t, info.nbch)


@pytest.mark.parametrize("rate_name,framesize,esn0,frames", [
    ("C1_2", NORMAL, 1.0, 12),   # BASELINE config 1: nothing converges, 25 full iterations
    ("C1_2", NORMAL, 2.0, 12),   # converging variant
    ("C3_4", NORMAL, 4.6, 8),    # config 2 (max 50 trials below)
    ("C3_5", NORMAL, 2.6, 8),    # config 3's code (QPSK-equivalent LLRs)
    ("C2_3", SHORT, 3.4, 16),    # config 4's code
    ("C9_10", NORMAL, 6.6, 8),   # config 5's code (64-bit check-node state)
])
def test_ldpc_per_frame_matches_oracle(gpu, oracle, rate_name, framesize, esn0, frames):
    d = gpu
    from dvbs2rx_b200 import vectors
    rate = d.RATE[rate_name]
    msg, cw, llr, info = vectors.make_llr_frames(S2, framesize, rate, frames, esn0, seed=100 + rate)
    trials = 50 if rate_name == "C3_4" else 25
    code = d.Code(S2, framesize, rate)
    for om in (d.OM_MESSAGE, d.OM_CODEWORD):
        hard, post, left = code.ldpc_decode(llr, trials, d.TERM_PER_FRAME, om, want_post=True)
        o_post, o_left = oracle.ldpc_decode(info.table, llr, trials)
        assert np.array_equal(left, o_left)
        assert np.array_equal(post, o_post)
        nbits = info.nbch if om == d.OM_MESSAGE else info.n_ldpc
        assert np.array_equal(hard, oracle.pack_hard(o_post, nbits))
    code.close()


@pytest.mark.parametrize("group", [16, 32])
def test_ldpc_group_termination_matches_reference_semantics(gpu, oracle, group):
    """term_group = SIMD width reproduces lib/ldpc_decoder/layered_decoder.hh:153 exactly:
    posterior bytes and the per-batch return value depend on the whole batch."""
    d = gpu
    from dvbs2rx_b200 import vectors
    msg, cw, llr, info = vectors.make_llr_frames(S2, SHORT, d.C1_2, 2 * group, 1.6, seed=11)
    code = d.Code(S2, SHORT, d.C1_2)
    hard, post, left = code.ldpc_decode(llr, 25, group, d.OM_CODEWORD, want_post=True)
    o_post, o_left = oracle.ldpc_decode(info.table, llr, 25, lanes=group)
    assert np.array_equal(left, o_left)
    assert np.array_equal(post, o_post)
    assert np.array_equal(hard, oracle.pack_hard(o_post, info.n_ldpc))
    code.close()


def test_ldpc_group_termination_more_frames_than_resident_ctas(gpu, oracle):
    """Group mode with the persistent loop actually looping: 2 x 444 resident CTAs' worth of frames (rounded to whole
    groups), so every CTA decodes several frames and the per-group arrival counters of different passes interleave.
    Sampled groups are compared with the oracle's coupled loop (lib/ldpc_decoder/layered_decoder.hh:153)."""
    d = gpu
    from dvbs2rx_b200 import vectors
    group = 32
    frames = (2 * 444 + group) // group * group  # 896
    base = 96
    msg, cw, llr, info = vectors.make_llr_frames(S2, SHORT, d.C1_2, base, 1.55, seed=13)
    reps = (frames + base - 1) // base
    # the batch is the 96 distinct frames repeated with a rotation per repetition, so the groups differ in composition
    llr = np.concatenate([np.roll(llr, 7 * r, axis=0) for r in range(reps)])[:frames]
    code = d.Code(S2, SHORT, d.C1_2)
    hard, post, left = code.ldpc_decode(llr, 25, group, d.OM_CODEWORD, want_post=True)
    assert len(set(left.tolist())) > 1  # the groups really stop at different iterations
    for g in (0, 5, 13, frames // group - 1):
        sl = slice(g * group, (g + 1) * group)
        o_post, o_left = oracle.ldpc_decode(info.table, llr[sl], 25, lanes=group)
        assert np.array_equal(left[sl], o_left), g
        assert np.array_equal(post[sl], o_post), g
        assert np.array_equal(hard[sl], oracle.pack_hard(o_post, info.n_ldpc)), g
    code.close()


def test_ldpc_clean_codeword_returns_max_trials(gpu):
    d = gpu
    from dvbs2rx_b200 import vectors
    rng = np.random.default_rng(3)
    msg, cw, info = vectors.encode_frames(S2, NORMAL, d.C1_2, 4, rng)
    llr = ((1 - 2 * cw.astype(np.int8)) * 9).astype(np.int8)
    code = d.Code(S2, NORMAL, d.C1_2)
    hard, post, left = code.ldpc_decode(llr, 0, 0, d.OM_CODEWORD, want_post=True)  # 0 -> 25 trials
    assert left.tolist() == [25] * 4
    assert np.array_equal(post, llr)
    assert np.array_equal(vectors.unpack_bits(hard), cw)
    code.close()


@pytest.mark.parametrize("rate_name,framesize", [("C1_2", NORMAL), ("C2_3", SHORT), ("C9_10", NORMAL), ("C2_3", NORMAL)])
def test_bch_injected_errors_match_oracle(gpu, oracle, rate_name, framesize):
    d = gpu
    rate = d.RATE[rate_name]
    info = d.lookup(S2, framesize, rate)
    h = _bch_of(oracle, d, framesize, info)
    rng = np.random.default_rng(17 + rate)
    nerr = [0, 1, 2, 3, info.t - 1, info.t, info.t + 1, 13, 30, 200] * 3
    F = len(nerr)
    msg = rng.integers(0, 256, size=(F, info.kbch // 8), dtype=np.uint8)
    cw = oracle.bch_encode(h, msg)
    for f in range(F):
        for p in rng.choice(info.nbch, size=nerr[f], replace=False):
            cw[f, p >> 3] ^= 0x80 >> (p & 7)
    garbage = rng.integers(0, 256, size=(8, info.nbch // 8), dtype=np.uint8)
    cw = np.concatenate([cw, garbage])
    code = d.Code(S2, framesize, rate)
    out, corr = code.bch_decode(cw)
    o_out, o_corr = oracle.bch_decode(h, cw)
    assert np.array_equal(corr, o_corr)
    assert np.array_equal(out, o_out)
    ok = np.array(nerr) <= info.t
    assert np.array_equal(out[:F][ok], msg[ok])
    code.close()


def test_demap_qpsk_matches_oracle(gpu, oracle):
    d = gpu
    rng = np.random.default_rng(23)
    iq = (rng.standard_normal(size=(5, 32400, 2)) * 0.8).astype(np.float32)
    iq[0, :8] = [[1, 1], [1, -1], [-1, -1], [-1, 1], [0.5, 0.25], [1e3, -1e3], [0.1249999, -0.125], [0, 0]]
    n0 = np.array([1.0, 0.794, 0.05, 2.5, 0.3], dtype=np.float32)
    code = d.Code(S2, NORMAL, d.C1_2)
    llr = code.demap(d.MOD_QPSK, iq, n0)
    assert np.array_equal(llr, oracle.demap_qpsk(iq, n0))
    code.close()


@pytest.mark.parametrize("rate_name", ["C3_5", "C2_3", "C25_36"])
def test_demap_8psk_matches_oracle(gpu, oracle, rate_name):
    d = gpu
    rate = d.RATE[rate_name]
    rng = np.random.default_rng(29)
    iq = (rng.standard_normal(size=(4, 21600, 2)) * 0.7).astype(np.float32)
    n0 = np.array([0.24, 0.05, 0.9, 0.31], dtype=np.float32)
    code = d.Code(S2, NORMAL, rate)
    llr = code.demap(d.MOD_8PSK, iq, n0)
    assert np.array_equal(llr, oracle.demap_8psk(iq, n0, rate))
    code.close()


def test_demap_unsupported_constellation(gpu):
    d = gpu
    code = d.Code(S2, NORMAL, d.C2_3)
    with pytest.raises(d.Dvbs2Error) as e:
        code.demap(d.MOD_16APSK, np.zeros((1, 16200, 2), np.float32), 1.0)
    assert e.value.code == d.EUNSUPPORTED
    code.close()


def test_fused_chain_from_symbols_8psk(gpu, oracle):
    """Config 3: 8PSK 3/5 normal, symbols -> demap -> LDPC -> BCH, TS-side bytes bit-exact."""
    d = gpu
    from dvbs2rx_b200 import vectors
    rng = np.random.default_rng(31)
    F = 8
    msg, cw, info = vectors.encode_frames(S2, NORMAL, d.C3_5, F, rng)
    iq, n0 = vectors.awgn(vectors.map_symbols(cw, d.MOD_8PSK, d.C3_5), 6.2, rng)
    code = d.Code(S2, NORMAL, d.C3_5)
    out, left, corr = code.fec_decode(iq=iq, n0=n0, constellation=d.MOD_8PSK, max_trials=25)
    llr = oracle.demap_8psk(iq, n0, d.C3_5)
    o_post, o_left = oracle.ldpc_decode(info.table, llr, 25)
    o_out, o_corr = oracle.bch_decode(_bch_of(oracle, d, NORMAL, info), oracle.pack_hard(o_post, info.nbch))
    assert np.array_equal(left, o_left)
    assert np.array_equal(corr, o_corr)
    assert np.array_equal(out, o_out)
    assert np.array_equal(out, msg)  # 6.2 dB decodes cleanly
    code.close()


def test_fused_chain_non_converging_matches_oracle(gpu, oracle):
    """Config 1 at 1.0 dB: LDPC never converges, so every frame takes the BCH failure path
    (lib/bch.cc:476-483); the emitted bytes must still match."""
    d = gpu
    from dvbs2rx_b200 import vectors
    msg, cw, llr, info = vectors.make_llr_frames(S2, NORMAL, d.C1_2, 6, 1.0, seed=41)
    code = d.Code(S2, NORMAL, d.C1_2)
    out, left, corr = code.fec_decode(llr=llr, max_trials=25)
    o_post, o_left = oracle.ldpc_decode(info.table, llr, 25)
    o_out, o_corr = oracle.bch_decode(_bch_of(oracle, d, NORMAL, info), oracle.pack_hard(o_post, info.nbch))
    assert left.tolist() == [-1] * 6 and np.array_equal(left, o_left)
    assert np.array_equal(corr, o_corr)
    assert np.array_equal(out, o_out)
    code.close()


def test_tables_roundtrip_create_from_blob(gpu):
    d = gpu
    from dvbs2rx_b200 import vectors
    a = d.Code(S2, SHORT, d.C2_3)
    b = d.Code(tables=a.export_tables())
    msg, cw, llr, info = vectors.make_llr_frames(S2, SHORT, d.C2_3, 4, 4.0, seed=5)
    ha, _, ta = a.ldpc_decode(llr)
    hb, _, tb = b.ldpc_decode(llr)
    assert np.array_equal(ha, hb) and np.array_equal(ta, tb)
    a.close()
    b.close()


def test_snr_estimate_matches_oracle(gpu, oracle):
    """dvbs2b200_estimate_snr (GPU) vs the oracle, QPSK and 8PSK, from sliced symbols and from posterior
    LLR signs.  Tolerance 1e-4 relative: float sums, the reference's own order (VOLK) is unspecified."""
    d = gpu
    from dvbs2rx_b200 import vectors
    rng = np.random.default_rng(41)
    for mod, rate_name, esn0 in ((d.MOD_QPSK, "C1_2", 1.0), (d.MOD_QPSK, "C3_4", 9.0), (d.MOD_8PSK, "C3_5", 6.2),
                                 (d.MOD_8PSK, "C2_3", 12.0), (d.MOD_8PSK, "C25_36", 8.0)):
        fs = 2 if rate_name == "C25_36" else 1
        std = 0
        rate = d.RATE[rate_name]
        if rate_name == "C25_36":
            fs = 1
        msg, cw, info = vectors.encode_frames(std, fs, rate, 5, rng)
        iq, n0 = vectors.awgn(vectors.map_symbols(cw, mod, rate), esn0, rng)
        code = d.Code(std, fs, rate)
        llr = ((1 - 2 * cw.astype(np.int16)) * rng.integers(1, 100, size=cw.shape)).astype(np.int8)
        llr[:, ::977] = 0  # a zero LLR counts as positive
        for l in (None, llr):
            got = code.estimate_snr(mod, iq, l)
            want = oracle.estimate_snr(mod, iq, l, rate)
            assert np.allclose(got, want, rtol=1e-4), (rate_name, got, want)
        # with the transmitted bits as reference the estimate is the channel Es/N0
        assert np.allclose(10 * np.log10(code.estimate_snr(mod, iq, llr)), esn0, atol=0.3)
        code.close()
    with pytest.raises(d.Dvbs2Error):
        d.Code(0, 1, d.C1_2).estimate_snr(d.MOD_16APSK, np.zeros((1, 16200, 2), np.float32))


def test_mixed_modcod_batch_matches_per_code_oracle(gpu, oracle):
    """BASELINE config 5: frames of five MODCODs interleaved in one batch, per-frame code id.  The result of
    every frame equals the single-code oracle run on that frame (oracle = per-code runs re-assembled)."""
    d = gpu
    from dvbs2rx_b200 import vectors
    modcods = [(0, 1, d.C1_2, 2.0), (0, 1, d.C3_4, 4.6), (0, 1, d.C3_5, 3.5), (0, 0, d.C2_3, 4.2), (0, 1, d.C9_10, 6.6)]
    mixed = d.MixedCodes([m[:3] for m in modcods])
    rng = np.random.default_rng(51)
    per_code = {}
    for c, (std, fs, rate, esn0) in enumerate(modcods):
        msg, cw, llr, info = vectors.make_llr_frames(std, fs, rate, 3 if fs else 6, esn0, seed=60 + c)
        per_code[c] = dict(msg=msg, llr=llr, info=info, fs=fs, next=0)
    order = np.concatenate([np.full(per_code[c]["llr"].shape[0], c, dtype=np.uint8) for c in per_code])
    rng.shuffle(order)
    llr_cat, want_msg, want_tr, want_co = [], [], [], []
    for c in order:
        pc = per_code[int(c)]
        i = pc["next"]
        pc["next"] += 1
        llr_cat.append(pc["llr"][i])
        info = pc["info"]
        o_post, o_ret = oracle.ldpc_decode(info.table, pc["llr"][i:i + 1], 25)
        o_msg, o_corr = oracle.bch_decode(oracle.bch(pc["fs"], info.t, info.nbch), oracle.pack_hard(o_post, info.nbch))
        want_msg.append(o_msg[0])
        want_tr.append(o_ret[0])
        want_co.append(o_corr[0])
    msg, tr, co = mixed.fec_decode(order, np.concatenate(llr_cat), 25)
    assert np.array_equal(msg, np.concatenate(want_msg))
    assert np.array_equal(tr, np.array(want_tr)) and np.array_equal(co, np.array(want_co))
    assert (tr >= 0).sum() >= len(order) - 2  # nearly everything converges at these SNRs
    mixed.close()


def test_table_demapper_16apsk_32apsk(gpu, oracle):
    """dvbs2b200_demap_table (no reference counterpart): within 1 LSB of a float64 max-log model, QPSK through
    it equals the reference-exact QPSK demapper, and 16APSK / 32APSK frames decode through LDPC + BCH."""
    d = gpu
    from dvbs2rx_b200 import apsk, vectors
    rng = np.random.default_rng(61)
    # QPSK as a table: same bytes as dvbs2b200_demap (lib/qpsk.h:208-214) away from rounding ties
    code = d.Code(0, 0, d.C2_3)
    a = np.float32(np.sqrt(0.5))
    qpsk = np.array([[a, a], [a, -a], [-a, a], [-a, -a]], dtype=np.float32)
    msg, cw, info = vectors.encode_frames(0, 0, d.C2_3, 3, rng)
    iq, n0 = vectors.awgn(vectors.map_symbols(cw, d.MOD_QPSK, d.C2_3), 4.0, rng)
    offs = np.array([0, 1], dtype=np.int32)  # QPSK is not interleaved: bit k of symbol j at 2j + k
    ref = code.demap(d.MOD_QPSK, iq, n0).astype(np.int16)
    got = code.demap_table(qpsk, apsk.row_offsets(16200, 2), iq, n0).astype(np.int16)
    got = np.stack([got[:, :8100], got[:, 8100:]], axis=2).reshape(3, 16200)  # rows -> interleaved pairs
    assert np.abs(got - ref).max() <= 1 and (got != ref).mean() < 0.01
    del offs
    for name, pts, fs, rate, esn0 in (("16APSK 2/3 short", apsk.points_16apsk(apsk.GAMMA_16APSK["C2_3"]), 0, d.C2_3, 10.5),
                                      ("32APSK 9/10 normal", apsk.points_32apsk(*apsk.GAMMA_32APSK["C9_10"]), 1, d.C9_10, 17.5)):
        assert abs(float((pts.astype(np.float64) ** 2).sum(axis=1).mean()) - 1.0) < 1e-6
        assert len({(round(float(x), 5), round(float(y), 5)) for x, y in pts}) == pts.shape[0]  # distinct points
        code = d.Code(0, fs, rate)
        bits = int(pts.shape[0]).bit_length() - 1
        F = 4
        msg, cw, info = vectors.encode_frames(0, fs, rate, F, rng)
        offs = apsk.row_offsets(info.n_ldpc, bits)
        iq, n0 = vectors.awgn(apsk.map_bits(cw, pts, offs), esn0, rng)
        llr = code.demap_table(pts, offs, iq, n0)
        model = apsk.maxlog_llr(iq, pts, offs, n0)
        want = np.clip(np.rint(model), -128, 127)
        assert np.abs(llr.astype(np.float64) - want).max() <= 1, name
        assert (llr.astype(np.float64) != want).mean() < 0.01, name
        out, trials, corr = code.fec_decode(llr=llr, max_trials=25)
        assert (trials >= 0).all() and np.array_equal(out, msg), name
        code.close()
    with pytest.raises(d.Dvbs2Error):
        d.Code(0, 1, d.C1_2).demap_table(np.zeros((64, 2), np.float32), np.zeros(6, np.int32), np.zeros((1, 10800, 2), np.float32), 1.0)


def test_host_register_takes_the_direct_copy_path(gpu, oracle):
    """A caller-owned pageable buffer page-locked in place (dvbs2b200_host_register) gives the same bytes as the
    staged pageable path and as pinned memory; unregistering twice is an error, not a crash."""
    d = gpu
    from dvbs2rx_b200 import vectors
    rate = d.RATE["C1_2"]
    msg, cw, llr, info = vectors.make_llr_frames(S2, NORMAL, rate, 40, 2.0, seed=9)
    code = d.Code(S2, NORMAL, rate)
    want, tr0, co0 = code.fec_decode(llr=llr)            # pageable: staged through the pinned ring
    buf = np.array(llr, copy=True)
    d.host_register(buf)
    try:
        got, tr1, co1 = code.fec_decode(llr=buf)         # registered: copied from directly
    finally:
        d.host_unregister(buf)
    assert np.array_equal(got, want) and np.array_equal(tr0, tr1) and np.array_equal(co0, co1)
    assert np.array_equal(got, msg)
    with pytest.raises(d.Dvbs2Error):
        d.host_unregister(buf)
    code.close()
