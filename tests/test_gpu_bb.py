"""BB layer on the GPU (through the C ABI) against the oracle: descrambler, stateful deheader over several
calls, descramble-on-the-fly, and the whole chain soft input -> TS packets."""
import numpy as np
import pytest

import bb_cases
from dvbs2rx_b200 import bbframes as bbf

pytestmark = pytest.mark.gpu

RATE_OF_KBCH = {16008: (1, "C1_4"), 3072: (0, "C1_4"), 58192: (1, "C9_10")}


def test_descrambler_matches_oracle(gpu, oracle):
    d = gpu
    rng = np.random.default_rng(5)
    for fs, rate in ((1, "C1_2"), (0, "C2_3"), (1, "C9_10")):
        code = d.Code(0, fs, d.RATE[rate])
        for frames in (1, 7, 300):
            bb = rng.integers(0, 256, size=(frames, code.kbch // 8), dtype=np.uint8)
            assert np.array_equal(code.bb_descramble(bb), oracle.bb_descramble(bb, code.kbch))
        code.close()


@pytest.mark.parametrize("scrambled", [False, True])
def test_deheader_matches_oracle_call_by_call(gpu, oracle, scrambled):
    """Same calls, same TS bytes per call, same counters as the reference's block (via the oracle)."""
    d = gpu
    codes = {}
    for name, kbch, calls in bb_cases.make_cases(ub=True):
        fs, rate = RATE_OF_KBCH[kbch]
        if kbch not in codes:
            codes[kbch] = d.Code(0, fs, d.RATE[rate])
        code = codes[kbch]
        assert code.kbch == kbch
        code.bb_reset()
        o = oracle.bbdeheader(kbch)
        for bb in calls:
            want = o.work(bb)
            got = code.bb_deheader(bbf.scramble(bb) if scrambled else bb, scrambled=scrambled)
            assert got.size == want.size, (name, got.size, want.size)
            assert np.array_equal(got, want), name
        assert code.bb_counters() == o.counters(), name
    for c in codes.values():
        c.close()


def test_deheader_large_batch_and_state_across_batches(gpu, oracle):
    """Thousands of BBFRAMEs per call (the scan walks them in chunks), faults sprinkled in, three calls."""
    d = gpu
    rng = np.random.default_rng(8)
    kbch = 32208  # QPSK 1/2 normal
    code = d.Code(0, 1, d.C1_2)
    n = 3000
    kb = kbch // 8
    up = bbf.ts_packets((n * (kb - 10) + 187) // 188 + 1, rng)
    bb = bbf.bbframe_stream(kbch, n, up)
    for i in rng.choice(n, 40, replace=False):
        bb[i, rng.integers(0, kb)] ^= 1 << rng.integers(0, 8)
    bb = np.delete(bb, [100, 1033, 2500], axis=0)
    o = oracle.bbdeheader(kbch)
    code.bb_reset()
    for part in (bb[:1100], bb[1100:1101], bb[1101:]):
        want = o.work(part)
        got = code.bb_deheader(bbf.scramble(part), scrambled=True)
        assert np.array_equal(got, want)
    c = code.bb_counters()
    assert c == o.counters() and c["errors"] > 0 and c["gaps"] >= 3
    code.close()


def test_chain_soft_input_to_ts_packets(gpu, oracle):
    """TS packets -> BBFRAMEs -> scrambler -> BCH -> LDPC -> QPSK + AWGN -> int8 LLRs, then the product's
    fec_decode_ts: the packets come back bit for bit, and equal the oracle's LDPC -> BCH -> descramble -> deheader."""
    d = gpu
    from dvbs2rx_b200 import vectors
    import oracle_lib  # noqa: F401
    rng = np.random.default_rng(9)
    fs, rate = 0, d.C2_3  # short frames keep the oracle's LDPC quick
    info = d.lookup(0, fs, rate)
    kb = info.kbch // 8
    F = 24
    up = bbf.ts_packets((F * (kb - 10) + 187) // 188 + 1, rng)
    bb = bbf.bbframe_stream(info.kbch, F, up)
    msg_bits = vectors.unpack_bits(bbf.scramble(bb), info.kbch)
    cw = vectors.ldpc_encode_bits(info.table, vectors.bch_encode_bits(msg_bits, fs, info.t, info.nbch)[:, :info.k_ldpc])
    iq, n0 = vectors.awgn(vectors.map_symbols(cw, d.MOD_QPSK, rate), 4.2, rng)
    llr = vectors.qpsk_llr(iq, n0)
    code = d.Code(0, fs, rate)
    ts, trials, corr = code.fec_decode_ts(llr=llr, max_trials=25)
    assert (trials >= 0).all()
    n_full = F * (kb - 10) // 188
    assert np.array_equal(ts, up[:n_full].ravel())
    # the same through the oracle, stage by stage
    o_post, o_ret = oracle.ldpc_decode(info.table, llr, 25)
    o_msg, o_corr = oracle.bch_decode(oracle.bch(fs, info.t, info.nbch), oracle.pack_hard(o_post, info.nbch))
    o_ts = oracle.bbdeheader(info.kbch).work(oracle.bb_descramble(o_msg, info.kbch))
    assert np.array_equal(ts, o_ts) and np.array_equal(corr, o_corr) and np.array_equal(trials, o_ret)
    assert code.bb_counters()["packets"] == n_full
    # from symbols, second call continues the stream state: nothing new is produced for an empty batch
    code.bb_reset()
    ts2, _, _ = code.fec_decode_ts(iq=iq, n0=n0, constellation=d.MOD_QPSK, max_trials=25)
    assert np.array_equal(ts2, ts)
    # noise so strong that LDPC and BCH fail: headers fail their CRC, frames are dropped, no packet comes out
    iq_bad, n0_bad = vectors.awgn(vectors.map_symbols(cw[:4], d.MOD_QPSK, rate), -3.0, rng)
    code.bb_reset()
    ts3, tr3, co3 = code.fec_decode_ts(llr=vectors.qpsk_llr(iq_bad, n0_bad), max_trials=5)
    o_post, _ = oracle.ldpc_decode(info.table, vectors.qpsk_llr(iq_bad, n0_bad), 5)
    o_msg, _ = oracle.bch_decode(oracle.bch(fs, info.t, info.nbch), oracle.pack_hard(o_post, info.nbch))
    assert np.array_equal(ts3, oracle.bbdeheader(info.kbch).work(oracle.bb_descramble(o_msg, info.kbch)))
    code.close()


def test_deheader_random_headers_match_oracle(gpu, oracle):
    """Random BBHEADERs (valid CRC, arbitrary DFL / SYNCD, some broken) over random payloads: every state
    transition class of the deheader's frame-to-frame recurrence, in random order, over several calls.
    The product's parallel scan must agree with the oracle's frame-by-frame walk byte for byte."""
    d = gpu
    kbch, kb = 3072, 384  # QPSK 1/4 short: the smallest BBFRAME, many frames per packet boundary pattern
    code = d.Code(0, 0, d.C1_4)
    assert code.kbch == kbch
    max_df = kb - 10
    for seed in range(6):
        rng = np.random.default_rng(100 + seed)
        n = 700
        bb = rng.integers(0, 256, size=(n, kb), dtype=np.uint8)
        pending = 0  # bytes of a cut packet a well-formed stream would announce
        for f in range(n):
            r = rng.random()
            df = int(rng.choice([max_df, max_df, max_df, 188 * (max_df // 188), 100, 0, 187, 188, 8 * int(rng.integers(0, max_df // 8 + 1)) // 8]))
            df = min(df, max_df)
            if r < 0.55:    # continues the stream
                syncd = (188 - pending) % 188
            elif r < 0.75:  # random but valid SYNCD
                syncd = int(rng.integers(0, df + 1)) if df else 0
            elif r < 0.85:  # SYNCD == DFL (the re-synchronisation edge)
                syncd = df
            else:
                syncd = int(rng.integers(0, 400))
            hdr = bbf.bbheader(df * 8, min(syncd, 8191) * 8).copy()
            if r > 0.93:
                hdr[int(rng.integers(0, 10))] ^= 1 << int(rng.integers(0, 8))  # CRC failure
            elif r > 0.90:
                hdr = bbf.bbheader(df * 8 + 4, 0).copy()                      # DFL not a multiple of 8
            bb[f, :10] = hdr
            pending = (pending + df) % 188
        o = oracle.bbdeheader(kbch)
        code.bb_reset()
        cuts = sorted(rng.choice(np.arange(1, n), size=3, replace=False).tolist())
        for lo, hi in zip([0] + cuts, cuts + [n]):
            want = o.work(bb[lo:hi])
            got = code.bb_deheader(bb[lo:hi], scrambled=False)
            assert got.size == want.size and np.array_equal(got, want), (seed, lo, hi)
        assert code.bb_counters() == o.counters(), seed
    code.close()


def test_deheader_refuses_a_ts_buffer_below_capacity(gpu):
    """A clipped batch would lose packets while the deheader state moves on: EINVAL, nothing consumed."""
    import ctypes as C
    d = gpu
    code = d.Code(0, 1, d.RATE["C1_2"])
    rng = np.random.default_rng(3)
    up = bbf.ts_packets(400, rng)
    bb = bbf.bbframe_stream(code.kbch, 8, up)
    cap = code.bb_ts_capacity(8)
    ts = np.zeros(cap, dtype=np.uint8)
    n = C.c_size_t(7)
    rc = d.lib().dvbs2b200_bb_deheader(code._h, bb.ctypes.data, 8, 0, ts.ctypes.data, cap - 188, C.byref(n))
    assert rc == -1 and n.value == 0  # DVBS2B200_EINVAL
    assert b"capacity" in d.lib().dvbs2b200_last_error()
    assert code.bb_counters()["bbframes"] == 0
    got = code.bb_deheader(bb)  # the same frames go through afterwards
    assert got.size > 0 and got.size % 188 == 0
    code.close()
