"""Container-only tests (need oracle/_ref = the compiled, unmodified reference): pin the oracle
restatement against the real thing on fresh inputs, including all 57 LDPC tables."""
import numpy as np
import pytest

import golden_inputs as gi

pytestmark = pytest.mark.ref


def test_ldpc_every_table_two_iterations(oracle, ref):
    """Every parity-address table: 16 frames, 2 iterations, generic ISA (16 lanes) -- checks the
    generated circulant data, the link order and the arithmetic for all degrees."""
    import dvbs2rx_b200 as d
    from dvbs2rx_b200 import vectors
    for table in range(d.lib().dvbs2b200_num_tables()):
        n, k = oracle.l.orc_table_n(table), oracle.l.orc_table_k(table)
        cw = vectors.ldpc_encode_bits(table, gi.random_bits(40 + table, (16, k)))
        llr = gi.noisy_llr(cw, 4, 1200, 90 + table)
        name = oracle.table_name(table)
        rp, rr = ref.ldpc_decode(name, llr, 2, isa="generic")
        op, orr = oracle.ldpc_decode(table, llr, 2, lanes=16)
        assert np.array_equal(rr, orr), name
        assert np.array_equal(rp, op), name


def test_ldpc_isas_agree(ref, oracle):
    import dvbs2rx_b200 as d
    from dvbs2rx_b200 import vectors
    info = d.lookup(0, 0, d.C1_2)
    cw = vectors.ldpc_encode_bits(info.table, gi.random_bits(1, (32, info.k_ldpc)))
    llr = gi.noisy_llr(cw, 4, 1300, 2)
    name = oracle.table_name(info.table)
    a, ra = ref.ldpc_decode(name, llr, 25, isa="avx2")
    b, rb = ref.ldpc_decode(name, llr, 25, isa="sse41")
    c, rc = ref.ldpc_decode(name, llr, 25, isa="generic")
    assert np.array_equal(b, c) and np.array_equal(rb, rc)
    if set(ra.tolist()) == {-1} and set(rb.tolist()) == {-1}:
        assert np.array_equal(a, b)  # same iteration count -> byte-identical posteriors


def test_bch_all_dvb_parameter_sets(oracle, ref):
    """encode -> flips -> decode for every (framesize, n, t) the block can be built with
    (lib/qa_bch.cc:652-747 in spirit), oracle vs compiled reference."""
    import dvbs2rx_b200 as d
    seen = set()
    for fs in (0, 1):
        for rate in range(len(d.RATE)):
            try:
                info = d.lookup(0, fs, rate)
            except d.Dvbs2Error:
                continue
            key = (fs, info.nbch, info.t)
            if key in seen or info.kbch % 8 or info.nbch % 8:
                continue
            seen.add(key)
            ho, hr = oracle.bch(fs, info.t, info.nbch), ref.bch(fs, info.t, info.nbch)
            assert oracle.l.orc_bch_k(ho) == ref.l.ref_bch_k(hr) == info.kbch
            msg = gi.random_bytes(7 + info.nbch, (6, info.kbch // 8))
            cw = oracle.bch_encode(ho, msg)
            assert np.array_equal(cw, ref.bch_encode(hr, msg, info.nbch))
            pos = gi.lcg_stream(11 + info.nbch, 6 * 64).reshape(6, 64) % np.uint32(info.nbch)
            for f, ne in enumerate([0, 1, 2, info.t, info.t + 1, 40]):
                for p in list(dict.fromkeys(pos[f].tolist()))[:ne]:
                    cw[f, p >> 3] ^= 0x80 >> (p & 7)
            mo, ro = oracle.bch_decode(ho, cw)
            mr, rr = ref.bch_decode(hr, cw, info.nbch)
            assert np.array_equal(ro, rr), key
            assert np.array_equal(mo, mr), key
    assert len(seen) >= 30


def test_snr_8psk_estimates_against_psk_hh(oracle, ref):
    """orc_snr_8psk vs lib/psk.hh hard()/map() driven as lib/xfecframe_demapper_cb_impl.cc:128-142,267-302."""
    import dvbs2rx_b200 as d
    from dvbs2rx_b200 import vectors
    rng = np.random.default_rng(31)
    for rate_name in ("C3_5", "C2_3", "C25_36"):
        rate = d.RATE[rate_name]
        msg, cw, info = vectors.encode_frames(0, 1, rate, 2, rng)
        iq, n0 = vectors.awgn(vectors.map_symbols(cw, d.MOD_8PSK, rate), 7.0, rng)
        rows = vectors.rows_8psk(rate, 21600)
        llr = (1 - 2 * cw.astype(np.int8)) * 9
        for l in (None, llr):
            a, b = oracle.estimate_snr(4, iq, l, rate), ref.snr_8psk(iq, l, rows)
            assert np.allclose(a, b, rtol=1e-5), (rate_name, a, b)
        # with the true bits as reference the estimate is the channel's Es/N0
        assert np.allclose(10 * np.log10(oracle.estimate_snr(4, iq, llr, rate)), 7.0, atol=0.1)
