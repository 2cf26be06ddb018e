"""BB layer on the CPU: the oracle restatement against (a) the reference's own bbdescrambler / bbdeheader
translation units compiled unmodified (oracle/_ref, container only), (b) the reference's QA cases
(python/dvbs2rx/qa_bbdeheader_bb.py) restated, (c) the committed fixture tests/golden/bb.json."""
import hashlib
import json
import os

import numpy as np
import pytest

import bb_cases
from dvbs2rx_b200 import bbframes as bbf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_prbs_and_descrambler_against_reference(oracle, ref):
    import dvbs2rx_b200 as d
    rng = np.random.default_rng(3)
    for fs, rate in ((1, "C1_4"), (1, "C9_10"), (0, "C1_4"), (0, "C8_9"), (1, "C3_5")):
        info = d.lookup(0, fs, d.RATE[rate])
        bb = rng.integers(0, 256, size=(3, info.kbch // 8), dtype=np.uint8)
        want = ref.bb_descramble(0, fs, d.RATE[rate], bb).reshape(bb.shape)
        assert np.array_equal(oracle.bb_descramble(bb, info.kbch), want)
        assert np.array_equal(bbf.scramble(bb), want)  # the generator's PRBS is the same sequence
    assert np.array_equal(oracle.bb_prbs(64), bbf.prbs(64))


def test_crc8_known_answers(oracle):
    # remainder of the byte string itself is zero <=> the last byte is the CRC-8 of the rest
    rng = np.random.default_rng(4)
    for n in (1, 9, 187):
        data = rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()
        c = bbf.crc8(data)
        assert oracle.crc8(data + bytes([c])) == 0
        assert oracle.crc8(data + bytes([c ^ 1])) != 0
    # the recursion yields rem(y * x^8): for y = 1 that is x^8 mod g = g - x^8 = 0xD5
    assert oracle.crc8(bytes([1])) == 0xD5


RATE_OF_KBCH = {16008: (1, "C1_4"), 3072: (0, "C1_4"), 58192: (1, "C9_10")}


def test_deheader_against_reference(oracle, ref):
    """Same calls, same bytes out, same counters as bbdeheader_bb_impl::general_work."""
    import dvbs2rx_b200 as d
    for name, kbch, calls in bb_cases.make_cases():
        fs, rate = RATE_OF_KBCH[kbch]
        o = oracle.bbdeheader(kbch)
        r = ref.bbdeheader(0, fs, d.RATE[rate], kbch)
        for bb in calls:
            a, b = o.work(bb), r.work(bb)
            assert np.array_equal(a, b), name
        assert o.counters() == r.counters(), name


def test_reference_qa_cases(oracle):
    """python/dvbs2rx/qa_bbdeheader_bb.py restated: expected outputs derived from the packet stream."""
    kbch, kb = 16008, 16008 // 8
    df = kb - 10
    rng = np.random.default_rng(21)

    def run(bb):
        return oracle.bbdeheader(kbch).work(bb)

    def expect(up, n_frames, discarded=0):
        first = -(-discarded * df // 188)  # ceil
        last = n_frames * df // 188
        return up[first:last].ravel()

    n = 10
    up = bbf.ts_packets((n * df + 187) // 188 + 1, rng)
    bb = bbf.bbframe_stream(kbch, n, up)
    assert np.array_equal(run(bb), expect(up, n))                              # test_successful_deframing
    b = bb[:2].copy()
    b[0, 9] ^= 255
    assert np.array_equal(run(b), expect(up, 2, discarded=1))                  # test_bbheader_crc_error
    for err, off in ((0x1D5, 4), (0x1D5 << 2, 4), (0x1D5 << 7, 7)):           # undetected dfl / syncd corruption
        b = bb[:2].copy()
        b[0, off] ^= (err >> 8) & 0xFF
        b[0, off + 1] ^= err & 0xFF
        assert bbf.crc8(b[0, :10]) == 0
        assert np.array_equal(run(b), expect(up, 2, discarded=1))
    dfp = (df // 188) * 188                                                    # test_padded_dfl
    upp = bbf.ts_packets(4 * (dfp // 188), rng)
    assert np.array_equal(run(bbf.bbframe_stream(kbch, 4, upp, dfl_bytes=dfp)), upp[:-1].ravel())
    b = bbf.bbframe_stream(kbch, n, up, syncd0_bits=5)                         # test_non_byte_aligned_syncd
    assert np.array_equal(run(b), expect(up, n, discarded=1))
    i_drop = 5                                                                 # test_non_consecutive_bbframes
    b = np.concatenate([bb[:i_drop], bb[i_drop + 1:]])
    pre = i_drop * df // 188
    dropped = -(-(df + i_drop * df - pre * 188) // 188)
    post = n * df // 188 - dropped - pre
    want = np.concatenate([up[:pre].ravel(), up[pre + dropped:pre + dropped + post].ravel()])
    assert np.array_equal(run(b), want)


def test_deheader_golden_fixture(oracle):
    """tests/golden/bb.json: outputs of the compiled reference (tools/gen_golden.py) for bb_cases."""
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "bb.json")))
    cases = {name: (kbch, calls) for name, kbch, calls in bb_cases.make_cases()}
    assert sorted(cases) == sorted(fx["cases"])
    for name, want in fx["cases"].items():
        kbch, calls = cases[name]
        o = oracle.bbdeheader(kbch)
        out = np.concatenate([o.work(bb) for bb in calls])
        assert out.size == want["ts_bytes"], name
        assert hashlib.sha256(out.tobytes()).hexdigest() == want["sha256"], name
        assert o.counters() == want["counters"], name
    assert hashlib.sha256(oracle.bb_prbs(7274).tobytes()).hexdigest() == fx["prbs_sha256"]
