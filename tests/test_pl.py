"""PL descrambler + pilot-segment de-rotation: scrambling codes against the compiled reference and
the golden fixture (bit-exact), the GPU kernel against the oracle's serial restatement of handle_payload (tolerance:
the reference's rotator is a serial float recurrence, VOLK, unpinned)."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_MAX = 360 * 90 + 22 * 36


def _orc(oracle):
    l = oracle.l
    l.orc_pl_rn.argtypes = [C.c_int, C.c_void_p, C.c_int]
    l.orc_pl_payload.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_void_p]
    return l


def orc_rn(oracle, gold, n=N_MAX):
    rn = np.zeros(n, np.uint8)
    _orc(oracle).orc_pl_rn(gold, rn.ctypes.data, n)
    return rn


def test_scrambling_codes_match_golden_fixture(oracle, built):
    """oracle and the library's host function vs tests/golden/pl.json (from lib/pl_descrambler.cc, compiled)."""
    import dvbs2rx_b200 as d
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "pl.json")))
    assert g["n"] == N_MAX
    for gold, want in g["gold_codes"].items():
        for rn in (orc_rn(oracle, int(gold)), d.pl_scrambling_codes(int(gold), N_MAX)):
            assert hashlib.sha256(rn.tobytes()).hexdigest() == want["sha256"], gold
            assert rn[:64].tolist() == want["first64"]


@pytest.mark.ref
def test_scrambling_codes_match_compiled_reference(oracle, ref):
    ref.l.ref_pl_rn.argtypes = [C.c_int, C.c_void_p, C.c_int]
    for gold in (0, 5, 77, 4242, 200000):
        rn = np.zeros(N_MAX, np.uint8)
        assert ref.l.ref_pl_rn(gold, rn.ctypes.data, N_MAX) == 0
        assert np.array_equal(rn, orc_rn(oracle, gold))


def test_oracle_payload_layout(oracle):
    """No rotation, no scrambling-dependent surprises: descrambling a scrambled constant gives the constant back, and the
    pilot blocks are dropped at the right places."""
    n_slots, gold = 40, 3
    n_pil = (n_slots - 1) // 16
    plen = n_slots * 90 + n_pil * 36
    rn = orc_rn(oracle, gold, plen)
    lut = np.array([1, -1j, -1, 1j])  # descrambling factors; scrambling is their conjugate
    data = (np.arange(n_slots * 90) % 7 + 1) * np.exp(1j * 0.3)
    payload = np.zeros(plen, np.complex64)
    pos = np.arange(n_slots * 90) + 36 * (np.arange(n_slots * 90) // 1440)
    payload[pos] = data
    payload[np.setdiff1d(np.arange(plen), pos)] = 99.0  # pilots
    payload = (payload * np.conj(lut[rn])).astype(np.complex64)
    out = np.zeros((n_slots * 90, 2), np.float32)
    pil = np.zeros(22, np.float32)
    _orc(oracle).orc_pl_payload(payload.view(np.float32).ctypes.data, n_slots, 1, rn.ctypes.data, C.c_float(0.0), C.c_float(0.0), 0,
                                pil.ctypes.data, out.ctypes.data)
    assert np.allclose(out[:, 0] + 1j * out[:, 1], data, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("n_slots,has_pilots,gold", [(360, 1, 0), (360, 0, 0), (144, 1, 13), (90, 1, 262141), (17, 1, 1), (16, 1, 1)])
def test_pl_kernel_matches_oracle(gpu, oracle, n_slots, has_pilots, gold):
    d = gpu
    rng = np.random.default_rng(7 + n_slots)
    F = 5
    plen = d.PlDescrambler.payload_len(n_slots, has_pilots)
    assert plen == n_slots * 90 + (((n_slots - 1) // 16) * 36 if has_pilots else 0)
    payload = (rng.standard_normal((F, plen, 2)) * 0.8).astype(np.float32)
    info = np.zeros(F, dtype=d.PL_FRAME_DTYPE)
    info["plheader_phase"] = rng.uniform(-3.1, 3.1, F)
    info["fine_foffset"] = rng.uniform(-3e-4, 3e-4, F)
    info["coarse_corrected"] = [1, 0, 1, 1, 0]
    info["pilot_phase"] = rng.uniform(-3.1, 3.1, (F, 22))
    pl = d.PlDescrambler(gold)
    got = pl.process(payload, n_slots, has_pilots, info)
    rn = orc_rn(oracle, gold, plen)
    assert np.array_equal(rn, d.pl_scrambling_codes(gold, plen))
    l = _orc(oracle)
    for f in range(F):
        want = np.zeros((n_slots * 90, 2), np.float32)
        pil = np.ascontiguousarray(info["pilot_phase"][f])
        l.orc_pl_payload(payload[f].ctypes.data, n_slots, has_pilots, rn.ctypes.data, C.c_float(info["plheader_phase"][f]),
                         C.c_float(info["fine_foffset"][f]), int(info["coarse_corrected"][f]), pil.ctypes.data, want.ctypes.data)
        # tolerance: 1440 serial float complex multiplies in the reference's rotator vs one closed-form phase here
        assert np.allclose(got[f], want, rtol=0, atol=2e-4), (f, np.abs(got[f] - want).max())
    pl.close()


@pytest.mark.gpu
def test_pl_then_demap_then_decode(gpu, oracle):
    """The stage in its place: scrambled, rotated PLFRAME payloads with pilots -> XFECFRAMEs -> demap -> LDPC -> BCH."""
    d = gpu
    from dvbs2rx_b200 import vectors
    rng = np.random.default_rng(3)
    F, gold, n_slots = 4, 5, 360
    msg, cw, info = vectors.encode_frames(0, 1, d.C1_2, F, rng)
    sym = vectors.map_symbols(cw, d.MOD_QPSK, d.C1_2)
    iq, n0 = vectors.awgn(sym, 4.0, rng)
    x = iq[..., 0] + 1j * iq[..., 1]
    plen = d.PlDescrambler.payload_len(n_slots, True)
    rn = d.pl_scrambling_codes(gold, plen)
    scr = np.array([1, 1j, -1, -1j])[rn]  # exp(j R_n pi / 2)
    pos = np.arange(n_slots * 90) + 36 * (np.arange(n_slots * 90) // 1440)
    frames_info = np.zeros(F, dtype=d.PL_FRAME_DTYPE)
    payload = np.zeros((F, plen), np.complex64)
    for f in range(F):
        ph0, fo = rng.uniform(-3, 3), rng.uniform(-2e-4, 2e-4)
        pil = rng.uniform(-3, 3, 22).astype(np.float32)
        k = np.arange(n_slots * 90) % 1440
        seg = np.arange(n_slots * 90) // 1440
        base = np.where(seg == 0, ph0, pil[np.maximum(seg - 1, 0)])
        rot = np.exp(1j * (base + 2 * np.pi * fo * k))  # what the channel did, as the synchroniser estimated it
        payload[f, pos] = x[f] * rot
        payload[f] *= scr
        frames_info[f] = (ph0, fo, 1, 0, pil)
    pl = d.PlDescrambler(gold)
    xfec = pl.process(payload.view(np.float32).reshape(F, plen, 2), n_slots, True, frames_info)
    assert np.allclose(xfec, iq, atol=3e-4)
    code = d.Code(0, 1, d.C1_2)
    out, trials, corr = code.fec_decode(iq=xfec, n0=n0, constellation=d.MOD_QPSK, max_trials=25)
    assert (trials >= 0).all() and np.array_equal(out, msg)
    code.close()
    pl.close()
