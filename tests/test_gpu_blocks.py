"""GPU tests of the host-side block mirrors (gr-dvbs2rx_b200/host): the reference's block contracts
(general_work sizes, counters, llr_pdu, get_average_trials / get_snr semantics) on top of the C ABI."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_P = C.c_void_p


@pytest.fixture(scope="module")
def blk(gpu):
    l = C.CDLL(os.path.join(ROOT, "gr-dvbs2rx_b200", "libdvbs2rx_b200_blocks.so"))
    l.blk_ldpc_make.restype = _P
    l.blk_ldpc_make.argtypes = [C.c_int] * 7
    l.blk_ldpc_free.argtypes = [_P]
    l.blk_ldpc_output_multiple.argtypes = [_P]
    l.blk_ldpc_forecast.argtypes = [_P, C.c_int]
    l.blk_ldpc_work.argtypes = [_P, C.c_int, _P, C.c_int, _P, C.POINTER(C.c_int)]
    l.blk_ldpc_average_trials.argtypes = [_P]
    l.blk_ldpc_pdu_bytes.argtypes = [_P, _P, C.c_size_t]
    l.blk_ldpc_pdu_bytes.restype = C.c_size_t
    l.blk_bch_make.restype = _P
    l.blk_bch_make.argtypes = [C.c_int] * 3
    l.blk_bch_free.argtypes = [_P]
    l.blk_bch_work.argtypes = [_P, C.c_int, _P, _P, C.POINTER(C.c_int)]
    l.blk_bch_frame_count.argtypes = [_P]
    l.blk_bch_frame_count.restype = C.c_uint64
    l.blk_bch_error_count.argtypes = [_P]
    l.blk_bch_error_count.restype = C.c_uint64
    l.blk_bbdescrambler_make.restype = _P
    l.blk_bbdescrambler_make.argtypes = [C.c_int] * 3
    l.blk_bbdescrambler_free.argtypes = [_P]
    l.blk_bbdescrambler_work.argtypes = [_P, C.c_int, _P, _P]
    l.blk_bbdeheader_make.restype = _P
    l.blk_bbdeheader_make.argtypes = [C.c_int] * 3
    l.blk_bbdeheader_free.argtypes = [_P]
    l.blk_bbdeheader_work.argtypes = [_P, C.c_int, C.c_int, _P, _P, C.POINTER(C.c_int)]
    l.blk_bbdeheader_forecast.argtypes = [_P, C.c_int]
    l.blk_bbdeheader_counters.argtypes = [_P, _P]
    l.blk_demap_make.restype = _P
    l.blk_demap_make.argtypes = [C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]
    l.blk_demap_make_apsk.restype = _P
    l.blk_demap_make_apsk.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float, C.c_char_p, C.c_int]
    l.blk_apsk_points.argtypes = [C.c_int, C.c_int, _P, C.c_int]
    l.blk_demap_free.argtypes = [_P]
    l.blk_demap_work.argtypes = [_P, C.c_int, _P, _P, C.POINTER(C.c_int)]
    l.blk_demap_snr.argtypes = [_P]
    l.blk_demap_snr.restype = C.c_float
    l.blk_demap_llr_pdu.argtypes = [_P, C.c_long, C.c_uint64, _P, C.c_size_t]
    l.blk_ldpc_cuda_init.argtypes = [C.c_int] * 4
    l.blk_ldpc_cuda_decode.argtypes = [_P, C.c_int]
    return l


def test_ldpc_block_contract(blk, gpu, oracle):
    d = gpu
    from dvbs2rx_b200 import vectors
    msg, cw, llr, info = vectors.make_llr_frames(0, 0, d.C1_2, 64, 1.7, seed=21)
    h = blk.blk_ldpc_make(0, 0, d.C1_2, d.OM_MESSAGE, 0, 2, 1)  # max_trials 0 -> 25, 2 batches per call
    assert h
    kb = info.nbch // 8
    assert blk.blk_ldpc_output_multiple(h) == kb * 64
    assert blk.blk_ldpc_forecast(h, kb * 64) == 64 * 16200
    out = np.zeros((64, kb), np.uint8)
    consumed = C.c_int()
    produced = blk.blk_ldpc_work(h, kb * 64, llr.ctypes.data, llr.size, out.ctypes.data, C.byref(consumed))
    assert produced == kb * 64 and consumed.value == 64 * 16200
    o_post, o_ret = oracle.ldpc_decode(info.table, llr, 25, lanes=32)
    assert np.array_equal(out, oracle.pack_hard(o_post, info.nbch))
    used = [25 if r < 0 else 25 - r for r in o_ret[::32]]
    assert blk.blk_ldpc_average_trials(h) == sum(used) // 2  # per batch, integer division as the reference
    pdu = np.zeros((64, 16200), np.int8)
    assert blk.blk_ldpc_pdu_bytes(h, pdu.ctypes.data, pdu.size) == pdu.size
    assert np.array_equal(pdu, o_post)  # what the reference publishes on "llr_pdu"
    blk.blk_ldpc_free(h)


def test_ldpc_cuda_seam_in_place(blk, gpu, oracle):
    """int (*decode)(void*, int8_t*, int) of lib/ldpc_decoder_bb_impl.h:41: posteriors overwrite `code`."""
    d = gpu
    from dvbs2rx_b200 import vectors
    msg, cw, llr, info = vectors.make_llr_frames(0, 0, d.C2_3, 32, 3.3, seed=4)
    assert blk.blk_ldpc_cuda_init(0, 0, d.C2_3, 32) == 0
    code = llr.copy()
    ret = blk.blk_ldpc_cuda_decode(code.ctypes.data, 25)
    o_post, o_ret = oracle.ldpc_decode(info.table, llr, 25, lanes=32)
    assert ret == o_ret[0]
    assert np.array_equal(code, o_post)
    blk.blk_ldpc_cuda_shutdown()


def test_bch_block_counters(blk, gpu, oracle):
    d = gpu
    info = d.lookup(0, 0, d.C2_3)
    hb = oracle.bch(0, info.t, info.nbch)
    rng = np.random.default_rng(8)
    msg = rng.integers(0, 256, size=(10, info.kbch // 8), dtype=np.uint8)
    cw = oracle.bch_encode(hb, msg)
    for f, ne in enumerate([0, 0, 1, 5, 12, 13, 20, 0, 2, 40]):
        for p in rng.choice(info.nbch, size=ne, replace=False):
            cw[f, p >> 3] ^= 0x80 >> (p & 7)
    h = blk.blk_bch_make(0, 0, d.C2_3)
    out = np.zeros_like(msg)
    consumed = C.c_int()
    r = blk.blk_bch_work(h, msg.size, cw.ctypes.data, out.ctypes.data, C.byref(consumed))
    assert r == msg.size and consumed.value == cw.size
    o_out, o_ret = oracle.bch_decode(hb, cw)
    assert np.array_equal(out, o_out)
    assert blk.blk_bch_frame_count(h) == 10
    assert blk.blk_bch_error_count(h) == int((o_ret == -1).sum()) == 3
    blk.blk_bch_free(h)


def test_demapper_block(blk, gpu, oracle):
    d = gpu
    from dvbs2rx_b200 import vectors
    err = C.create_string_buffer(128)
    assert not blk.blk_demap_make(1, d.C2_3, d.MOD_16APSK, err, 128)
    assert err.value == b"Unsupported constellation"  # lib/xfecframe_demapper_cb_impl.cc:70-72
    rng = np.random.default_rng(12)
    msg, cw, info = vectors.encode_frames(0, 1, d.C1_2, 3, rng)
    iq, n0 = vectors.awgn(vectors.map_symbols(cw, d.MOD_QPSK, d.C1_2), 6.0, rng)
    h = blk.blk_demap_make(1, d.C1_2, d.MOD_QPSK, err, 128)
    out = np.zeros((3, 64800), np.int8)
    consumed = C.c_int()
    assert blk.blk_demap_work(h, out.size, iq.ctypes.data, out.ctypes.data, C.byref(consumed)) == out.size
    assert consumed.value == 3 * 32400
    snr = blk.blk_demap_snr(h)
    assert abs(snr - 6.0) < 0.5  # hard-slice estimate (lib/qa_qpsk.cc:105-153 allows 10 %)
    # the LLRs are the oracle's for the N0 the block estimated for that frame (last frame's is visible)
    n0_last = np.float32(1.0) / np.float32(10.0 ** (np.float64(snr) / 10))
    ref_last = oracle.demap_qpsk(iq[2:3], n0_last)
    assert np.abs(out[2].astype(int) - ref_last[0]).max() <= 1
    # post-decoder refinement from posterior LLR signs: exact reference symbols -> estimate ~ true Es/N0
    post = np.where(cw == 1, -20, 20).astype(np.int8)
    pdu = np.concatenate([post, np.zeros((29, 64800), np.int8)])
    blk.blk_demap_llr_pdu(h, 32, 0, pdu.ctypes.data, pdu.size)
    assert abs(blk.blk_demap_snr(h) - 6.0) < 0.1
    blk.blk_demap_free(h)


def test_bb_blocks_against_oracle(blk, gpu, oracle):
    """bbdescrambler_bb::work and bbdeheader_bb::general_work / forecast / counters, driven the way the GNU
    Radio scheduler drives the reference blocks (python/dvbs2rx/qa_bbdeheader_bb.py: QPSK 1/4 normal)."""
    d = gpu
    from dvbs2rx_b200 import bbframes as bbf
    kbch, kb = 16008, 16008 // 8
    rng = np.random.default_rng(12)
    n = 10
    up = bbf.ts_packets((n * (kb - 10) + 187) // 188 + 1, rng)
    bb = bbf.bbframe_stream(kbch, n, up)
    bb[4, 9] ^= 0xFF  # one BBHEADER fails its CRC
    scr = bbf.scramble(bb)
    hs = blk.blk_bbdescrambler_make(0, 1, d.C1_4)
    out = np.zeros_like(scr)
    assert blk.blk_bbdescrambler_work(hs, scr.size, scr.ctypes.data, out.ctypes.data) == scr.size
    assert np.array_equal(out, bb) and np.array_equal(out, oracle.bb_descramble(scr, kbch))
    blk.blk_bbdescrambler_free(hs)
    hd = blk.blk_bbdeheader_make(0, 1, d.C1_4)
    max_dfl_bytes = (kbch - 80) // 8
    assert blk.blk_bbdeheader_forecast(hd, 2 * max_dfl_bytes) == 2 * kb  # lib/bbdeheader_bb_impl.cc:69-74
    o = oracle.bbdeheader(kbch)
    got, consumed = [], C.c_int()
    for f0 in range(0, n, 2):  # the scheduler hands over two BBFRAMEs' worth of output space at a time
        part = np.ascontiguousarray(out[f0:f0 + 2])
        ts = np.zeros(2 * max_dfl_bytes + 188, np.uint8)
        r = blk.blk_bbdeheader_work(hd, 2 * max_dfl_bytes, part.size, part.ctypes.data, ts.ctypes.data, C.byref(consumed))
        assert consumed.value == 2 * kb
        want = o.work(part)
        assert r == want.size and np.array_equal(ts[:r], want)
        got.append(ts[:r].copy())
    c = (C.c_uint64 * 5)()
    blk.blk_bbdeheader_counters(hd, c)
    assert dict(zip(("packets", "errors", "bbframes", "dropped", "gaps"), [int(v) for v in c])) == o.counters()
    assert o.counters()["dropped"] == 1
    blk.blk_bbdeheader_free(hd)


def test_demapper_block_apsk_opt_in(blk, gpu):
    """16APSK / 32APSK are rejected by default (as the reference throws) and accepted behind the explicit opt-in:
    the block then demaps with the table-driven demapper and the frames decode through LDPC + BCH."""
    d = gpu
    from dvbs2rx_b200 import apsk, vectors
    err = C.create_string_buffer(128)
    rng = np.random.default_rng(19)
    for mod, fs, rate_name, esn0, pts_py in ((d.MOD_16APSK, 0, "C2_3", 10.5, apsk.points_16apsk(apsk.GAMMA_16APSK["C2_3"])),
                                             (d.MOD_32APSK, 1, "C9_10", 17.5, apsk.points_32apsk(*apsk.GAMMA_32APSK["C9_10"]))):
        rate = d.RATE[rate_name]
        assert not blk.blk_demap_make(fs, rate, mod, err, 128) and err.value == b"Unsupported constellation"
        # the C++ tables are the Python ones
        pts = np.zeros((32, 2), np.float32)
        n = blk.blk_apsk_points(mod, rate, pts.ctypes.data, pts.size)
        assert n == pts_py.shape[0] and np.allclose(pts[:n], pts_py, atol=1e-6)
        bits = n.bit_length() - 1
        F = 3
        msg, cw, info = vectors.encode_frames(0, fs, rate, F, rng)
        offs = apsk.row_offsets(info.n_ldpc, bits)
        iq, n0 = vectors.awgn(apsk.map_bits(cw, pts_py, offs), esn0, rng)
        iq = np.ascontiguousarray(iq)
        h = blk.blk_demap_make_apsk(fs, rate, mod, esn0, err, 128)
        assert h, err.value
        out = np.zeros((F, info.n_ldpc), np.int8)
        consumed = C.c_int()
        assert blk.blk_demap_work(h, out.size, iq.ctypes.data, out.ctypes.data, C.byref(consumed)) == out.size
        assert consumed.value == F * (info.n_ldpc // bits)
        code = d.Code(0, fs, rate)
        assert np.array_equal(out, code.demap_table(pts_py, offs, iq, n0))  # the block is a thin shell over the C ABI
        dec, trials, corr = code.fec_decode(llr=out, max_trials=25)
        assert (trials >= 0).all() and np.array_equal(dec, msg)
        code.close()
        blk.blk_demap_free(h)
