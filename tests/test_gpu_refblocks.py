"""The reference's OWN block sources running on the GPU path.

oracle/_ref/libref_blocks.so      lib/ldpc_decoder_bb_impl.cc + lib/bch_decoder_bb_impl.cc compiled UNMODIFIED over
                                  oracle/shim (no GNU Radio in this image): the CPU blocks as they are.
oracle/_ref/libpatched_blocks.so  the same files with patches/0001-ldpc-cuda-seam.patch and 0002-bch-cuda.patch
                                  applied, built with -DDVBS2RX_WITH_B200 against libdvbs2_b200.so.
Both are built in the container by `make -C oracle blocks` (they need /root/reference) and travel to the GPU box.
The test drives the two with the same input through the reference's general_work and compares everything the
blocks expose: output bytes, items consumed, get_average_trials(), the "llr_pdu" messages, BCH counters."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BLOCKS = os.path.join(ROOT, "oracle", "_ref", "libref_blocks.so")
PATCHED_BLOCKS = os.path.join(ROOT, "oracle", "_ref", "libpatched_blocks.so")
_P = C.c_void_p


def _load(path):
    l = C.CDLL(path)
    l.blk_ldpc_create.restype = _P
    l.blk_ldpc_create.argtypes = [C.c_int] * 7
    l.blk_ldpc_destroy.argtypes = [_P]
    l.blk_ldpc_output_multiple.argtypes = [_P]
    l.blk_ldpc_work.argtypes = [_P, _P, C.c_int, C.c_int, _P]
    l.blk_ldpc_consumed.argtypes = [_P]
    l.blk_ldpc_consumed.restype = C.c_long
    l.blk_ldpc_average_trials.argtypes = [_P]
    l.blk_ldpc_average_trials.restype = C.c_uint
    l.blk_ldpc_pdu_count.argtypes = [_P]
    l.blk_ldpc_pdu.argtypes = [_P, C.c_int, _P, C.c_long, C.POINTER(C.c_long), C.POINTER(C.c_uint64)]
    l.blk_ldpc_pdu.restype = C.c_long
    l.blk_bch_create.restype = _P
    l.blk_bch_create.argtypes = [C.c_int] * 4
    l.blk_bch_destroy.argtypes = [_P]
    l.blk_bch_output_multiple.argtypes = [_P]
    l.blk_bch_work.argtypes = [_P, _P, C.c_int, C.c_int, _P]
    l.blk_bch_consumed.argtypes = [_P]
    l.blk_bch_consumed.restype = C.c_long
    l.blk_bch_frame_count.argtypes = [_P]
    l.blk_bch_frame_count.restype = C.c_uint64
    l.blk_bch_error_count.argtypes = [_P]
    l.blk_bch_error_count.restype = C.c_uint64
    return l


def test_patches_apply_to_the_reference():
    """CPU, container only: the committed patches apply cleanly to the reference checkout."""
    import subprocess
    if not os.path.isdir("/root/reference"):
        pytest.skip("no reference checkout here")
    for name in ("0001-ldpc-cuda-seam.patch", "0002-bch-cuda.patch"):
        subprocess.check_call(["git", "apply", "--check", os.path.join(ROOT, "patches", name)], cwd="/root/reference")


def _ldpc_run(l, fs, rate, om, trials, llr, n_ldpc, out_bytes_per_frame):
    h = l.blk_ldpc_create(0, fs, rate, 0, om, 0, trials)
    assert h
    mult = l.blk_ldpc_output_multiple(h)
    assert mult == out_bytes_per_frame * 32  # both blocks batch 32 frames (AVX2 here / the CUDA seam)
    frames = llr.shape[0]
    out = np.zeros(frames * out_bytes_per_frame, dtype=np.uint8)
    ret = l.blk_ldpc_work(h, llr.ctypes.data, llr.size, out.size, out.ctypes.data)
    pdus = []
    for i in range(l.blk_ldpc_pdu_count(h)):
        buf = np.zeros(32 * n_ldpc, dtype=np.uint8)
        simd, cnt = C.c_long(), C.c_uint64()
        n = l.blk_ldpc_pdu(h, i, buf.ctypes.data, buf.size, C.byref(simd), C.byref(cnt))
        pdus.append((simd.value, cnt.value, buf[:n].copy()))
    res = dict(ret=ret, out=out, consumed=l.blk_ldpc_consumed(h), avg_trials=l.blk_ldpc_average_trials(h), pdus=pdus)
    l.blk_ldpc_destroy(h)
    return res


@pytest.mark.gpu
@pytest.mark.parametrize("rate_name,fs,esn0,om", [("C1_2", 0, 1.6, 1), ("C1_2", 1, 2.0, 1), ("C3_4", 1, 4.6, 0)])
def test_reference_ldpc_block_with_cuda_seam(gpu, rate_name, fs, esn0, om):
    """lib/ldpc_decoder_bb_impl.cc:309-350,394-455 with decode == &ldpc_cuda::ldpc_dec_decode vs the unpatched block."""
    d = gpu
    from dvbs2rx_b200 import vectors
    if not (os.path.exists(REF_BLOCKS) and os.path.exists(PATCHED_BLOCKS)):
        pytest.skip("oracle/_ref block libraries not built (make -C oracle blocks, container only)")
    ref, pat = _load(REF_BLOCKS), _load(PATCHED_BLOCKS)
    rate = d.RATE[rate_name]
    frames = 64
    msg, cw, llr, info = vectors.make_llr_frames(0, fs, rate, frames, esn0, seed=77)
    llr = np.ascontiguousarray(llr)
    ob = (info.nbch if om else info.n_ldpc) // 8  # the reference's d_kldpc is the BCH n
    a = _ldpc_run(ref, fs, rate, om, 25, llr, info.n_ldpc, ob)
    b = _ldpc_run(pat, fs, rate, om, 25, llr, info.n_ldpc, ob)
    assert a["ret"] == b["ret"] == frames * ob
    assert a["consumed"] == b["consumed"] == frames * info.n_ldpc
    assert np.array_equal(a["out"], b["out"])
    assert a["avg_trials"] == b["avg_trials"]
    assert len(a["pdus"]) == len(b["pdus"]) == frames // 32
    for (s0, c0, p0), (s1, c1, p1) in zip(a["pdus"], b["pdus"]):
        assert (s0, c0) == (s1, c1) and np.array_equal(p0, p1)  # posterior LLRs of the whole SIMD batch


@pytest.mark.gpu
@pytest.mark.parametrize("rate_name,fs", [("C1_2", 1), ("C2_3", 0), ("C9_10", 1)])
def test_reference_bch_block_with_cuda(gpu, oracle, rate_name, fs):
    """lib/bch_decoder_bb_impl.cc:84-117 with dvbs2b200_bch_decode inside vs the unpatched block."""
    d = gpu
    if not (os.path.exists(REF_BLOCKS) and os.path.exists(PATCHED_BLOCKS)):
        pytest.skip("oracle/_ref block libraries not built (make -C oracle blocks, container only)")
    ref, pat = _load(REF_BLOCKS), _load(PATCHED_BLOCKS)
    rate = d.RATE[rate_name]
    info = d.lookup(0, fs, rate)
    rng = np.random.default_rng(5)
    nerr = [0, 1, 2, 5, info.t, info.t + 1, 40, 0] * 4
    F = len(nerr)
    msg = rng.integers(0, 256, size=(F, info.kbch // 8), dtype=np.uint8)
    cw = oracle.bch_encode(oracle.bch(fs, info.t, info.nbch), msg)
    for f in range(F):
        for p in rng.choice(info.nbch, size=nerr[f], replace=False):
            cw[f, p >> 3] ^= 0x80 >> (p & 7)
    cw = np.ascontiguousarray(cw)
    res = []
    for l in (ref, pat):
        h = l.blk_bch_create(0, fs, rate, 1)
        assert h and l.blk_bch_output_multiple(h) == info.kbch // 8
        out = np.zeros(F * (info.kbch // 8), dtype=np.uint8)
        ret = l.blk_bch_work(h, cw.ctypes.data, cw.size, out.size, out.ctypes.data)
        res.append((ret, out, l.blk_bch_consumed(h), l.blk_bch_frame_count(h), l.blk_bch_error_count(h)))
        l.blk_bch_destroy(h)
    assert res[0][0] == res[1][0] and res[0][2:] == res[1][2:]
    assert np.array_equal(res[0][1], res[1][1])
    assert res[0][4] == sum(1 for e in nerr if e > info.t)
