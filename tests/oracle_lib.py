"""ctypes access to the checker libraries (tests and bench.py's CPU legs only):
oracle/liboracle.so (C restatement) and oracle/_ref/libdvbs2_ref.so (compiled reference)."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_PATH = os.path.join(ROOT, "oracle", "liboracle.so")
REF_PATH = os.path.join(ROOT, "oracle", "_ref", "libdvbs2_ref.so")
_P = C.c_void_p

PRIM_POLY = {1: 0x1002D, 0: 0x402B, 2: 0x802D}  # framesize ordinal -> lib/bch_decoder_bb_impl.cc:58-63


def _p(a):
    return None if a is None else a.ctypes.data


class OracleDeheader:
    """Stateful restatement of bbdeheader_bb (oracle/dvbs2_oracle.c)."""

    def __init__(self, l, kbch):
        self.l, self.kbch = l, kbch
        self.h = l.orc_bbdeheader_create(kbch)

    def work(self, bbframes):
        bb = np.ascontiguousarray(bbframes, dtype=np.uint8).reshape(-1, self.kbch // 8)
        out = np.zeros(bb.shape[0] * (self.kbch // 8 + 188) + 188, dtype=np.uint8)
        n = self.l.orc_bbdeheader_work(self.h, bb.ctypes.data, bb.shape[0], out.ctypes.data)
        return out[:n].copy()

    def counters(self):
        c = (C.c_uint64 * 5)()
        self.l.orc_bbdeheader_counters(self.h, c)
        return dict(zip(("packets", "errors", "bbframes", "dropped", "gaps"), [int(v) for v in c]))

    def __del__(self):
        try:
            self.l.orc_bbdeheader_destroy(self.h)
        except Exception:
            pass


class Oracle:
    def __init__(self):
        l = self.l = C.CDLL(ORACLE_PATH)
        l.orc_table_name.restype = C.c_char_p
        l.orc_ldpc_create.restype = _P
        l.orc_ldpc_create.argtypes = [C.c_int]
        l.orc_ldpc_destroy.argtypes = [_P]
        l.orc_ldpc_decode.argtypes = [_P, _P, C.c_int, C.c_int]
        l.orc_ldpc_bad.argtypes = [_P, _P, C.c_int]
        l.orc_ldpc_encode.argtypes = [_P, _P, _P]
        l.orc_pack_hard.argtypes = [_P, C.c_int, _P]
        l.orc_lookup.argtypes = [C.c_int] * 3 + [C.POINTER(C.c_int)] * 3
        l.orc_bch_create.restype = _P
        l.orc_bch_create.argtypes = [C.c_uint32, C.c_int, C.c_int]
        l.orc_bch_destroy.argtypes = [_P]
        l.orc_bch_n.argtypes = [_P]
        l.orc_bch_k.argtypes = [_P]
        l.orc_bch_genpoly.argtypes = [_P, _P, C.c_int]
        l.orc_gf_min_poly.argtypes = [_P, C.c_uint32]
        l.orc_gf_min_poly.restype = C.c_uint32
        l.orc_gf_alpha.argtypes = [_P, C.c_uint32]
        l.orc_gf_alpha.restype = C.c_uint32
        l.orc_bch_encode.argtypes = [_P, _P, _P]
        l.orc_bch_decode.argtypes = [_P, _P, _P]
        l.orc_bch_syndrome.argtypes = [_P, _P, _P]
        l.orc_bch_err_loc_poly.argtypes = [_P, _P, _P]
        l.orc_bch_err_loc_numbers.argtypes = [_P, _P, C.c_int, _P]
        l.orc_demap_qpsk.argtypes = [_P, C.c_int, C.c_float, _P]
        l.orc_demap_8psk.argtypes = [_P, C.c_int, C.c_float, C.c_int, _P]
        l.orc_snr_qpsk.argtypes = [_P, C.c_int, _P]
        l.orc_snr_qpsk.restype = C.c_float
        l.orc_snr_8psk.argtypes = [_P, C.c_int, _P, C.c_int]
        l.orc_snr_8psk.restype = C.c_float
        l.orc_bb_prbs.argtypes = [_P, C.c_int]
        l.orc_bb_descramble.argtypes = [_P, C.c_int, C.c_int, _P]
        l.orc_crc8.argtypes = [_P, C.c_int]
        l.orc_crc8.restype = C.c_uint8
        l.orc_bbdeheader_create.restype = _P
        l.orc_bbdeheader_create.argtypes = [C.c_int]
        l.orc_bbdeheader_destroy.argtypes = [_P]
        l.orc_bbdeheader_work.argtypes = [_P, _P, C.c_int, _P]
        l.orc_bbdeheader_work.restype = C.c_long
        l.orc_bbdeheader_counters.argtypes = [_P, _P]
        self._ldpc = {}
        self._bch = {}

    def estimate_snr(self, constellation, iq, llr=None, rate=0):
        """Linear Es/N0 per frame (constellation 0 = QPSK, 4 = 8PSK); llr: posterior LLRs or None."""
        iq = np.ascontiguousarray(iq, dtype=np.float32)
        F, n_syms = iq.shape[0], iq.shape[1]
        out = np.zeros(F, dtype=np.float32)
        for f in range(F):
            lp = None if llr is None else np.ascontiguousarray(llr[f], dtype=np.int8)
            if constellation == 0:
                out[f] = self.l.orc_snr_qpsk(iq[f].ctypes.data, n_syms, _p(lp))
            else:
                out[f] = self.l.orc_snr_8psk(iq[f].ctypes.data, n_syms, _p(lp), rate)
        return out

    # ---- BB layer ----
    def bb_prbs(self, nbytes):
        seq = np.zeros(nbytes, dtype=np.uint8)
        self.l.orc_bb_prbs(seq.ctypes.data, nbytes)
        return seq

    def bb_descramble(self, bbframes, kbch):
        bb = np.ascontiguousarray(bbframes, dtype=np.uint8).reshape(-1, kbch // 8)
        out = np.empty_like(bb)
        self.l.orc_bb_descramble(bb.ctypes.data, bb.shape[0], kbch // 8, out.ctypes.data)
        return out

    def crc8(self, data):
        a = np.ascontiguousarray(np.frombuffer(bytes(data), dtype=np.uint8))
        return int(self.l.orc_crc8(a.ctypes.data, a.size))

    def bbdeheader(self, kbch):
        return OracleDeheader(self.l, kbch)

    def lookup(self, standard, framesize, rate):
        k, n, t = C.c_int(), C.c_int(), C.c_int()
        tab = self.l.orc_lookup(standard, framesize, rate, C.byref(k), C.byref(n), C.byref(t))
        return tab, k.value, n.value, t.value

    def table_name(self, table):
        return self.l.orc_table_name(table).decode()

    def ldpc(self, table):
        if table not in self._ldpc:
            self._ldpc[table] = self.l.orc_ldpc_create(table)
        return self._ldpc[table]

    def ldpc_decode(self, table, llr, trials=25, lanes=None):
        """llr [F, N] int8.  lanes=None: every frame alone (per-frame termination); else frames
        are decoded in consecutive groups of `lanes` with the reference's coupled loop.
        Returns (posterior [F, N], ret [F])."""
        N = self.l.orc_table_n(table)
        post = np.array(llr, dtype=np.int8, copy=True).reshape(-1, N)
        F = post.shape[0]
        g = lanes or 1
        assert F % g == 0
        ret = np.zeros(F, dtype=np.int32)
        h = self.ldpc(table)
        for f0 in range(0, F, g):
            blk = np.ascontiguousarray(post[f0:f0 + g])
            ret[f0:f0 + g] = self.l.orc_ldpc_decode(h, blk.ctypes.data, g, trials)
            post[f0:f0 + g] = blk
        return post, ret

    def ldpc_bad(self, table, llr):
        N = self.l.orc_table_n(table)
        a = np.ascontiguousarray(llr, dtype=np.int8).reshape(-1, N)
        return self.l.orc_ldpc_bad(self.ldpc(table), a.ctypes.data, a.shape[0])

    def ldpc_encode(self, table, msg_bits):
        N, K = self.l.orc_table_n(table), self.l.orc_table_k(table)
        msg_bits = np.ascontiguousarray(msg_bits, dtype=np.uint8).reshape(-1, K)
        out = np.zeros((msg_bits.shape[0], N), dtype=np.uint8)
        for f in range(msg_bits.shape[0]):
            self.l.orc_ldpc_encode(self.ldpc(table), msg_bits[f].ctypes.data, out[f].ctypes.data)
        return out

    def pack_hard(self, llr, nbits):
        llr = np.ascontiguousarray(llr, dtype=np.int8)
        F = llr.shape[0]
        out = np.zeros((F, nbits // 8), dtype=np.uint8)
        for f in range(F):
            self.l.orc_pack_hard(llr[f].ctypes.data, nbits, out[f].ctypes.data)
        return out

    def bch(self, framesize, t, n):
        key = (framesize, t, n)
        if key not in self._bch:
            self._bch[key] = self.l.orc_bch_create(PRIM_POLY[framesize], t, n)
        return self._bch[key]

    def bch_raw(self, prim_poly, t, n):
        return self.l.orc_bch_create(prim_poly, t, n)

    def bch_encode(self, h, msg):
        k, n = self.l.orc_bch_k(h), self.l.orc_bch_n(h)
        msg = np.ascontiguousarray(msg, dtype=np.uint8).reshape(-1, k // 8)
        cw = np.zeros((msg.shape[0], n // 8), dtype=np.uint8)
        for f in range(msg.shape[0]):
            self.l.orc_bch_encode(h, msg[f].ctypes.data, cw[f].ctypes.data)
        return cw

    def bch_decode(self, h, cw):
        k, n = self.l.orc_bch_k(h), self.l.orc_bch_n(h)
        cw = np.ascontiguousarray(cw, dtype=np.uint8).reshape(-1, n // 8)
        msg = np.zeros((cw.shape[0], k // 8), dtype=np.uint8)
        ret = np.zeros(cw.shape[0], dtype=np.int32)
        for f in range(cw.shape[0]):
            ret[f] = self.l.orc_bch_decode(h, cw[f].ctypes.data, msg[f].ctypes.data)
        return msg, ret

    def demap_qpsk(self, iq, n0):
        iq = np.ascontiguousarray(iq, dtype=np.float32)
        F = iq.shape[0]
        n_syms = iq[0].size // 2
        out = np.zeros((F, 2 * n_syms), dtype=np.int8)
        n0 = np.broadcast_to(np.asarray(n0, dtype=np.float32), (F,))
        for f in range(F):
            self.l.orc_demap_qpsk(iq[f].ctypes.data, n_syms, C.c_float(n0[f]), out[f].ctypes.data)
        return out

    def demap_8psk(self, iq, n0, rate):
        iq = np.ascontiguousarray(iq, dtype=np.float32)
        F = iq.shape[0]
        n_syms = iq[0].size // 2
        out = np.zeros((F, 3 * n_syms), dtype=np.int8)
        n0 = np.broadcast_to(np.asarray(n0, dtype=np.float32), (F,))
        for f in range(F):
            self.l.orc_demap_8psk(iq[f].ctypes.data, n_syms, C.c_float(n0[f]), rate, out[f].ctypes.data)
        return out


class RefDeheader:
    """bbdeheader_bb_impl of the reference, compiled unmodified (oracle/ref_bb_harness.cc)."""

    def __init__(self, l, standard, framesize, rate, kbch):
        self.l, self.kbch = l, kbch
        self.h = l.ref_bbdeheader_create(standard, framesize, rate)
        assert self.h

    def work(self, bbframes):
        bb = np.ascontiguousarray(bbframes, dtype=np.uint8).reshape(-1, self.kbch // 8)
        out = np.zeros(bb.shape[0] * (self.kbch // 8 + 188) + 188, dtype=np.uint8)
        n = self.l.ref_bbdeheader_work(self.h, bb.ctypes.data, bb.size, out.ctypes.data, out.size - 188)
        return out[:n].copy()

    def counters(self):
        c = (C.c_uint64 * 5)()
        self.l.ref_bbdeheader_counters(self.h, c)
        return dict(zip(("packets", "errors", "bbframes", "dropped", "gaps"), [int(v) for v in c]))

    def __del__(self):
        try:
            self.l.ref_bbdeheader_destroy(self.h)
        except Exception:
            pass


class Ref:
    """The compiled, unmodified reference (oracle/_ref)."""

    ISA = {"generic": 0, "sse41": 1, "avx2": 2}

    def __init__(self):
        l = self.l = C.CDLL(REF_PATH)
        l.ref_ldpc_init.argtypes = [C.c_char_p, C.c_int]
        l.ref_ldpc_decode.argtypes = [C.c_int, C.c_int, _P, C.c_int]
        l.ref_ldpc_decode_mt.argtypes = [C.c_char_p, _P, C.c_int, C.c_int, C.c_int, _P]
        l.ref_ldpc_decode_mt.restype = C.c_double
        l.ref_bch_create.restype = _P
        l.ref_bch_create.argtypes = [C.c_uint32, C.c_int, C.c_int]
        l.ref_bch_destroy.argtypes = [_P]
        l.ref_bch_k.argtypes = [_P]
        l.ref_gf_alpha.argtypes = [_P, C.c_uint32]
        l.ref_gf_alpha.restype = C.c_uint32
        l.ref_gf_min_poly.argtypes = [_P, C.c_uint32]
        l.ref_gf_min_poly.restype = C.c_uint32
        l.ref_bch_genpoly.argtypes = [_P, _P, C.c_int]
        l.ref_bch_encode.argtypes = [_P, _P, _P]
        l.ref_bch_decode.argtypes = [_P, _P, _P]
        l.ref_bch_syndrome.argtypes = [_P, _P, _P]
        l.ref_bch_err_loc_poly.argtypes = [_P, _P, _P]
        l.ref_bch_decode_mt.argtypes = [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]
        l.ref_bch_decode_mt.restype = C.c_double
        l.ref_demap_8psk.argtypes = [_P, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, _P]
        l.ref_demap_qpsk_psk4.argtypes = [_P, C.c_int, C.c_float, _P]
        l.ref_snr_8psk.argtypes = [_P, C.c_int, _P, C.c_int, C.c_int, C.c_int]
        l.ref_snr_8psk.restype = C.c_float
        l.ref_bb_descramble.argtypes = [C.c_int, C.c_int, C.c_int, _P, C.c_int, _P]
        l.ref_bbdeheader_create.restype = _P
        l.ref_bbdeheader_create.argtypes = [C.c_int, C.c_int, C.c_int]
        l.ref_bbdeheader_destroy.argtypes = [_P]
        l.ref_bbdeheader_work.argtypes = [_P, _P, C.c_int, _P, C.c_int]
        l.ref_bbdeheader_counters.argtypes = [_P, _P]
        self._init = {}

    def snr_8psk(self, iq, llr, rows):
        """lib/psk.hh hard()/map() driven as the demapper block does; rows = (r0, r1, r2)."""
        iq = np.ascontiguousarray(iq, dtype=np.float32)
        out = np.zeros(iq.shape[0], dtype=np.float32)
        for f in range(iq.shape[0]):
            lp = None if llr is None else np.ascontiguousarray(llr[f], dtype=np.int8)
            out[f] = self.l.ref_snr_8psk(iq[f].ctypes.data, iq.shape[1], _p(lp), *rows)
        return out

    # ---- BB layer: the reference's own blocks over the gr::block shim ----
    def bb_descramble(self, standard, framesize, rate, bbframes):
        bb = np.ascontiguousarray(bbframes, dtype=np.uint8).ravel()
        out = np.empty_like(bb)
        n = self.l.ref_bb_descramble(standard, framesize, rate, bb.ctypes.data, bb.size, out.ctypes.data)
        assert n == bb.size
        return out

    def bbdeheader(self, standard, framesize, rate, kbch):
        return RefDeheader(self.l, standard, framesize, rate, kbch)

    def ldpc_decode(self, table_name, llr, trials=25, isa="avx2"):
        """Batches of the ISA's SIMD width through ldpc_<isa>::ldpc_dec_decode, in place."""
        code = self.ISA[isa]
        if self._init.get(code) != table_name:
            n = self.l.ref_ldpc_init(table_name.encode(), code)
            assert n > 0, table_name
            self._init[code] = table_name
            self._n = n
        simd = self.l.ref_simd_width(code)
        post = np.array(llr, dtype=np.int8, copy=True).reshape(-1, self._n)
        F = post.shape[0]
        assert F % simd == 0, (F, simd)
        ret = np.zeros(F, dtype=np.int32)
        for f0 in range(0, F, simd):
            blk = np.ascontiguousarray(post[f0:f0 + simd])
            ret[f0:f0 + simd] = self.l.ref_ldpc_decode(code, self._n, blk.ctypes.data, trials)
            post[f0:f0 + simd] = blk
        return post, ret

    def ldpc_decode_mt(self, table_name, llr, trials, threads):
        """In place on a copy; returns (posterior, per-batch ret, seconds)."""
        post = np.array(llr, dtype=np.int8, copy=True)
        F = post.shape[0]
        ret = np.zeros(F // 32, dtype=np.int32)
        secs = self.l.ref_ldpc_decode_mt(table_name.encode(), post.ctypes.data, F, trials, threads, ret.ctypes.data)
        return post, ret, secs

    def bch(self, framesize, t, n):
        return self.l.ref_bch_create(PRIM_POLY[framesize], t, n)

    def bch_raw(self, prim_poly, t, n):
        return self.l.ref_bch_create(prim_poly, t, n)

    def bch_encode(self, h, msg, n):
        k = self.l.ref_bch_k(h)
        msg = np.ascontiguousarray(msg, dtype=np.uint8).reshape(-1, k // 8)
        cw = np.zeros((msg.shape[0], n // 8), dtype=np.uint8)
        for f in range(msg.shape[0]):
            self.l.ref_bch_encode(h, msg[f].ctypes.data, cw[f].ctypes.data)
        return cw

    def bch_decode(self, h, cw, n):
        k = self.l.ref_bch_k(h)
        cw = np.ascontiguousarray(cw, dtype=np.uint8).reshape(-1, n // 8)
        msg = np.zeros((cw.shape[0], k // 8), dtype=np.uint8)
        ret = np.zeros(cw.shape[0], dtype=np.int32)
        for f in range(cw.shape[0]):
            ret[f] = self.l.ref_bch_decode(h, cw[f].ctypes.data, msg[f].ctypes.data)
        return msg, ret

    def bch_decode_mt(self, h, cw, n, threads):
        k = self.l.ref_bch_k(h)
        cw = np.ascontiguousarray(cw, dtype=np.uint8).reshape(-1, n // 8)
        F = cw.shape[0]
        msg = np.zeros((F, k // 8), dtype=np.uint8)
        ret = np.zeros(F, dtype=np.int32)
        secs = self.l.ref_bch_decode_mt(h, cw.ctypes.data, msg.ctypes.data, F, n // 8, k // 8, threads, ret.ctypes.data)
        return msg, ret, secs

    def demap_8psk(self, iq, n0, rows):
        iq = np.ascontiguousarray(iq, dtype=np.float32)
        F = iq.shape[0]
        n_syms = iq[0].size // 2
        out = np.zeros((F, 3 * n_syms), dtype=np.int8)
        n0 = np.broadcast_to(np.asarray(n0, dtype=np.float32), (F,))
        for f in range(F):
            self.l.ref_demap_8psk(iq[f].ctypes.data, n_syms, C.c_float(n0[f]), rows[0], rows[1], rows[2],
                                  out[f].ctypes.data)
        return out
