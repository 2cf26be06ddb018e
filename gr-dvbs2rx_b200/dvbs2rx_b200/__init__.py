"""dvbs2rx_b200 -- ctypes binding of libdvbs2_b200.so (C ABI in include/dvbs2_b200.h).

Host side of the B200 DVB-S2 FEC decode path.  The enum ordinals and block-level names mirror
gr-dvbs2rx (include/gnuradio/dvbs2rx/dvb_config.h:15-121, python/dvbs2rx/params.py).  There is
no CPU fallback: compute calls raise Dvbs2Error when the CUDA library or a device is missing.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# DVBS2B200_LIB: another build of the same library (diagnostics builds such as -DDVBS2_PHASE_PROFILE)
LIB_PATH = os.environ.get("DVBS2B200_LIB") or os.path.join(os.path.dirname(_HERE), "libdvbs2_b200.so")

# ---- enums (dvb_config.h) -----------------------------------------------------------------------
STANDARD_DVBS2, STANDARD_DVBT2 = 0, 1
FECFRAME_SHORT, FECFRAME_NORMAL, FECFRAME_MEDIUM = 0, 1, 2
OM_CODEWORD, OM_MESSAGE = 0, 1
INFO_OFF, INFO_ON = 0, 1
_RATES = ("C1_4 C1_3 C2_5 C1_2 C3_5 C2_3 C3_4 C4_5 C5_6 C7_8 C8_9 C9_10 C13_45 C9_20 C90_180 C96_180 "
          "C11_20 C100_180 C104_180 C26_45 C18_30 C28_45 C23_36 C116_180 C20_30 C124_180 C25_36 "
          "C128_180 C13_18 C132_180 C22_30 C135_180 C140_180 C7_9 C154_180 C11_45 C4_15 C14_45 C7_15 "
          "C8_15 C32_45 C2_9_VLSNR C1_5_MEDIUM C11_45_MEDIUM C1_3_MEDIUM C1_5_VLSNR_SF2 "
          "C11_45_VLSNR_SF2 C1_5_VLSNR C4_15_VLSNR C1_3_VLSNR C_OTHER").split()
RATE = {name: i for i, name in enumerate(_RATES)}
globals().update(RATE)
_MODS = ("MOD_QPSK MOD_16QAM MOD_64QAM MOD_256QAM MOD_8PSK MOD_8APSK MOD_16APSK MOD_8_8APSK "
         "MOD_32APSK MOD_4_12_16APSK MOD_4_8_4_16APSK MOD_64APSK MOD_8_16_20_20APSK "
         "MOD_4_12_20_28APSK MOD_128APSK MOD_256APSK MOD_BPSK MOD_BPSK_SF2 MOD_8VSB MOD_OTHER").split()
MOD = {name: i for i, name in enumerate(_MODS)}
globals().update(MOD)

OK, EINVAL, EUNSUPPORTED, ECUDA, ENOMEM = 0, -1, -2, -3, -4
TERM_PER_FRAME = 0


class Dvbs2Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("dvbs2b200 error %d: %s" % (code, msg))
        self.code = code


class CodeInfo(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("table", "n_ldpc", "k_ldpc", "q", "n_circ", "links_total",
                                       "max_cn_deg", "kbch", "nbch", "t", "gf_m")]


_lib = None

# every symbol include/dvbs2_b200.h declares: (restype, argtypes)
_P = C.c_void_p
_SIGS = {
    "dvbs2b200_version": (C.c_int, []),
    "dvbs2b200_last_error": (C.c_char_p, []),
    "dvbs2b200_device_count": (C.c_int, []),
    "dvbs2b200_host_register": (C.c_int, [_P, C.c_size_t]),
    "dvbs2b200_host_unregister": (C.c_int, [_P]),
    "dvbs2b200_num_tables": (C.c_int, []),
    "dvbs2b200_table_name": (C.c_char_p, [C.c_int]),
    "dvbs2b200_lookup": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(CodeInfo)]),
    "dvbs2b200_table_circulants": (C.c_int, [C.c_int, _P, C.c_int]),
    "dvbs2b200_bch_genpoly": (C.c_int, [C.c_int, C.c_int, _P, C.c_int]),
    "dvbs2b200_schedule_stats": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "dvbs2b200_tables_build": (C.c_int, [C.c_int, C.c_int, C.c_int, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "dvbs2b200_code_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.c_int, C.c_int]),
    "dvbs2b200_code_export_tables": (C.c_int, [_P, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "dvbs2b200_code_create_from_tables": (C.c_int, [C.POINTER(_P), C.c_int, _P, C.c_size_t]),
    "dvbs2b200_code_destroy": (None, [_P]),
    "dvbs2b200_code_info_get": (C.c_int, [_P, C.POINTER(CodeInfo)]),
    "dvbs2b200_launch_count": (C.c_uint64, [_P]),
    "dvbs2b200_ldpc_decode": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "dvbs2b200_ldpc_decode_dev": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P]),
    "dvbs2b200_bch_decode": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "dvbs2b200_bch_decode_dev": (C.c_int, [_P, _P, C.c_int, _P, _P, _P]),
    "dvbs2b200_demap": (C.c_int, [_P, C.c_int, _P, C.c_int, _P, _P]),
    "dvbs2b200_demap_dev": (C.c_int, [_P, C.c_int, _P, C.c_int, _P, _P, _P]),
    "dvbs2b200_fec_decode": (C.c_int, [_P, C.c_int, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "dvbs2b200_fec_decode_dev": (C.c_int, [_P, C.c_int, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P]),
    "dvbs2b200_demap_table": (C.c_int, [_P, C.c_int, _P, _P, _P, C.c_int, _P, _P]),
    "dvbs2b200_demap_table_dev": (C.c_int, [_P, C.c_int, _P, _P, _P, C.c_int, _P, _P, _P]),
    "dvbs2b200_mixed_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, _P, _P, _P]),
    "dvbs2b200_mixed_destroy": (None, [_P]),
    "dvbs2b200_mixed_code_info": (C.c_int, [_P, C.c_int, C.POINTER(CodeInfo)]),
    "dvbs2b200_mixed_fec_decode": (C.c_int, [_P, C.c_int, _P, _P, C.c_int, _P, _P, _P]),
    "dvbs2b200_mixed_launch_count": (C.c_uint64, [_P]),
    "dvbs2b200_multi_create": (C.c_int, [C.POINTER(_P), _P, C.c_int, C.c_int, C.c_int, C.c_int]),
    "dvbs2b200_multi_destroy": (None, [_P]),
    "dvbs2b200_multi_device_count": (C.c_int, [_P]),
    "dvbs2b200_multi_code": (_P, [_P, C.c_int]),
    "dvbs2b200_multi_shard": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "dvbs2b200_multi_fec_decode": (C.c_int, [_P, C.c_int, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "dvbs2b200_pl_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int]),
    "dvbs2b200_pl_destroy": (None, [_P]),
    "dvbs2b200_pl_payload_len": (C.c_int, [C.c_int, C.c_int]),
    "dvbs2b200_pl_scrambling_codes": (C.c_int, [C.c_int, _P, C.c_int]),
    "dvbs2b200_pl_descramble_derotate": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "dvbs2b200_pl_descramble_derotate_dev": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "dvbs2b200_mixed_fec_decode_dev": (C.c_int, [_P, C.c_int, _P, _P, C.c_int, _P, _P, _P, _P]),
    "dvbs2b200_mixed_create_from_tables": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, _P, _P]),
    "dvbs2b200_estimate_snr": (C.c_int, [_P, C.c_int, _P, _P, C.c_int, _P]),
    "dvbs2b200_estimate_snr_dev": (C.c_int, [_P, C.c_int, _P, _P, C.c_int, _P, _P]),
    "dvbs2b200_bb_descramble": (C.c_int, [_P, _P, C.c_int, _P]),
    "dvbs2b200_bb_descramble_dev": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "dvbs2b200_bb_ts_capacity": (C.c_size_t, [_P, C.c_int]),
    "dvbs2b200_bb_deheader": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "dvbs2b200_bb_deheader_dev": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, C.c_size_t, _P]),
    "dvbs2b200_bb_produced_dev": (C.c_int, [_P, _P, C.POINTER(C.c_size_t)]),
    "dvbs2b200_bb_reset": (C.c_int, [_P]),
    "dvbs2b200_bb_counters_get": (C.c_int, [_P, _P]),
    "dvbs2b200_fec_decode_ts": (C.c_int, [_P, C.c_int, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P, C.c_size_t,
                                          C.POINTER(C.c_size_t), _P, _P]),
    "dvbs2b200_fec_decode_ts_dev": (C.c_int, [_P, C.c_int, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P, C.c_size_t, _P, _P, _P]),
}
EXPORTED_SYMBOLS = tuple(sorted(_SIGS))


def lib():
    """Load libdvbs2_b200.so (built by __graft_entry__.build()).  Fails loudly if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Dvbs2Error(ECUDA, "%s not built: run `python __graft_entry__.py build`" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def _check(rc):
    if rc != 0:
        raise Dvbs2Error(rc, lib().dvbs2b200_last_error().decode())


def host_register(array):
    """Page-lock a numpy array the caller keeps across calls (dvbs2b200_host_register)."""
    _check(lib().dvbs2b200_host_register(array.ctypes.data, array.nbytes))


def host_unregister(array):
    _check(lib().dvbs2b200_host_unregister(array.ctypes.data))


def device_count():
    return lib().dvbs2b200_device_count()


def lookup(standard, framesize, rate):
    info = CodeInfo()
    _check(lib().dvbs2b200_lookup(standard, framesize, rate, C.byref(info)))
    return info


def table_circulants(table):
    """(layer, group, shift) int arrays of one LDPC table."""
    n = lib().dvbs2b200_table_circulants(table, None, 0)
    if n < 0:
        _check(n)
    w = np.zeros(n, dtype=np.uint32)
    lib().dvbs2b200_table_circulants(table, w.ctypes.data, n)
    return (w >> 17).astype(np.int64), ((w >> 9) & 0xff).astype(np.int64), (w & 0x1ff).astype(np.int64)


def bch_genpoly(framesize, t):
    g = np.zeros(256, dtype=np.uint8)
    deg = lib().dvbs2b200_bch_genpoly(framesize, t, g.ctypes.data, 256)
    if deg < 0:
        _check(deg)
    return g[:deg + 1].copy()


def schedule_stats(table):
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    _check(lib().dvbs2b200_schedule_stats(table, C.byref(a), C.byref(b), C.byref(c)))
    return dict(steps_per_iter=a.value, max_depth=b.value, conflict_layers=c.value)


def build_tables(standard, framesize, rate):
    """Packed code tables (uint8 array) built on the host; no device needed."""
    size = C.c_size_t()
    _check(lib().dvbs2b200_tables_build(standard, framesize, rate, None, 0, C.byref(size)))
    buf = np.zeros(size.value, dtype=np.uint8)
    _check(lib().dvbs2b200_tables_build(standard, framesize, rate, buf.ctypes.data, buf.size, C.byref(size)))
    return buf


def bits_per_symbol(constellation):
    return {MOD_QPSK: 2, MOD_8PSK: 3}.get(constellation, 0)  # noqa: F821


def _ptr(a):
    return None if a is None else a.ctypes.data


def _np(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


class Code:
    """One (standard, framesize, rate) bound to one CUDA device -- the state a gr-dvbs2rx block
    instance holds (ldpc_decoder_bb_impl.cc:79-368, bch_decoder_bb_impl.cc:43-71)."""

    def __init__(self, standard=STANDARD_DVBS2, framesize=FECFRAME_NORMAL, rate=None, device=0, tables=None):
        self._h = _P()
        if tables is not None:
            buf = np.ascontiguousarray(tables, dtype=np.uint8)
            _check(lib().dvbs2b200_code_create_from_tables(C.byref(self._h), device, buf.ctypes.data, buf.size))
        else:
            _check(lib().dvbs2b200_code_create(C.byref(self._h), device, standard, framesize, rate))
        self.info = CodeInfo()
        _check(lib().dvbs2b200_code_info_get(self._h, C.byref(self.info)))
        self.device = device
        self.N = self.info.n_ldpc
        self.kbch, self.nbch = self.info.kbch, self.info.nbch

    def close(self):
        if self._h:
            lib().dvbs2b200_code_destroy(self._h)
            self._h = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self):
        return int(lib().dvbs2b200_launch_count(self._h))

    def export_tables(self):
        size = C.c_size_t()
        _check(lib().dvbs2b200_code_export_tables(self._h, None, 0, C.byref(size)))
        buf = np.zeros(size.value, dtype=np.uint8)
        _check(lib().dvbs2b200_code_export_tables(self._h, buf.ctypes.data, buf.size, C.byref(size)))
        return buf

    def out_bytes(self, output_mode):
        return (self.nbch if output_mode else self.N) // 8

    # ---- host-buffer entry points (numpy) ------------------------------------------------------
    def ldpc_decode(self, llr, max_trials=25, term_group=TERM_PER_FRAME, output_mode=OM_MESSAGE,
                    want_post=False, want_trials=True):
        llr = _np(llr, np.int8).reshape(-1, self.N)
        F = llr.shape[0]
        hard = np.empty((F, self.out_bytes(output_mode)), dtype=np.uint8)
        post = np.empty((F, self.N), dtype=np.int8) if want_post else None
        trials = np.empty(F, dtype=np.int32) if want_trials else None
        _check(lib().dvbs2b200_ldpc_decode(self._h, llr.ctypes.data, F, max_trials, term_group, output_mode,
                                           hard.ctypes.data, _ptr(post), _ptr(trials)))
        return hard, post, trials

    def bch_decode(self, cw):
        cw = _np(cw, np.uint8).reshape(-1, self.nbch // 8)
        F = cw.shape[0]
        msg = np.empty((F, self.kbch // 8), dtype=np.uint8)
        corr = np.empty(F, dtype=np.int32)
        _check(lib().dvbs2b200_bch_decode(self._h, cw.ctypes.data, F, msg.ctypes.data, corr.ctypes.data))
        return msg, corr

    def demap(self, constellation, iq, n0):
        bits = bits_per_symbol(constellation)
        if not bits:
            raise Dvbs2Error(EUNSUPPORTED, "Unsupported constellation")
        iq = _np(iq, np.float32).reshape(-1, self.N // bits, 2)
        F = iq.shape[0]
        n0 = _np(np.broadcast_to(np.asarray(n0, dtype=np.float32), (F,)), np.float32)
        llr = np.empty((F, self.N), dtype=np.int8)
        _check(lib().dvbs2b200_demap(self._h, constellation, iq.ctypes.data, F, n0.ctypes.data, llr.ctypes.data))
        return llr

    def fec_decode(self, llr=None, iq=None, n0=None, constellation=0, max_trials=25, term_group=TERM_PER_FRAME):
        if iq is not None:
            bits = bits_per_symbol(constellation)
            iq = _np(iq, np.float32).reshape(-1, self.N // bits, 2)
            F = iq.shape[0]
            n0 = _np(np.broadcast_to(np.asarray(n0, dtype=np.float32), (F,)), np.float32)
        else:
            llr = _np(llr, np.int8).reshape(-1, self.N)
            F = llr.shape[0]
        msg = np.empty((F, self.kbch // 8), dtype=np.uint8)
        trials = np.empty(F, dtype=np.int32)
        corr = np.empty(F, dtype=np.int32)
        _check(lib().dvbs2b200_fec_decode(self._h, constellation, _ptr(iq), _ptr(n0), _ptr(llr), F, max_trials,
                                          term_group, msg.ctypes.data, trials.ctypes.data, corr.ctypes.data))
        return msg, trials, corr

    def demap_table(self, points, row_offsets, iq, n0):
        """Max-log soft demap against a constellation table [2^bits, 2]; bit k of symbol j -> llr[row_offsets[k] + j]."""
        points = _np(points, np.float32).reshape(-1, 2)
        bits = int(points.shape[0]).bit_length() - 1
        rows = _np(row_offsets, np.int32)
        iq = _np(iq, np.float32).reshape(-1, self.N // bits, 2)
        F = iq.shape[0]
        n0 = _np(np.broadcast_to(np.asarray(n0, dtype=np.float32), (F,)), np.float32)
        llr = np.empty((F, self.N), dtype=np.int8)
        _check(lib().dvbs2b200_demap_table(self._h, bits, points.ctypes.data, rows.ctypes.data, iq.ctypes.data, F,
                                           n0.ctypes.data, llr.ctypes.data))
        return llr

    def demap_table_dev(self, points, row_offsets, d_iq, frames, d_n0, d_llr, stream):
        points = _np(points, np.float32).reshape(-1, 2)
        bits = int(points.shape[0]).bit_length() - 1
        rows = _np(row_offsets, np.int32)
        _check(lib().dvbs2b200_demap_table_dev(self._h, bits, points.ctypes.data, rows.ctypes.data, d_iq, frames, d_n0, d_llr, stream))

    def estimate_snr(self, constellation, iq, llr_post=None):
        """Linear Es/N0 per frame: from sliced symbols, or from posterior LLR signs when given."""
        bits = bits_per_symbol(constellation)
        if not bits:
            raise Dvbs2Error(EUNSUPPORTED, "Unsupported constellation")
        iq = _np(iq, np.float32).reshape(-1, self.N // bits, 2)
        F = iq.shape[0]
        if llr_post is not None:
            llr_post = _np(llr_post, np.int8).reshape(F, self.N)
        out = np.empty(F, dtype=np.float32)
        _check(lib().dvbs2b200_estimate_snr(self._h, constellation, iq.ctypes.data, _ptr(llr_post), F, out.ctypes.data))
        return out

    def estimate_snr_dev(self, constellation, d_iq, d_llr_post, frames, d_snr, stream):
        _check(lib().dvbs2b200_estimate_snr_dev(self._h, constellation, d_iq, d_llr_post, frames, d_snr, stream))

    # ---- BB layer: BBFRAMEs -> TS packets (bbdescrambler_bb / bbdeheader_bb) ---------------------
    def bb_descramble(self, bbframes):
        bb = _np(bbframes, np.uint8).reshape(-1, self.kbch // 8)
        out = np.empty_like(bb)
        _check(lib().dvbs2b200_bb_descramble(self._h, bb.ctypes.data, bb.shape[0], out.ctypes.data))
        return out

    def bb_ts_capacity(self, frames):
        return int(lib().dvbs2b200_bb_ts_capacity(self._h, frames))

    def bb_deheader(self, bbframes, scrambled=False):
        """Stateful, like the reference block: returns the TS bytes this call produced."""
        bb = _np(bbframes, np.uint8).reshape(-1, self.kbch // 8)
        F = bb.shape[0]
        ts = np.empty(max(self.bb_ts_capacity(F), 188), dtype=np.uint8)
        n = C.c_size_t()
        _check(lib().dvbs2b200_bb_deheader(self._h, bb.ctypes.data, F, 1 if scrambled else 0, ts.ctypes.data, ts.size, C.byref(n)))
        return ts[:n.value].copy()

    def bb_reset(self):
        _check(lib().dvbs2b200_bb_reset(self._h))

    def bb_counters(self):
        c = (C.c_uint64 * 5)()
        _check(lib().dvbs2b200_bb_counters_get(self._h, c))
        return dict(zip(("packets", "errors", "bbframes", "dropped", "gaps"), [int(v) for v in c]))

    def fec_decode_ts(self, llr=None, iq=None, n0=None, constellation=0, max_trials=25, term_group=TERM_PER_FRAME):
        """Soft input -> TS packets with every intermediate in device memory."""
        if iq is not None:
            bits = bits_per_symbol(constellation)
            iq = _np(iq, np.float32).reshape(-1, self.N // bits, 2)
            F = iq.shape[0]
            n0 = _np(np.broadcast_to(np.asarray(n0, dtype=np.float32), (F,)), np.float32)
        else:
            llr = _np(llr, np.int8).reshape(-1, self.N)
            F = llr.shape[0]
        ts = np.empty(max(self.bb_ts_capacity(F), 188), dtype=np.uint8)
        n = C.c_size_t()
        trials = np.empty(F, dtype=np.int32)
        corr = np.empty(F, dtype=np.int32)
        _check(lib().dvbs2b200_fec_decode_ts(self._h, constellation, _ptr(iq), _ptr(n0), _ptr(llr), F, max_trials, term_group,
                                             ts.ctypes.data, ts.size, C.byref(n), trials.ctypes.data, corr.ctypes.data))
        return ts[:n.value].copy(), trials, corr

    def fec_decode_ts_ptr(self, constellation, iq_ptr, n0_ptr, llr_ptr, frames, max_trials, term_group, ts_ptr, ts_cap,
                          trials_ptr, corr_ptr):
        n = C.c_size_t()
        _check(lib().dvbs2b200_fec_decode_ts(self._h, constellation, iq_ptr, n0_ptr, llr_ptr, frames, max_trials, term_group,
                                             ts_ptr, ts_cap, C.byref(n), trials_ptr, corr_ptr))
        return n.value

    def bb_deheader_dev(self, d_bb, frames, scrambled, d_ts, ts_cap, stream):
        _check(lib().dvbs2b200_bb_deheader_dev(self._h, d_bb, frames, scrambled, d_ts, ts_cap, stream))

    def bb_produced_dev(self, stream):
        n = C.c_size_t()
        _check(lib().dvbs2b200_bb_produced_dev(self._h, stream, C.byref(n)))
        return n.value

    def fec_decode_ts_dev(self, constellation, d_iq, d_n0, d_llr, frames, max_trials, term_group, d_ts, ts_cap, d_trials,
                          d_corr, stream):
        _check(lib().dvbs2b200_fec_decode_ts_dev(self._h, constellation, d_iq, d_n0, d_llr, frames, max_trials, term_group,
                                                 d_ts, ts_cap, d_trials, d_corr, stream))

    # ---- raw-pointer entry points (device or pinned host addresses as ints) ----------------------
    def ldpc_decode_ptr(self, llr_ptr, frames, max_trials, term_group, output_mode, hard_ptr, post_ptr, trials_ptr):
        _check(lib().dvbs2b200_ldpc_decode(self._h, llr_ptr, frames, max_trials, term_group, output_mode,
                                           hard_ptr, post_ptr, trials_ptr))

    def fec_decode_ptr(self, constellation, iq_ptr, n0_ptr, llr_ptr, frames, max_trials, term_group, msg_ptr,
                       trials_ptr, corr_ptr):
        _check(lib().dvbs2b200_fec_decode(self._h, constellation, iq_ptr, n0_ptr, llr_ptr, frames, max_trials,
                                          term_group, msg_ptr, trials_ptr, corr_ptr))

    def ldpc_decode_dev(self, d_llr, frames, max_trials, term_group, output_mode, d_hard, d_post, d_trials, stream):
        _check(lib().dvbs2b200_ldpc_decode_dev(self._h, d_llr, frames, max_trials, term_group, output_mode,
                                               d_hard, d_post, d_trials, stream))

    def bch_decode_dev(self, d_cw, frames, d_msg, d_corr, stream):
        _check(lib().dvbs2b200_bch_decode_dev(self._h, d_cw, frames, d_msg, d_corr, stream))

    def demap_dev(self, constellation, d_iq, frames, d_n0, d_llr, stream):
        _check(lib().dvbs2b200_demap_dev(self._h, constellation, d_iq, frames, d_n0, d_llr, stream))

    def fec_decode_dev(self, constellation, d_iq, d_n0, d_llr, frames, max_trials, term_group, d_msg, d_trials,
                       d_corr, stream):
        _check(lib().dvbs2b200_fec_decode_dev(self._h, constellation, d_iq, d_n0, d_llr, frames, max_trials,
                                              term_group, d_msg, d_trials, d_corr, stream))


class MixedCodes:
    """A set of MODCODs on one device for mixed (VCM/ACM) batches: frames carry a per-frame code id."""

    def __init__(self, modcods=None, device=0, tables=None):
        """modcods: [(standard, framesize, rate), ...], or tables: [packed blob (uint8 array), ...]"""
        self._h = _P()
        if tables is not None:
            n = len(tables)
            bufs = [np.ascontiguousarray(t, dtype=np.uint8) for t in tables]
            ptrs = (_P * n)(*[b.ctypes.data for b in bufs])
            sizes = (C.c_size_t * n)(*[b.size for b in bufs])
            _check(lib().dvbs2b200_mixed_create_from_tables(C.byref(self._h), device, n, ptrs, sizes))
        else:
            n = len(modcods)
            std = (C.c_int * n)(*[m[0] for m in modcods])
            fs = (C.c_int * n)(*[m[1] for m in modcods])
            rt = (C.c_int * n)(*[m[2] for m in modcods])
            _check(lib().dvbs2b200_mixed_create(C.byref(self._h), device, n, std, fs, rt))
        self.infos = []
        for c in range(n):
            info = CodeInfo()
            _check(lib().dvbs2b200_mixed_code_info(self._h, c, C.byref(info)))
            self.infos.append(info)

    def close(self):
        if self._h:
            lib().dvbs2b200_mixed_destroy(self._h)
            self._h = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self):
        return int(lib().dvbs2b200_mixed_launch_count(self._h))

    def sizes(self, code_id):
        """(input bytes, output bytes) per frame of a batch."""
        n = np.array([i.n_ldpc for i in self.infos], dtype=np.int64)[code_id]
        k = np.array([i.kbch // 8 for i in self.infos], dtype=np.int64)[code_id]
        return n, k

    def fec_decode_dev(self, code_id, d_llr, max_trials, d_msg, d_trials, d_corr, stream):
        code_id = _np(code_id, np.uint8)
        _check(lib().dvbs2b200_mixed_fec_decode_dev(self._h, code_id.size, code_id.ctypes.data, d_llr, max_trials, d_msg, d_trials,
                                                    d_corr, stream))

    def fec_decode_ptr(self, code_id, llr_ptr, max_trials, msg_ptr, trials_ptr, corr_ptr):
        code_id = _np(code_id, np.uint8)
        _check(lib().dvbs2b200_mixed_fec_decode(self._h, code_id.size, code_id.ctypes.data, llr_ptr, max_trials, msg_ptr, trials_ptr,
                                                corr_ptr))

    def fec_decode(self, code_id, llr, max_trials=25):
        """code_id [F] uint8; llr: the frames' int8 LLRs back to back.  Returns (msg bytes back to back,
        trials_left [F], corrections [F])."""
        code_id = _np(code_id, np.uint8)
        F = code_id.size
        n, k = self.sizes(code_id)
        llr = _np(llr, np.int8).ravel()
        assert llr.size == int(n.sum())
        msg = np.empty(int(k.sum()), dtype=np.uint8)
        trials = np.empty(F, dtype=np.int32)
        corr = np.empty(F, dtype=np.int32)
        _check(lib().dvbs2b200_mixed_fec_decode(self._h, F, code_id.ctypes.data, llr.ctypes.data, max_trials, msg.ctypes.data,
                                                trials.ctypes.data, corr.ctypes.data))
        return msg, trials, corr


class MultiCode:
    """One code on several devices of this process (dvbs2b200_multi_*): the batch is cut into contiguous frame ranges."""

    def __init__(self, devices, standard=STANDARD_DVBS2, framesize=FECFRAME_NORMAL, rate=None):
        self._h = _P()
        devs = (C.c_int * len(devices))(*devices)
        _check(lib().dvbs2b200_multi_create(C.byref(self._h), devs, len(devices), standard, framesize, rate))
        self.info = lookup(standard, framesize, rate)
        self.N, self.kbch = self.info.n_ldpc, self.info.kbch

    def close(self):
        if self._h:
            lib().dvbs2b200_multi_destroy(self._h)
            self._h = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def shard(self, frames, index):
        a, b = C.c_int(), C.c_int()
        _check(lib().dvbs2b200_multi_shard(self._h, frames, index, C.byref(a), C.byref(b)))
        return a.value, b.value

    def fec_decode(self, llr, max_trials=25, term_group=TERM_PER_FRAME):
        llr = _np(llr, np.int8).reshape(-1, self.N)
        F = llr.shape[0]
        msg = np.empty((F, self.kbch // 8), dtype=np.uint8)
        trials = np.empty(F, dtype=np.int32)
        corr = np.empty(F, dtype=np.int32)
        _check(lib().dvbs2b200_multi_fec_decode(self._h, 0, None, None, llr.ctypes.data, F, max_trials, term_group, msg.ctypes.data,
                                                trials.ctypes.data, corr.ctypes.data))
        return msg, trials, corr


PL_FRAME_DTYPE = np.dtype([("plheader_phase", np.float32), ("fine_foffset", np.float32), ("coarse_corrected", np.int32),
                           ("reserved", np.int32), ("pilot_phase", np.float32, (22,))])


def pl_scrambling_codes(gold_code, n):
    rn = np.zeros(n, dtype=np.uint8)
    _check(lib().dvbs2b200_pl_scrambling_codes(gold_code, rn.ctypes.data, n))
    return rn


class PlDescrambler:
    """PL descrambler + pilot-segment de-rotation of one Gold code on one device (dvbs2b200_pl_*)."""

    def __init__(self, gold_code=0, device=0):
        self._h = _P()
        _check(lib().dvbs2b200_pl_create(C.byref(self._h), device, gold_code))

    def close(self):
        if self._h:
            lib().dvbs2b200_pl_destroy(self._h)
            self._h = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def payload_len(n_slots, has_pilots):
        return lib().dvbs2b200_pl_payload_len(n_slots, 1 if has_pilots else 0)

    def process(self, payload, n_slots, has_pilots, info):
        """payload [F, payload_len, 2] float32, info: array of PL_FRAME_DTYPE [F] -> XFECFRAMEs [F, n_slots * 90, 2]."""
        payload = _np(payload, np.float32).reshape(-1, self.payload_len(n_slots, has_pilots), 2)
        F = payload.shape[0]
        info = np.ascontiguousarray(info, dtype=PL_FRAME_DTYPE)
        assert info.size == F
        out = np.empty((F, n_slots * 90, 2), dtype=np.float32)
        _check(lib().dvbs2b200_pl_descramble_derotate(self._h, payload.ctypes.data, F, n_slots, 1 if has_pilots else 0,
                                                      info.ctypes.data, out.ctypes.data))
        return out

    def process_dev(self, d_payload, frames, n_slots, has_pilots, d_info, d_out, stream):
        _check(lib().dvbs2b200_pl_descramble_derotate_dev(self._h, d_payload, frames, n_slots, 1 if has_pilots else 0, d_info, d_out, stream))
