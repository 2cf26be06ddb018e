"""Transmit-side vector generator for tests and bench.py (synthetic FECFRAMEs).

Not part of the decode path and not in the reference either (its Tx chain is gr-dtv): BCH
encoder, DVB-S2 IRA LDPC encoder, bit interleaver, QPSK/8PSK mapper and AWGN, written with
numpy over whole batches.  Conventions are the ones the reference's receive side assumes:
  * BCH: systematic, first transmitted bit = coefficient of x^(n-1)   (lib/bch.cc:157-173,436-450)
  * QPSK: bit 0 -> +1/sqrt(2), first bit of a pair on I                (lib/qpsk.h:100-149)
  * 8PSK: symbol = m_8psk[4*b0 + 2*b1 + b2]                            (lib/psk.hh:102-109,152-157)
          with b0/b1/b2 taken from the three interleaver columns in the rate-dependent order of
          lib/xfecframe_demapper_cb_impl.cc:48-69
  * AWGN: sigma = sqrt(N0/2) per real dimension, Es = 1                (lib/qa_util.h:33-53)
The LDPC encoder is validated in tests by requiring a clean syndrome from the oracle and from
the compiled reference decoder.
"""
import functools

import numpy as np

from . import (MOD_8PSK, MOD_QPSK, RATE, bch_genpoly, lookup,  # noqa: F401
               table_circulants)

M = 360


def rows_8psk(rate, n_syms):
    """Deinterleaver row offsets (lib/xfecframe_demapper_cb_impl.cc:48-69)."""
    if rate == RATE["C3_5"]:
        return 2 * n_syms, n_syms, 0
    if rate in (RATE["C25_36"], RATE["C13_18"], RATE["C7_15"], RATE["C8_15"], RATE["C26_45"]):
        return n_syms, 0, 2 * n_syms
    return 0, n_syms, 2 * n_syms


@functools.lru_cache(maxsize=None)
def _bch_parity_matrix(framesize, t, n, k):
    """[k, n-k] GF(2) matrix: row i = (x^(n-1-i) mod g), column c = coefficient of x^(n-k-1-c)."""
    g = bch_genpoly(framesize, t)
    deg = len(g) - 1
    assert deg == n - k, (deg, n, k)
    gint = 0
    for d, c in enumerate(g):
        gint |= int(c) << d
    rows = np.zeros((k, deg), dtype=np.uint8)
    r = 1  # x^0
    top = 1 << deg
    for p in range(1, n):  # r = x^p mod g
        r <<= 1
        if r & top:
            r ^= gint
        if p >= deg:
            i = n - 1 - p
            bits = np.frombuffer(r.to_bytes((deg + 7) // 8, "little"), dtype=np.uint8)
            rows[i] = np.unpackbits(bits, bitorder="little")[:deg][::-1]
    return rows.astype(np.float32)


def bch_encode_bits(msg_bits, framesize, t, n):
    """msg_bits [F, k] (0/1) -> codeword bits [F, n]."""
    msg_bits = np.asarray(msg_bits, dtype=np.uint8)
    F, k = msg_bits.shape
    P = _bch_parity_matrix(framesize, t, n, k)
    out = np.empty((F, n), dtype=np.uint8)
    out[:, :k] = msg_bits
    for f0 in range(0, F, 512):
        blk = msg_bits[f0:f0 + 512].astype(np.float32)
        out[f0:f0 + 512, k:] = (blk @ P).astype(np.int64) & 1
    return out


def ldpc_encode_bits(table, info_bits):
    """Systematic IRA encoder (EN 302 307-1 5.3.2): info_bits [F, K] -> codeword bits [F, N]."""
    layer, group, shift = table_circulants(table)
    info_bits = np.asarray(info_bits, dtype=np.uint8)
    F, K = info_bits.shape
    q = int(layer.max()) + 1
    D = info_bits.reshape(F, K // M, M)
    P = np.zeros((F, q, M), dtype=np.uint8)
    for i, g, a in zip(layer, group, shift):
        P[:, i, :] ^= np.roll(D[:, g, :], int(a), axis=-1)
    p = np.ascontiguousarray(P.transpose(0, 2, 1)).reshape(F, q * M)  # parity index q*j + i
    p = np.bitwise_xor.accumulate(p, axis=1)
    return np.concatenate([info_bits, p], axis=1)


def pack_bits(bits):
    return np.packbits(np.asarray(bits, dtype=np.uint8), axis=-1)  # MSB first


def unpack_bits(bytes_, nbits=None):
    b = np.unpackbits(np.asarray(bytes_, dtype=np.uint8), axis=-1)
    return b if nbits is None else b[..., :nbits]


def encode_frames(standard, framesize, rate, F, rng, kldpc_pad_random=True):
    """Random BBFRAMEs -> BCH -> LDPC.  Returns (msg_bytes [F, kbch/8], cw_bits [F, N], info)."""
    info = lookup(standard, framesize, rate)
    msg_bits = rng.integers(0, 2, size=(F, info.kbch), dtype=np.uint8)
    bch_bits = bch_encode_bits(msg_bits, framesize, info.t, info.nbch)
    if info.k_ldpc > info.nbch:  # shortened VL-SNR style codes: fill the rest of the LDPC message
        pad = (rng.integers(0, 2, size=(F, info.k_ldpc - info.nbch), dtype=np.uint8) if kldpc_pad_random
               else np.zeros((F, info.k_ldpc - info.nbch), dtype=np.uint8))
        bch_bits = np.concatenate([bch_bits, pad], axis=1)
    cw = ldpc_encode_bits(info.table, bch_bits[:, :info.k_ldpc])
    return pack_bits(msg_bits), cw, info


_M8PSK = np.array([[np.sqrt(0.5), np.sqrt(0.5)], [1, 0], [-1, 0], [-np.sqrt(0.5), -np.sqrt(0.5)],
                   [0, 1], [np.sqrt(0.5), -np.sqrt(0.5)], [-np.sqrt(0.5), np.sqrt(0.5)], [0, -1]],
                  dtype=np.float32)


def map_symbols(cw_bits, constellation, rate):
    """Codeword bits [F, N] -> XFECFRAME symbols [F, N/bits, 2] float32 (bit interleaved for 8PSK)."""
    cw_bits = np.asarray(cw_bits, dtype=np.uint8)
    F, N = cw_bits.shape
    if constellation == MOD_QPSK:
        s = np.float32(np.sqrt(0.5)) * (1 - 2 * cw_bits.astype(np.float32))
        return s.reshape(F, N // 2, 2)
    if constellation == MOD_8PSK:
        n = N // 3
        r0, r1, r2 = rows_8psk(rate, n)
        idx = (cw_bits[:, r0:r0 + n].astype(np.int64) << 2) | (cw_bits[:, r1:r1 + n].astype(np.int64) << 1) | \
            cw_bits[:, r2:r2 + n]
        return _M8PSK[idx]
    raise ValueError("Unsupported constellation")


def awgn(iq, esn0_db, rng):
    n0 = 10.0 ** (-esn0_db / 10.0)
    noise = rng.standard_normal(size=iq.shape, dtype=np.float32) * np.float32(np.sqrt(n0 / 2))
    return (iq + noise).astype(np.float32), np.float32(n0)


def qpsk_llr(iq, n0):
    """numpy restatement of lib/qpsk.h:208-214 for building LLR inputs without a device."""
    scalar = np.float32(2 * np.sqrt(2.0) / np.float64(np.float32(n0)))
    v = iq.reshape(iq.shape[0], -1).astype(np.float32) * scalar
    return np.clip(np.rint(v), -128, 127).astype(np.int8)


def make_llr_frames(standard, framesize, rate, F, esn0_db, seed):
    """Encoded random frames through QPSK + AWGN, quantised to int8 LLRs [F, N]."""
    rng = np.random.default_rng(seed)
    msg, cw, info = encode_frames(standard, framesize, rate, F, rng)
    iq, n0 = awgn(map_symbols(cw, MOD_QPSK, rate), esn0_db, rng)
    return msg, cw, qpsk_llr(iq, n0), info
