"""Synthetic BBFRAME streams carrying MPEG-TS packets (test / bench input for the BB layer).

Builds what a DVB-S2 modulator's baseband framer emits (EN 302 307-1 clause 5.1): user packets of 188
bytes whose sync byte is replaced by the CRC-8 of the preceding packet, sliced into DATAFIELDs behind a
10-byte BBHEADER (MATYPE, UPL, DFL, SYNC, SYNCD, CRC-8), optionally zero padded and scrambled.  The same
cases as the reference's python/dvbs2rx/qa_bbdeheader_bb.py are expressible (its generator is not used).
"""
import numpy as np

TS_LEN = 188
HDR_LEN = 10
_CRC_POLY = 0x1D5  # x^8 + x^7 + x^6 + x^4 + x^2 + 1


def _crc_table():
    t = np.zeros(256, dtype=np.uint8)
    for i in range(256):
        r = i << 8
        for b in range(15, 7, -1):
            if r & (1 << b):
                r ^= _CRC_POLY << (b - 8)
        t[i] = r & 0xFF
    return t


_T = _crc_table()


def crc8(data):
    c = 0
    for b in bytes(data):
        c = int(_T[c ^ b])
    return c


def prbs(nbytes):
    """BB scrambling sequence 1 + x^14 + x^15, initial register 100101010000000, MSB first."""
    reg = [1, 0, 0, 1, 0, 1, 0, 1, 0, 0, 0, 0, 0, 0, 0]
    bits = np.zeros(nbytes * 8, dtype=np.uint8)
    for i in range(nbytes * 8):
        fb = reg[13] ^ reg[14]
        bits[i] = fb
        reg = [fb] + reg[:14]
    return np.packbits(bits)


def ts_packets(n, rng):
    """n random TS packets: 0x47, three zero header bytes, random payload."""
    up = rng.integers(0, 256, size=(n, TS_LEN), dtype=np.uint8)
    up[:, 0] = 0x47
    up[:, 1:4] = 0
    return up


def crc_encode(up):
    """Sync byte of packet i replaced by the CRC-8 of bytes 1..187 of packet i-1 (the first keeps 0x47)."""
    out = up.copy()
    for i in range(1, up.shape[0]):
        out[i, 0] = crc8(up[i - 1, 1:])
    return out


def bbheader(dfl_bits, syncd_bits, upl_bits=TS_LEN * 8, matype1=0xF2, matype2=0, sync=0x47):
    h = bytes([matype1, matype2, upl_bits >> 8, upl_bits & 0xFF, dfl_bits >> 8, dfl_bits & 0xFF, sync,
               syncd_bits >> 8, syncd_bits & 0xFF])
    return np.frombuffer(h + bytes([crc8(h)]), dtype=np.uint8)


def bbframe_stream(kbch, n_frames, up, dfl_bytes=None, syncd0_bits=0, first_packet=0):
    """BBFRAMEs [n_frames, kbch/8] (not scrambled) filled from the CRC-encoded packets `up`.
    dfl_bytes: DATAFIELD length (default: the maximum, kbch/8 - 10); shorter fields are zero padded."""
    kb = kbch // 8
    max_df = kb - HDR_LEN
    df = max_df if dfl_bytes is None else dfl_bytes
    stream = crc_encode(up).ravel()[first_packet * TS_LEN:]
    assert stream.size >= n_frames * df
    out = np.zeros((n_frames, kb), dtype=np.uint8)
    off = 0
    syncd = syncd0_bits
    for f in range(n_frames):
        out[f, :HDR_LEN] = bbheader(df * 8, syncd)
        out[f, HDR_LEN:HDR_LEN + df] = stream[off:off + df]
        off += df
        syncd = ((TS_LEN - off % TS_LEN) % TS_LEN) * 8
    return out


def scramble(bbframes):
    bb = np.ascontiguousarray(bbframes, dtype=np.uint8)
    return bb ^ prbs(bb.shape[1])[None, :]
