"""16APSK / 32APSK constellation tables and bit-interleaver row offsets for dvbs2b200_demap_table.

The reference (gr-dvbs2rx) has no APSK demapper, so nothing here is checked against it.  The tables follow
EN 302 307-1 clause 5.4.3 (16APSK: 4+12 points, ring ratio gamma = R2/R1 by code rate, table 9) and clause
5.4.4 (32APSK: 4+12+16 points, gamma1 = R2/R1, gamma2 = R3/R1, table 10), bit labels as in figures 11 and 12,
energy normalised to 1.  PARITY UNPINNED: they were written down without access to the standard's text in this
build environment and no reference implementation exists to check them against.  What IS checked
(tests/test_apsk_tables.py): ring sizes 4 + 12 (+ 16), equal phase spacing, unit energy, the ring ratios of every code
rate, Gray labelling along the 4- and 12-point rings (the 16-point ring of 32APSK is quasi-Gray: one or two bits
between neighbours), one-bit label symmetry under the I / Q mirror images, and the C++ twin of the tables
(host/dvbs2rx_b200_blocks.cc:apsk_points).  Verify the label tables against the standard before relying on them
for on-air signals.  The demapper itself is table driven and independent of them (any labelled constellation of up
to 32 points).
"""
import numpy as np

# ring ratios by code rate name (EN 302 307-1 tables 9 and 10)
GAMMA_16APSK = {"C2_3": 3.15, "C3_4": 2.85, "C4_5": 2.75, "C5_6": 2.70, "C8_9": 2.60, "C9_10": 2.57}
GAMMA_32APSK = {"C3_4": (2.84, 5.27), "C4_5": (2.72, 4.87), "C5_6": (2.64, 4.64), "C8_9": (2.54, 4.33), "C9_10": (2.53, 4.30)}


def _pol(r, phase):
    return [r * np.cos(phase), r * np.sin(phase)]


def points_16apsk(gamma):
    """[16, 2] float32, index = 4-bit label (first bit = MSB); unit average energy."""
    r1 = np.sqrt(4.0 / (1.0 + 3.0 * gamma * gamma))
    r2 = gamma * r1
    p = np.pi
    outer = [p / 4, -p / 4, 3 * p / 4, -3 * p / 4, p / 12, -p / 12, 11 * p / 12, -11 * p / 12,
             5 * p / 12, -5 * p / 12, 7 * p / 12, -7 * p / 12]
    inner = [p / 4, -p / 4, 3 * p / 4, -3 * p / 4]
    return np.array([_pol(r2, a) for a in outer] + [_pol(r1, a) for a in inner], dtype=np.float32)


def points_32apsk(gamma1, gamma2):
    """[32, 2] float32, index = 5-bit label (first bit = MSB); unit average energy."""
    r1 = np.sqrt(8.0 / (1.0 + 3.0 * gamma1 * gamma1 + 4.0 * gamma2 * gamma2))
    r2, r3 = gamma1 * r1, gamma2 * r1
    p = np.pi
    tab = [(r2, p / 4), (r2, 5 * p / 12), (r2, -p / 4), (r2, -5 * p / 12), (r2, 3 * p / 4), (r2, 7 * p / 12),
           (r2, -3 * p / 4), (r2, -7 * p / 12), (r3, p / 8), (r3, 3 * p / 8), (r3, -p / 4), (r3, -p / 2),
           (r3, 3 * p / 4), (r3, p / 2), (r3, -7 * p / 8), (r3, -5 * p / 8), (r2, p / 12), (r1, p / 4),
           (r2, -p / 12), (r1, -p / 4), (r2, 11 * p / 12), (r1, 3 * p / 4), (r2, -11 * p / 12), (r1, -3 * p / 4),
           (r3, 0.0), (r3, p / 4), (r3, -p / 8), (r3, -3 * p / 8), (r3, 7 * p / 8), (r3, 5 * p / 8), (r3, p), (r3, -3 * p / 4)]
    return np.array([_pol(r, a) for r, a in tab], dtype=np.float32)


def row_offsets(n_ldpc, bits, column_order=None):
    """Bit k of symbol j comes from interleaver column column_order[k] (default 0, 1, ..): codeword index
    column * rows + j (EN 302 307-1 clause 5.3.3: written column-wise, read row-wise, MSB first)."""
    rows = n_ldpc // bits
    order = list(range(bits)) if column_order is None else list(column_order)
    return np.array([c * rows for c in order], dtype=np.int32)


def map_bits(cw_bits, points, offsets):
    """Codeword bits [F, N] -> symbols [F, N/bits, 2] with the interleaver of `offsets` (test / bench input)."""
    cw_bits = np.asarray(cw_bits, dtype=np.uint8)
    F, N = cw_bits.shape
    bits = int(points.shape[0]).bit_length() - 1
    n = N // bits
    idx = np.zeros((F, n), dtype=np.int64)
    for k in range(bits):
        idx = (idx << 1) | cw_bits[:, offsets[k]:offsets[k] + n]
    return points[idx]


def maxlog_llr(iq, points, offsets, n0):
    """float64 model of dvbs2b200_demap_table (the checker of the GPU kernel, tolerance 1 LSB)."""
    iq = np.asarray(iq, dtype=np.float32)
    F, n, _ = iq.shape
    bits = int(points.shape[0]).bit_length() - 1
    y = iq.astype(np.float64)
    pts = points.astype(np.float64)
    d = (y[:, :, None, 0] - pts[None, None, :, 0]) ** 2 + (y[:, :, None, 1] - pts[None, None, :, 1]) ** 2
    out = np.zeros((F, n * bits), dtype=np.float64)
    labels = np.arange(points.shape[0])
    for k in range(bits):
        one = ((labels >> (bits - 1 - k)) & 1).astype(bool)
        v = (d[:, :, one].min(axis=2) - d[:, :, ~one].min(axis=2)) / np.float64(n0)
        out[:, offsets[k]:offsets[k] + n] = v
    return out
