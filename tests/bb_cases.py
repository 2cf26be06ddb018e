"""BBFRAME streams for the BB-layer tests (oracle vs compiled reference on the CPU, product vs oracle on
the GPU): clean streams, every fault the reference's qa_bbdeheader_bb.py injects, and a few it does not."""
import numpy as np

from dvbs2rx_b200 import bbframes as bbf


def _stream(kbch, n_frames, rng, dfl_bytes=None, syncd0_bits=0):
    kb = kbch // 8
    df = (kb - 10) if dfl_bytes is None else dfl_bytes
    n_up = (n_frames * df + 187) // 188 + 1
    up = bbf.ts_packets(n_up, rng)
    return up, bbf.bbframe_stream(kbch, n_frames, up, dfl_bytes=dfl_bytes, syncd0_bits=syncd0_bits)


def make_cases(seed=11, ub=False):
    """[(name, kbch, [calls: uint8 [frames, kbch/8]])].  ub=True adds the one case where the reference has
    undefined behaviour (re-synchronisation with syncd == dfl), meaningful for oracle vs product only."""
    rng = np.random.default_rng(seed)
    cases = []
    kbch = 16008  # QPSK 1/4 normal, the code of the reference's QA
    kb = kbch // 8
    # 1. consecutive frames, one call (qa: test_successful_deframing)
    up, bb = _stream(kbch, 10, rng)
    cases.append(("clean", kbch, [bb]))
    # 2. the same stream cut into uneven calls: the partial packet and the sync state cross call boundaries
    cases.append(("clean_split_calls", kbch, [bb[:1], bb[1:4], bb[4:5], bb[5:]]))
    # 3. BBHEADER CRC error on the first frame and in the middle (qa: test_bbheader_crc_error)
    b = bb.copy()
    b[0, 9] ^= 0xFF
    b[6, 3] ^= 0x10
    cases.append(("header_crc_errors", kbch, [b[:5], b[5:]]))
    # 4. CRC-consistent but invalid headers (qa: test_undetected_dfl_corruption / syncd_corruption)
    b = bb.copy()
    b[1, :10] = bbf.bbheader(kbch - 80 + 8, 0)          # dfl > kbch - 80
    b[3, :10] = bbf.bbheader(kbch - 80 - 4, 0)          # dfl not a multiple of 8
    b[5, :10] = bbf.bbheader(800, 808)                  # syncd > dfl
    b[7, :10] = bbf.bbheader(kbch - 80, 0, upl_bits=1496)  # upl != 188 bytes
    cases.append(("invalid_headers", kbch, [b]))
    # 5. byte-misaligned SYNCD on the first frame (qa: test_non_byte_aligned_syncd)
    up5, b5 = _stream(kbch, 10, rng, syncd0_bits=5)
    cases.append(("misaligned_syncd", kbch, [b5]))
    # 6. a frame missing in the middle (qa: test_non_consecutive_bbframes)
    cases.append(("dropped_frame", kbch, [np.concatenate([bb[:5], bb[6:]])]))
    # 7. zero-padded DATAFIELDs holding whole packets (qa: test_padded_dfl)
    df = ((kb - 10) // 188) * 188
    up7, b7 = _stream(kbch, 4, rng, dfl_bytes=df)
    cases.append(("padded_dfl", kbch, [b7]))
    # 8. payload corruption: CRC-8 of two packets fails -> transport error indicator, error count
    b = bb.copy()
    b[2, 500] ^= 0x01
    b[8, 1500] ^= 0x80
    cases.append(("payload_errors", kbch, [b]))
    # 9. short DATAFIELDs (< 188 bytes): the packet loop is never entered and the partial buffer is overwritten
    up9, b9 = _stream(kbch, 12, rng, dfl_bytes=100)
    cases.append(("short_datafields", kbch, [b9[:7], b9[7:]]))
    # 10. everything at once on the smallest BBFRAME (short 1/4: kbch = 3072) and the largest (9/10 normal)
    for kbch2, n in ((3072, 40), (58192, 9)):
        up10, b10 = _stream(kbch2, n, rng)
        b10 = b10.copy()
        b10[3, 9] ^= 0x55
        b10[n // 2, 200] ^= 0x04
        keep = [i for i in range(n) if i != n - 3]
        b10 = b10[keep]
        cases.append(("mixed_kbch%d" % kbch2, kbch2, [b10[:2], b10[2:n // 2], b10[n // 2:]]))
    # 11. garbage: random bytes (BCH failures look like this) between good frames
    b = bb.copy()
    b[4] = rng.integers(0, 256, size=kb, dtype=np.uint8)
    cases.append(("garbage_frame", kbch, [b]))
    if ub:
        b = bb.copy()
        b[0, :10] = bbf.bbheader(1600, 1600)  # first frame re-synchronises with syncd == dfl
        cases.append(("resync_syncd_eq_dfl", kbch, [b]))
    return cases
