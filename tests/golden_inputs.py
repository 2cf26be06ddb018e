"""Deterministic, platform-independent test inputs for the golden fixtures.

Everything here is integer arithmetic on numpy uint64 (an LCG), so the same seed produces the same
bytes on every machine and numpy version.  tools/gen_golden.py feeds these inputs to the compiled
reference (oracle/_ref) in the build container and stores only the expected OUTPUTS under
tests/golden/; the tests regenerate the inputs and compare.
"""
import numpy as np

_A = np.uint64(6364136223846793005)
_C = np.uint64(1442695040888963407)


def lcg_stream(seed, n):
    """n pseudo-random uint32 values from a 64-bit LCG (high half of the state)."""
    out = np.empty(n, dtype=np.uint32)
    # vectorised in blocks: state_k for k consecutive steps via repeated squaring is overkill here,
    # a 4096-lane interleave keeps it fast and still deterministic
    lanes = 4096
    st = (np.arange(lanes, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed)) | np.uint64(1)
    with np.errstate(over="ignore"):
        for _ in range(4):
            st = st * _A + _C
        pos = 0
        while pos < n:
            st = st * _A + _C
            take = min(lanes, n - pos)
            out[pos:pos + take] = (st[:take] >> np.uint64(32)).astype(np.uint32)
            pos += take
    return out


def random_bits(seed, shape):
    n = int(np.prod(shape))
    return ((lcg_stream(seed, n) >> np.uint32(16)) & np.uint32(1)).astype(np.uint8).reshape(shape)


def random_bytes(seed, shape):
    n = int(np.prod(shape))
    return ((lcg_stream(seed, n) >> np.uint32(13)) & np.uint32(0xff)).astype(np.uint8).reshape(shape)


def noisy_llr(cw_bits, amp, sigma_q8, seed):
    """int8 LLRs: +-amp for bit 0/1 plus approximately Gaussian integer noise.
    noise = (sum of four 8-bit uniforms - 510) * sigma_q8 / 256 / 1.15 (sigma of the sum is ~148)."""
    cw_bits = np.asarray(cw_bits, dtype=np.uint8)
    n = cw_bits.size
    r = lcg_stream(seed, n).astype(np.int64)
    s = (r & 0xff) + ((r >> 8) & 0xff) + ((r >> 16) & 0xff) + ((r >> 24) & 0xff) - 510
    noise = (s * sigma_q8) // (256 * 148)
    x = (1 - 2 * cw_bits.astype(np.int64).reshape(-1)) * amp + noise
    return np.clip(x, -128, 127).astype(np.int8).reshape(cw_bits.shape)


def complex_symbols(seed, shape_syms, scale_q8=180):
    """float32 [..., 2] symbols with exactly representable values k/256 (|k| < 2^15)."""
    n = int(np.prod(shape_syms)) * 2
    r = lcg_stream(seed, n).astype(np.int64)
    s = (r & 0xff) + ((r >> 8) & 0xff) + ((r >> 16) & 0xff) + ((r >> 24) & 0xff) - 510
    k = (s * scale_q8) // 148
    return (k.astype(np.float32) / np.float32(256.0)).reshape(tuple(shape_syms) + (2,))
