#!/usr/bin/env python3
"""Golden fixture for the PL descrambler: the scrambling codes R_n of several Gold codes as produced by the
reference's own lib/pl_descrambler.cc, compiled unmodified into oracle/_ref (container only: needs /root/reference).

    python tools/gen_golden_pl.py   ->  tests/golden/pl.json
Stored per Gold code: SHA-256 of the 33192 codes (one byte each), the first 64 codes, the histogram."""
import ctypes as C
import hashlib
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libdvbs2_ref.so"))
N = 360 * 90 + 22 * 36
out = {"n": N, "source": "lib/pl_descrambler.cc compiled unmodified (oracle/Makefile ref, oracle/ref_bb_harness.cc:ref_pl_rn)", "gold_codes": {}}
for g in (0, 1, 2, 13, 1000, 7777, 131072, 262141):
    rn = np.zeros(N, np.uint8)
    assert ref.ref_pl_rn(g, rn.ctypes.data, N) == 0
    out["gold_codes"][str(g)] = {"sha256": hashlib.sha256(rn.tobytes()).hexdigest(), "first64": rn[:64].tolist(),
                                 "histogram": np.bincount(rn, minlength=4).tolist()}
with open(os.path.join(ROOT, "tests", "golden", "pl.json"), "w") as f:
    json.dump(out, f, indent=1)
print("wrote tests/golden/pl.json")
