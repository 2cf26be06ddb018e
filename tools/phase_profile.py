#!/usr/bin/env python3
"""Per-phase cycle profile of the LDPC kernel (diagnostics).  Needs a library built with
-DDVBS2_PHASE_PROFILE, e.g. gr-dvbs2rx_b200/libdvbs2_b200_prof.so, selected with DVBS2B200_LIB:

    DVBS2B200_LIB=$PWD/gr-dvbs2rx_b200/libdvbs2_b200_prof.so python tools/phase_profile.py C1_2 1 1.0 25

Prints mean cycles per CTA and phase (measured on thread 0 of each CTA) and the frame rate."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gr-dvbs2rx_b200"))
os.environ.setdefault("DVBS2B200_PHASE_PROFILE", "/tmp/dvbs2_phase.txt")
import dvbs2rx_b200 as d  # noqa: E402
from dvbs2rx_b200 import vectors  # noqa: E402

rate = sys.argv[1] if len(sys.argv) > 1 else "C1_2"
fs = int(sys.argv[2]) if len(sys.argv) > 2 else 1
esn0 = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
trials = int(sys.argv[4]) if len(sys.argv) > 4 else 25
frames = int(sys.argv[5]) if len(sys.argv) > 5 else 888
msg, cw, llr, info = vectors.make_llr_frames(d.STANDARD_DVBS2, fs, d.RATE[rate], 64, esn0, seed=5)
import numpy as np  # noqa: E402
llr = np.tile(llr, (frames // 64 + 1, 1))[:frames]
code = d.Code(d.STANDARD_DVBS2, fs, d.RATE[rate])
code.ldpc_decode(llr, trials, d.TERM_PER_FRAME, d.OM_MESSAGE)
t0 = time.time()
code.ldpc_decode(llr, trials, d.TERM_PER_FRAME, d.OM_MESSAGE)
dt = time.time() - t0
print("%s fs=%d esn0=%.1f trials=%d frames=%d  host-call %.1f ms (incl. copies and profile readback)" % (rate, fs, esn0, trials, frames, dt * 1e3))
print(open(os.environ["DVBS2B200_PHASE_PROFILE"]).read())
