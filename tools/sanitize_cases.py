"""Small LDPC decodes that exercise every synchronisation protocol of the LDPC kernel, for
`compute-sanitizer --tool racecheck|synccheck python tools/sanitize_cases.py [case ...]`.

Cases (each compared with nothing here -- parity is the test suite's job; the sanitizer is the judge):
  pair      QPSK 1/2 normal   pair steps + chain-form split steps, per-frame stop
  level     9/10 normal       level-form split steps (node operands through shared memory, lane per shared link,
                              shuffle butterfly, block barrier per level), wide state
  short     2/3 short         chain and level form on a short frame
  c34       3/4 normal        chain + level mix, two CTAs per SM
  group     1/2 short         group-of-32 termination (cooperative launch, arrival counters), frames = 2 x resident CTAs
Environment knobs (DVBS2B200_CHAIN=0: level form for every split step) apply as usual.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gr-dvbs2rx_b200"))
sys.path.insert(0, ROOT)

import dvbs2rx_b200 as d  # noqa: E402
from dvbs2rx_b200 import vectors  # noqa: E402

CASES = {
    "pair": (1, "C1_2", 2.0, 4, 3, 0),
    "level": (1, "C9_10", 6.6, 4, 3, 0),
    "short": (0, "C2_3", 3.4, 6, 3, 0),
    "c34": (1, "C3_4", 4.6, 4, 3, 0),
    "group": (0, "C1_2", 1.6, 0, 4, 32),
}


def main(argv):
    names = argv or list(CASES)
    for name in names:
        fs, rate_name, esn0, frames, trials, group = CASES[name]
        rate = d.RATE[rate_name]
        code = d.Code(0, fs, rate)
        if group:
            frames = int(os.environ.get("SANITIZE_GROUP_FRAMES", "64"))
        msg, cw, llr, info = vectors.make_llr_frames(0, fs, rate, min(frames, 32), esn0, seed=5)
        if frames > llr.shape[0]:
            llr = np.concatenate([llr] * ((frames + llr.shape[0] - 1) // llr.shape[0]))[:frames]
        hard, post, left = code.ldpc_decode(llr, trials, group, d.OM_MESSAGE, want_post=True)
        print("%s: %d frames of %s %s, trials left %s" % (name, frames, "normal" if fs else "short", rate_name, left[:8].tolist()), flush=True)
        code.close()


if __name__ == "__main__":
    main(sys.argv[1:])
