#!/usr/bin/env python3
"""Generate tests/golden/*.json from the compiled, unmodified reference (oracle/_ref).

Build-container only (needs /root/reference to build oracle/_ref):
    python tools/gen_golden.py
Inputs come from tests/golden_inputs.py (pure-integer PRNG, reproducible anywhere); the fixtures
hold the reference's outputs: return values, sha256 of posterior LLR bytes, packed hard decisions
(hex), BCH outputs / return codes, 8PSK demapper bytes (sha256).
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gr-dvbs2rx_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

import __graft_entry__ as ge  # noqa: E402

ge.build()
import dvbs2rx_b200 as d  # noqa: E402
from dvbs2rx_b200 import vectors  # noqa: E402
import golden_inputs as gi  # noqa: E402
import oracle_lib  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
ref = oracle_lib.Ref()
orc = oracle_lib.Oracle()


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# (name, framesize, rate, amp, sigma_q8 list, trials): sigma chosen so that one case does not
# converge and one converges after a data-dependent number of iterations
LDPC_CASES = [
    ("c1_qpsk_1_2_normal", 1, "C1_2", 4, [1300, 800], 25),
    ("c2_qpsk_3_4_normal", 1, "C3_4", 6, [1150, 900], 50),
    ("c3_8psk_3_5_normal", 1, "C3_5", 5, [1350, 900], 25),
    ("c4_16apsk_2_3_short", 0, "C2_3", 6, [1300, 1000], 25),
    ("c5_32apsk_9_10_normal", 1, "C9_10", 8, [1000, 760], 25),
    ("s2_1_4_normal", 1, "C1_4", 3, [1300, 820], 25),
    ("s2x_1_3_medium", 2, "C1_3_MEDIUM", 3, [1200, 700], 25),
]


def gen_ldpc():
    cases = []
    for name, fs, rate_name, amp, sigmas, trials in LDPC_CASES:
        rate = d.RATE[rate_name]
        info = d.lookup(0, fs, rate)
        tname = orc.table_name(info.table)
        for sigma in sigmas:
            frames = 32  # one AVX2 batch = two generic/SSE4.1 batches
            seed = 1000 + 7 * rate + sigma
            msg_bits = gi.random_bits(seed, (frames, info.k_ldpc))
            cw = vectors.ldpc_encode_bits(info.table, msg_bits)
            llr = gi.noisy_llr(cw, amp, sigma, seed + 1)
            post32, ret32 = ref.ldpc_decode(tname, llr, trials, isa="avx2")
            post16, ret16 = ref.ldpc_decode(tname, llr, trials, isa="generic")
            post16s, ret16s = ref.ldpc_decode(tname, llr, trials, isa="sse41")
            assert np.array_equal(post16, post16s) and np.array_equal(ret16, ret16s)
            hard32 = orc.pack_hard(post32, info.n_ldpc)
            hard16 = orc.pack_hard(post16, info.n_ldpc)
            cases.append(dict(
                name=name, framesize=fs, rate=rate_name, amp=amp, sigma_q8=sigma, trials=trials, frames=frames,
                seed=seed, table=tname, llr_sha256=sha(llr),
                group32=dict(ret=ret32.tolist(), post_sha256=sha(post32), hard_sha256=sha(hard32),
                             hard_frame0_hex=hard32[0].tobytes().hex()[:128]),
                group16=dict(ret=ret16.tolist(), post_sha256=sha(post16), hard_sha256=sha(hard16)),
                bit_errors=int((vectors.unpack_bits(hard32) != cw).sum()),
            ))
            print(name, sigma, "ret32", sorted(set(ret32.tolist())), "ret16", sorted(set(ret16.tolist())),
                  "bit errors", cases[-1]["bit_errors"])
    with open(os.path.join(OUT, "ldpc.json"), "w") as f:
        json.dump(dict(generator="tools/gen_golden.py", reference="igorauad/gr-dvbs2rx v1.4.0 (130c315), "
                       "ldpc_decoder_{avx2,sse41,generic}.cc compiled unmodified", cases=cases), f, indent=1)


BCH_CASES = [("normal_1_2", 1, "C1_2"), ("short_2_3", 0, "C2_3"), ("normal_9_10", 1, "C9_10"), ("normal_2_3", 1, "C2_3"),
             ("normal_3_4", 1, "C3_4"), ("normal_3_5", 1, "C3_5")]


def gen_bch():
    cases = []
    for name, fs, rate_name in BCH_CASES:
        rate = d.RATE[rate_name]
        info = d.lookup(0, fs, rate)
        h = ref.bch(fs, info.t, info.nbch)
        nerr = [0, 1, 2, 3, info.t - 1, info.t, info.t + 1, 13, 30, 200] * 2
        F = len(nerr)
        seed = 5000 + rate
        msg = gi.random_bytes(seed, (F, info.kbch // 8))
        cw = ref.bch_encode(h, msg, info.nbch)
        pos = gi.lcg_stream(seed + 1, F * 256).reshape(F, 256) % np.uint32(info.nbch)
        for f in range(F):
            seen = []
            for p in pos[f]:
                if len(seen) == nerr[f]:
                    break
                if int(p) not in seen:
                    seen.append(int(p))
                    cw[f, p >> 3] ^= 0x80 >> (p & 7)
        garbage = gi.random_bytes(seed + 2, (6, info.nbch // 8))
        allcw = np.concatenate([cw, garbage])
        out, ret = ref.bch_decode(h, allcw, info.nbch)
        g = np.zeros(256, np.uint8)
        deg = ref.l.ref_bch_genpoly(h, g.ctypes.data, 256)
        cases.append(dict(name=name, framesize=fs, rate=rate_name, n=info.nbch, k=info.kbch, t=info.t, seed=seed,
                          nerr=nerr, garbage_frames=6, cw_sha256=sha(allcw), ret=ret.tolist(), out_sha256=sha(out),
                          encoded_sha256=sha(ref.bch_encode(h, msg, info.nbch)),
                          genpoly="".join(str(int(x)) for x in g[:deg + 1])))
        print(name, ret.tolist())
    with open(os.path.join(OUT, "bch.json"), "w") as f:
        json.dump(dict(generator="tools/gen_golden.py", reference="lib/gf.cc + lib/bch.cc compiled unmodified",
                       cases=cases), f, indent=1)


def gen_demap():
    cases = []
    for rate_name in ("C3_5", "C2_3", "C25_36"):
        rate = d.RATE[rate_name]
        for n0 in (0.24, 0.05, 0.9):
            seed = 9000 + rate
            iq = gi.complex_symbols(seed, (3, 21600))
            out = ref.demap_8psk(iq, n0, vectors.rows_8psk(rate, 21600))
            cases.append(dict(constellation="MOD_8PSK", rate=rate_name, n0=n0, seed=seed, frames=3, n_syms=21600,
                              iq_sha256=sha(iq), llr_sha256=sha(out), llr_head=out[0, :24].tolist()))
    with open(os.path.join(OUT, "demap.json"), "w") as f:
        json.dump(dict(generator="tools/gen_golden.py",
                       reference="lib/psk.hh driven as lib/xfecframe_demapper_cb_impl.cc:148-176, -ffp-contract=off",
                       cases=cases), f, indent=1)


def gen_bb():
    """BB deheader: the reference's bbdeheader_bb_impl.cc (compiled unmodified over the gr::block shim) on
    the streams of tests/bb_cases.py."""
    import bb_cases
    rate_of = {16008: (1, "C1_4"), 3072: (0, "C1_4"), 58192: (1, "C9_10")}
    cases = {}
    for name, kbch, calls in bb_cases.make_cases():
        fs, rate = rate_of[kbch]
        r = ref.bbdeheader(0, fs, d.RATE[rate], kbch)
        out = np.concatenate([r.work(bb) for bb in calls])
        cases[name] = dict(kbch=kbch, calls=[int(bb.shape[0]) for bb in calls], ts_bytes=int(out.size), sha256=sha(out),
                           counters=r.counters())
        print(name, out.size, r.counters())
    zero = np.zeros((1, 58192 // 8), dtype=np.uint8)
    prbs = ref.bb_descramble(0, 1, d.RATE["C9_10"], zero)
    with open(os.path.join(OUT, "bb.json"), "w") as f:
        json.dump(dict(generator="tools/gen_golden.py",
                       reference="lib/bbdeheader_bb_impl.cc + lib/bbdescrambler_bb_impl.cc compiled unmodified (oracle/ref_bb_harness.cc)",
                       prbs_sha256=sha(prbs), cases=cases), f, indent=1)


if __name__ == "__main__":
    which = sys.argv[1:] or ["ldpc", "bch", "demap", "bb"]
    if "ldpc" in which:
        gen_ldpc()
    if "bch" in which:
        gen_bch()
    if "demap" in which:
        gen_demap()
    if "bb" in which:
        gen_bb()
