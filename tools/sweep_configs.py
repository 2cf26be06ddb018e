#!/usr/bin/env python3
"""Throughput of every BASELINE.json config and of the standalone kernels on one B200 (run on the GPU
box: python tools/sweep_configs.py > gpurun_out/sweep.json).  Not the bench line: the five configs are
parity-test cases; this records what each costs, with the reference CPU path timed beside it."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gr-dvbs2rx_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ge.build()
import dvbs2rx_b200 as d  # noqa: E402
from dvbs2rx_b200 import vectors  # noqa: E402

HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3


def cpu_ref(info, llr, trials, fs):
    # the reference CPU leg lives in bench.py (the only measurement code that may execute oracle/)
    import bench
    name = d.lib().dvbs2b200_table_name(info.table).decode()
    return bench.cpu_reference_ldpc_rate(name, llr, trials)


CONFIGS = [
    ("C1 QPSK 1/2 normal, 25 it, 1.0 dB", 1, "C1_2", d.MOD_QPSK, 1.0, 25, 0),
    ("C1b QPSK 1/2 normal, 25 it, 2.0 dB (converging)", 1, "C1_2", d.MOD_QPSK, 2.0, 25, 0),
    ("C2 QPSK 3/4 normal, <=50 it early termination, 4.6 dB", 1, "C3_4", d.MOD_QPSK, 4.6, 50, 0),
    ("C2g same, reference batch-of-32 termination", 1, "C3_4", d.MOD_QPSK, 4.6, 50, 32),
    ("C3 8PSK 3/5 normal from symbols, 6.2 dB", 1, "C3_5", d.MOD_8PSK, 6.2, 25, 0),
    ("C4 2/3 short, 8192 frames, BCH t=12", 0, "C2_3", d.MOD_QPSK, 3.3, 25, 0),
    ("C5 9/10 normal", 1, "C9_10", d.MOD_QPSK, 6.6, 25, 0),
]


def main():
    out = []
    for name, fs, rate_name, mod, esn0, trials, group in CONFIGS:
        rate = d.RATE[rate_name]
        code = d.Code(0, fs, rate)
        info = code.info
        F = 8192 if fs == 0 else 2664
        if group:
            F = F // 416 * 416 if F >= 416 else F
        rng = np.random.default_rng(5)
        msg, cw, _ = vectors.encode_frames(0, fs, rate, 64, rng)
        cw_all = np.tile(cw, (F // 64 + 1, 1))[:F]
        iq, n0 = vectors.awgn(vectors.map_symbols(cw_all, mod, rate), esn0, rng)
        d_iq = torch.from_numpy(iq).to(dev)
        d_n0 = torch.full((F,), float(n0), dtype=torch.float32, device=dev)
        d_llr = torch.empty((F, info.n_ldpc), dtype=torch.int8, device=dev)
        d_msg = torch.empty((F, info.kbch // 8), dtype=torch.uint8, device=dev)
        d_tr = torch.empty(F, dtype=torch.int32, device=dev)
        d_co = torch.empty(F, dtype=torch.int32, device=dev)
        t_demap = timed(lambda: code.demap_dev(mod, d_iq.data_ptr(), F, d_n0.data_ptr(), d_llr.data_ptr(), stream))
        t_chain = timed(lambda: code.fec_decode_dev(mod, None, None, d_llr.data_ptr(), F, trials, group,
                                                    d_msg.data_ptr(), d_tr.data_ptr(), d_co.data_ptr(), stream))
        t_full = timed(lambda: code.fec_decode_dev(mod, d_iq.data_ptr(), d_n0.data_ptr(), None, F, trials, group,
                                                   d_msg.data_ptr(), d_tr.data_ptr(), d_co.data_ptr(), stream))
        tr = d_tr.cpu().numpy()
        ok = (d_msg.cpu().numpy() == np.tile(msg, (F // 64 + 1, 1))[:F]).all(axis=1)
        bits = d.bits_per_symbol(mod)
        demap_bytes = F * (8 * info.n_ldpc // bits + info.n_ldpc)
        rec = dict(config=name, frames=F, ldpc_bch_frames_per_s=F / t_chain, from_symbols_frames_per_s=F / t_full,
                   demap_ms=t_demap * 1e3, demap_GBps=demap_bytes / t_demap / 1e9, demap_frac_of_hbm=demap_bytes / t_demap / 1e9 / HBM,
                   converged_frac=float((tr >= 0).mean()), frames_correct_frac=float(ok.mean()),
                   mean_iterations=float(np.where(tr >= 0, trials - tr, trials).mean()),
                   cpu_reference=cpu_ref(info, d_llr.cpu().numpy(), trials, fs))
        out.append(rec)
        print(json.dumps(rec), flush=True)
        code.close()


if __name__ == "__main__":
    main()
