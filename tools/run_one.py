#!/usr/bin/env python3
"""One pass of a chosen piece of the path over a synthetic batch, for profiling under ncu on the GPU box:

    python tools/run_one.py fec   C3_4 1 4.6 50 2664     # LDPC + BCH (device-resident input)
    python tools/run_one.py ts    C1_2 1 2.0 25 2664     # LDPC + BCH + descrambler + deheader -> TS packets
    python tools/run_one.py bb    C1_2 1 0 0 8192        # descrambler + deheader alone on clean BBFRAMEs
    python tools/run_one.py snr   C3_5 1 6.2 0 2664      # SNR estimate (8PSK for 3/5, else QPSK) + demap
    python tools/run_one.py mixed C1_2 1 0 25 5000       # five MODCODs interleaved in one batch (host API)
    python tools/run_one.py apsk  C9_10 1 17.5 0 2664    # table-driven demapper (32APSK for 8/9, 9/10, else 16APSK)
    python tools/run_one.py pl    C1_2 1 0 0 2664        # PL descrambler + pilot de-rotation (QPSK normal frame with pilots)

Prints device time per pass (CUDA events) and the derived rates; not the bench line."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gr-dvbs2rx_b200"))
import dvbs2rx_b200 as d  # noqa: E402
from dvbs2rx_b200 import bbframes as bbf, vectors  # noqa: E402

what, rate_name = sys.argv[1], sys.argv[2]
fs, esn0, trials, F = int(sys.argv[3]), float(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
rate = d.RATE[rate_name]
code = d.Code(0, fs, rate)
info = code.info
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream
rng = np.random.default_rng(1)
kb = info.kbch // 8


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3


def bb_stream(n):
    up = bbf.ts_packets((n * (kb - 10) + 187) // 188 + 1, rng)
    return up, bbf.bbframe_stream(info.kbch, n, up)


if what in ("fec", "ts"):
    base = 64
    up, bb = bb_stream(base)
    msg_bits = vectors.unpack_bits(bbf.scramble(bb), info.kbch)
    cw = vectors.ldpc_encode_bits(info.table, vectors.bch_encode_bits(msg_bits, fs, info.t, info.nbch)[:, :info.k_ldpc])
    iq, n0 = vectors.awgn(vectors.map_symbols(cw, d.MOD_QPSK, rate), esn0, rng)
    llr = np.tile(vectors.qpsk_llr(iq, n0), (F // base + 1, 1))[:F]
    d_llr = torch.from_numpy(llr).to(dev)
    d_tr = torch.empty(F, dtype=torch.int32, device=dev)
    d_co = torch.empty(F, dtype=torch.int32, device=dev)
    if what == "fec":
        d_msg = torch.empty((F, kb), dtype=torch.uint8, device=dev)
        t = timed(lambda: code.fec_decode_dev(0, None, None, d_llr.data_ptr(), F, trials, 0, d_msg.data_ptr(), d_tr.data_ptr(),
                                              d_co.data_ptr(), stream))
        print("fec %s: %.3f ms, %.0f frames/s, mean iterations %.1f" % (
            rate_name, t * 1e3, F / t, float(np.where(d_tr.cpu().numpy() >= 0, trials - d_tr.cpu().numpy(), trials).mean())))
    else:
        cap = code.bb_ts_capacity(F)
        d_ts = torch.empty(cap, dtype=torch.uint8, device=dev)

        def run():
            code.bb_reset()
            code.fec_decode_ts_dev(0, None, None, d_llr.data_ptr(), F, trials, 0, d_ts.data_ptr(), cap, d_tr.data_ptr(), d_co.data_ptr(), stream)
        t = timed(run)
        print("ts %s: %.3f ms, %.0f frames/s, %d TS bytes, counters %s" % (rate_name, t * 1e3, F / t, code.bb_produced_dev(stream),
                                                                        code.bb_counters()))
elif what == "mixed":
    # BASELINE config 5: the five MODCODs interleaved in one batch, per-frame code id, host buffers in and out
    import time
    modcods = [(0, 1, d.C1_2, 2.0), (0, 1, d.C3_4, 4.6), (0, 1, d.C3_5, 3.5), (0, 0, d.C2_3, 4.2), (0, 1, d.C9_10, 6.6)]
    mixed = d.MixedCodes([m[:3] for m in modcods])
    per = F // len(modcods)
    llrs, ids = [], []
    for c, (std, f_s, r, e) in enumerate(modcods):
        msg, cw, llr, inf = vectors.make_llr_frames(std, f_s, r, 32, e, seed=70 + c)
        llrs.append(np.tile(llr, (per // 32 + 1, 1))[:per])
        ids.append(np.full(per, c, dtype=np.uint8))
    order = np.concatenate(ids)
    perm = rng.permutation(order.size)
    order = order[perm]
    counters = [0] * len(modcods)
    parts = []
    for c in order:
        parts.append(llrs[c][counters[c]])
        counters[c] += 1
    cat = np.concatenate(parts)
    mixed.fec_decode(order, cat, trials)
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        msg, tr, co = mixed.fec_decode(order, cat, trials)
    t = (time.perf_counter() - t0) / reps
    print("mixed: %d frames of %d MODCODs, %.1f ms per batch through the host API, %.0f frames/s, %.0f%% converged" % (
        order.size, len(modcods), t * 1e3, order.size / t, 100.0 * (tr >= 0).mean()))
elif what == "apsk":
    # table-driven demapper: 16APSK for 2/3, 3/4, 4/5, 5/6 (bits = 4), 32APSK for 8/9, 9/10 (bits = 5)
    from dvbs2rx_b200 import apsk
    if rate_name in ("C8_9", "C9_10"):
        pts = apsk.points_32apsk(*apsk.GAMMA_32APSK[rate_name])
    else:
        pts = apsk.points_16apsk(apsk.GAMMA_16APSK[rate_name])
    bits = int(pts.shape[0]).bit_length() - 1
    offs = apsk.row_offsets(info.n_ldpc, bits)
    msg, cw, info2 = vectors.encode_frames(0, fs, rate, 8, rng)
    iq, n0 = vectors.awgn(apsk.map_bits(cw, pts, offs), esn0, rng)
    iq = np.tile(iq, (F // 8 + 1, 1, 1))[:F]
    d_iq = torch.from_numpy(iq).to(dev)
    d_n0 = torch.full((F,), float(n0), dtype=torch.float32, device=dev)
    d_llr = torch.empty((F, info.n_ldpc), dtype=torch.int8, device=dev)
    t = timed(lambda: code.demap_table_dev(pts, offs, d_iq.data_ptr(), F, d_n0.data_ptr(), d_llr.data_ptr(), stream), reps=10)
    nbytes = F * (info.n_ldpc // bits) * 8 + F * info.n_ldpc
    print("apsk %s (%d points): %.1f us per %d frames, %.0f GB/s of symbols in + LLRs out, %.1f M symbols/ms" % (
        rate_name, pts.shape[0], t * 1e6, F, nbytes / t / 1e9, F * (info.n_ldpc // bits) / t / 1e9))
elif what == "bb":
    up, bb = bb_stream(F)
    d_bb = torch.from_numpy(bbf.scramble(bb)).to(dev)
    cap = code.bb_ts_capacity(F)
    d_ts = torch.empty(cap, dtype=torch.uint8, device=dev)

    def run():
        code.bb_reset()
        code.bb_deheader_dev(d_bb.data_ptr(), F, 1, d_ts.data_ptr(), cap, stream)
    t = timed(run, reps=10)
    n = code.bb_produced_dev(stream)
    ok = np.array_equal(d_ts[:n].cpu().numpy(), up[:n // 188].ravel())
    print("bb %s: %.1f us per call of %d BBFRAMEs, %.1f GB/s (read kbch/8 + write TS), packets ok %s" % (
        rate_name, t * 1e6, F, (F * kb + n) / t / 1e9, ok))
elif what == "snr":
    mod = d.MOD_8PSK if rate_name == "C3_5" else d.MOD_QPSK
    msg, cw, info2 = vectors.encode_frames(0, fs, rate, 16, rng)
    iq, n0 = vectors.awgn(vectors.map_symbols(cw, mod, rate), esn0, rng)
    iq = np.tile(iq, (F // 16 + 1, 1, 1))[:F]
    d_iq = torch.from_numpy(iq).to(dev)
    d_snr = torch.empty(F, dtype=torch.float32, device=dev)
    d_n0 = torch.full((F,), float(n0), dtype=torch.float32, device=dev)
    d_llr = torch.empty((F, info.n_ldpc), dtype=torch.int8, device=dev)
    t1 = timed(lambda: code.estimate_snr_dev(mod, d_iq.data_ptr(), None, F, d_snr.data_ptr(), stream), reps=10)
    t2 = timed(lambda: code.demap_dev(mod, d_iq.data_ptr(), F, d_n0.data_ptr(), d_llr.data_ptr(), stream), reps=10)
    t3 = timed(lambda: code.estimate_snr_dev(mod, d_iq.data_ptr(), d_llr.data_ptr(), F, d_snr.data_ptr(), stream), reps=10)
    bits = d.bits_per_symbol(mod)
    iq_bytes = F * (info.n_ldpc // bits) * 8
    print("snr %s: symbols-only %.1f us (%.0f GB/s), demap %.1f us (%.0f GB/s), with LLR %.1f us (%.0f GB/s); mean %.2f dB" % (
        rate_name, t1 * 1e6, iq_bytes / t1 / 1e9, t2 * 1e6, (iq_bytes + F * info.n_ldpc) / t2 / 1e9, t3 * 1e6,
        (iq_bytes + F * info.n_ldpc) / t3 / 1e9, float(10 * np.log10(d_snr.cpu().numpy().mean()))))
elif what == "pl":
    # PL descrambler + pilot de-rotation: QPSK normal frame with pilots (360 slots, 22 pilot blocks)
    n_slots, pilots = 360, 1
    plen = d.PlDescrambler.payload_len(n_slots, pilots)
    pl = d.PlDescrambler(0, 0)
    d_pay = torch.randn((F, plen, 2), dtype=torch.float32, device=dev)
    finfo = np.zeros(F, dtype=d.PL_FRAME_DTYPE)
    finfo["plheader_phase"] = rng.uniform(-3, 3, F)
    finfo["fine_foffset"] = rng.uniform(-1e-4, 1e-4, F)
    finfo["coarse_corrected"] = 1
    finfo["pilot_phase"] = rng.uniform(-3, 3, (F, 22))
    d_info = torch.from_numpy(finfo.view(np.uint8)).to(dev)
    d_out = torch.empty((F, n_slots * 90, 2), dtype=torch.float32, device=dev)
    t = timed(lambda: pl.process_dev(d_pay.data_ptr(), F, n_slots, pilots, d_info.data_ptr(), d_out.data_ptr(), stream), reps=10)
    nbytes = F * (plen + n_slots * 90) * 8
    print("pl: %.1f us per %d PLFRAME payloads (QPSK normal, pilots), %.0f GB/s of symbols in + out, %.2f M frames/s" % (
        t * 1e6, F, nbytes / t / 1e9, F / t / 1e6))
