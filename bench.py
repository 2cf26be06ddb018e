#!/usr/bin/env python3
"""bench.py -- FECFRAMEs/s of the DVB-S2 FEC decode hot path (LDPC -> BCH) on B200.

Workloads
  c1 (default; BASELINE.json configs[0], the configuration the metric is quoted on): QPSK 1/2 normal
     FECFRAMEs (64800 bit), 25 offset-min-sum iterations, AWGN Es/N0 = 1.0 dB (no frame converges, so every
     frame runs all 25 iterations and takes the BCH failure path), int8 LLR input, decoded BBFRAME bytes out.
  mixed (BASELINE.json configs[4]): five MODCODs interleaved in one batch with a per-frame code id
     (1/2, 3/4, 3/5 normal, 2/3 short, 9/10 normal), the batch sharded across the ranks by frame count
     (sharding.shard_mixed); every rank checks a sample of its shard against the per-code oracle and the
     mismatch count is all-reduced (must be 0).
A "step" is one pass of LDPC+BCH over one batch of frames per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c1|mixed]

N > 1 is launched by torchrun (one rank per GPU); frames shard across ranks with no data-path collective
(weak scaling: per-GPU batch fixed); the code tables are built on rank 0 and broadcast once over NCCL.
--impl reference times the reference's own CPU implementation (oracle/_ref: its unmodified translation
units, AVX2 SIMD, one decoder per host thread) and never loads the product library.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "gr-dvbs2rx_b200"))

ESN0_DB = 1.0
MAX_TRIALS = 25
FRAMES_PER_GPU = 148 * 18  # 2664 frames = 6 per resident CTA (3 CTAs x 148 SMs); 173 MB of LLRs per step > 126 MB L2
WORKLOAD = "QPSK 1/2 normal FECFRAME (64800), 25 iters, AWGN Es/N0=1.0 dB, LDPC+BCH, int8 LLR in"
# (standard, framesize, rate name, Es/N0 of the QPSK-equivalent LLRs): BASELINE configs 1-5's codes
MIXED_MODCODS = [(0, 1, "C1_2", 2.0), (0, 1, "C3_4", 4.6), (0, 1, "C3_5", 3.5), (0, 0, "C2_3", 4.2), (0, 1, "C9_10", 6.6)]
MIXED_FRAMES_PER_CODE = 592  # x 5 codes = 2960 frames per GPU per step, 163 MB of LLRs > L2
MIXED_WORKLOAD = ("mixed-MODCOD batch: 1/2, 3/4, 3/5 normal, 2/3 short, 9/10 normal interleaved, per-frame code id, "
                  "25 max iters per-frame stop, LDPC+BCH, int8 LLR in")
RATE_ORD = {"C1_2": 3, "C3_5": 4, "C2_3": 5, "C3_4": 6, "C9_10": 11}  # dvb_config.h ordinals (no product import)


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def window(self, t0, t1):
        """Keep only the samples taken while the GPU was under load ([t0, t1] in perf_counter time)."""
        self.t0, self.t1 = t0, t1

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0, t1 = getattr(self, "t0", None), getattr(self, "t1", None)
        lines = [l for t, l in self.samples if t0 is None or t0 <= t <= t1] or [l for _, l in self.samples[-3:]]
        for s in lines:
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                pw.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w": float(np.median(pw)) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(frames, seed):
    import dvbs2rx_b200 as d
    from dvbs2rx_b200 import vectors
    msg, cw, llr, info = vectors.make_llr_frames(d.STANDARD_DVBS2, d.FECFRAME_NORMAL, d.C1_2, frames, ESN0_DB, seed)
    return msg, llr, info


# --------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own CPU implementation (oracle/_ref).  Nothing here
# imports or loads the product library: inputs come from the oracle's encoders.
# --------------------------------------------------------------------------------------------
def _oracle_lib():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    return oracle_lib


def ref_inputs(orc, standard, framesize, rate, frames, esn0_db, seed):
    """Random BBFRAMEs -> BCH -> LDPC (oracle encoders) -> QPSK + AWGN -> int8 LLRs (lib/qpsk.h:208-214)."""
    table, kbch, nbch, t = orc.lookup(standard, framesize, rate)
    rng = np.random.default_rng(seed)
    msg = rng.integers(0, 256, size=(frames, kbch // 8), dtype=np.uint8)
    bch = orc.bch(framesize, t, nbch)
    cw_bch = np.unpackbits(orc.bch_encode(bch, msg), axis=1)
    K = orc.l.orc_table_k(table)
    if K > nbch:
        cw_bch = np.concatenate([cw_bch, rng.integers(0, 2, size=(frames, K - nbch), dtype=np.uint8)], axis=1)
    cw = orc.ldpc_encode(table, cw_bch[:, :K])
    n0 = np.float32(10.0 ** (-esn0_db / 10.0))
    sym = np.float32(np.sqrt(0.5)) * (1 - 2 * cw.astype(np.float32))
    sym += rng.standard_normal(size=sym.shape, dtype=np.float32) * np.float32(np.sqrt(n0 / 2))
    scalar = np.float32(2 * np.sqrt(2.0) / np.float64(n0))
    llr = np.clip(np.rint(sym * scalar), -128, 127).astype(np.int8)
    return llr, dict(table=table, kbch=kbch, nbch=nbch, t=t, framesize=framesize, name=orc.table_name(table))


def cpu_reference_run(frames, steps, warmup, threads, modcods=None):
    """LDPC (AVX2, 32 frames per SIMD batch, one decoder instance per thread) + BCH on the resulting bytes,
    exactly the reference's translation units; for a list of MODCODs one such run per code per step (the
    reference is CCM: a mixed stream is one chain per MODCOD).  Returns frames/s, ms/step, kind, threads."""
    oracle_lib = _oracle_lib()
    kind = "reference" if os.path.exists(oracle_lib.REF_PATH) else "port"
    orc = oracle_lib.Oracle()
    modcods = modcods or [(0, 1, "C1_2", ESN0_DB)]
    sets = []
    for c, (std, fs, rate_name, esn0) in enumerate(modcods):
        llr, meta = ref_inputs(orc, std, fs, RATE_ORD[rate_name], frames, esn0, seed=1234 + c)
        sets.append((llr, meta))
    times = []
    if kind == "reference":
        ref = oracle_lib.Ref()
        for it in range(warmup + steps):
            tot = 0.0
            for llr, meta in sets:
                bch = ref.bch(meta["framesize"], meta["t"], meta["nbch"])
                post, ret, t_ldpc = ref.ldpc_decode_mt(meta["name"], llr, MAX_TRIALS, threads)
                hard = np.ascontiguousarray(np.packbits(post[:, :meta["nbch"]] < 0, axis=1))  # untimed glue
                out, corr, t_bch = ref.bch_decode_mt(bch, hard, meta["nbch"], threads)
                tot += t_ldpc + t_bch
            if it >= warmup:
                times.append(tot)
    else:
        threads = 1
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            for llr, meta in sets:
                bch = orc.bch(meta["framesize"], meta["t"], meta["nbch"])
                post, ret = orc.ldpc_decode(meta["table"], llr, MAX_TRIALS, lanes=32)
                out, corr = orc.bch_decode(bch, orc.pack_hard(post, meta["nbch"]))
            if it >= warmup:
                times.append(time.perf_counter() - t0)
    total = float(np.sum(times))
    nframes = frames * len(sets)
    return nframes * len(times) / total, 1e3 * total / len(times), kind, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    mixed = args.workload == "mixed"
    frames = 32 * threads * (1 if mixed else 2)  # bounded sample: SIMD batches per host thread (per code) per step
    fps, ms, kind, threads = cpu_reference_run(frames, args.steps, args.warmup, threads, MIXED_MODCODS if mixed else None)
    nf = frames * (len(MIXED_MODCODS) if mixed else 1)
    sample = "%d frames per step (32-frame AVX2 batches%s), %d host threads" % (nf, ", five single-code runs" if mixed else "", threads)
    emit(json.dumps({
        "impl": "reference", "metric": "FECFRAMEs/s", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int8", "data": "synthetic",
        "config": {"workload": MIXED_WORKLOAD if mixed else WORKLOAD, "frames_per_step": nf, "max_trials": MAX_TRIALS,
                   "esn0_db": [m[3] for m in MIXED_MODCODS] if mixed else ESN0_DB, "term": "reference SIMD batch of 32"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def cpu_reference_ldpc_rate(table_name, llr, trials, threads=None):
    """Frames/s of the reference's own AVX2 LDPC decoder (oracle/_ref) on `llr`, one decoder instance per host
    thread; None where oracle/_ref is not built.  The one CPU leg other measurement tools (tools/sweep_configs.py)
    go through, so that nothing outside tests/, smoke() and this file touches oracle/."""
    oracle_lib = _oracle_lib()
    if not os.path.exists(oracle_lib.REF_PATH):
        return None
    ref = oracle_lib.Ref()
    threads = threads or (os.cpu_count() or 1)
    F = min(32 * threads, llr.shape[0] // 32 * 32)
    ref.ldpc_decode_mt(table_name, llr[:F], trials, threads)
    post, ret, t1 = ref.ldpc_decode_mt(table_name, llr[:F], trials, threads)
    return dict(ldpc_frames_per_s=F / t1, threads=threads, frames=F)


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def _issue_roofline():
    """Issue-side view of the LDPC kernel, from the committed ncu metrics of the same bench command
    (profiles/ldpc_issue.json, written by tools/summarize_ncu.py): the kernel is bound by instruction issue /
    the ALU pipe, not HBM, so this is the fraction that says how close to the machine it runs."""
    try:
        with open(os.path.join(ROOT, "profiles", "ldpc_issue.json")) as f:
            return json.load(f)
    except Exception:
        return None


def run_ours(args):
    import torch
    import torch.distributed as dist
    import dvbs2rx_b200 as d

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.workload == "mixed":
        return run_ours_mixed(args, torch, dist, d, world, rank, local, dev)

    # ---- code tables: built on rank 0, ONE NCCL broadcast, every rank creates from the blob ----
    from dvbs2rx_b200 import sharding
    code = sharding.make_code(d.STANDARD_DVBS2, d.FECFRAME_NORMAL, d.C1_2, local)
    info = code.info
    F = FRAMES_PER_GPU
    N, nb, kb = info.n_ldpc, info.nbch // 8, info.kbch // 8

    # ---- synthetic input: every rank its own frames (shard = contiguous frame range) ------------
    msg, llr_np, _ = make_inputs(F, seed=1000 + rank)
    h_llr = torch.from_numpy(llr_np).pin_memory()
    # two device copies of the batch at different addresses, used alternately: a step never finds its input in L2
    d_llrs = [h_llr.to(dev, non_blocking=True), torch.roll(h_llr, 1, 0).to(dev, non_blocking=True)]
    d_mid = torch.empty((F, nb), dtype=torch.uint8, device=dev)
    d_msg = torch.empty((F, kb), dtype=torch.uint8, device=dev)
    d_trials = torch.empty(F, dtype=torch.int32, device=dev)
    d_corr = torch.empty(F, dtype=torch.int32, device=dev)
    h_msg = torch.empty((F, kb), dtype=torch.uint8).pin_memory()
    h_trials = torch.empty(F, dtype=torch.int32).pin_memory()
    h_corr = torch.empty(F, dtype=torch.int32).pin_memory()
    stream = torch.cuda.current_stream().cuda_stream

    def step(k, ev=None):
        if ev:
            ev[0].record()
        code.ldpc_decode_dev(d_llrs[k & 1].data_ptr(), F, MAX_TRIALS, d.TERM_PER_FRAME, d.OM_MESSAGE, d_mid.data_ptr(),
                             None, d_trials.data_ptr(), stream)
        if ev:
            ev[1].record()
        code.bch_decode_dev(d_mid.data_ptr(), F, d_msg.data_ptr(), d_corr.data_ptr(), stream)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # nvidia-smi takes a moment to come up: started ahead of the warm-up, windowed below
    warm = max(args.warmup, 3)
    for k in range(warm):
        step(k + 1)
    sync_all()
    launches0 = code.launches
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    t_load0 = time.perf_counter()
    t_start.record()
    for k in range(args.steps):
        step(k, evs[k])
    t_end.record()
    sync_all()
    launches = code.launches - launches0
    elapsed_ms = t_start.elapsed_time(t_end)
    ldpc_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    el = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
    elapsed_ms = float(el.item())
    value = world * F * args.steps / (elapsed_ms * 1e-3)

    # sanity: the timed path produced what the host API produces (and no frame converged)
    assert int((d_trials.cpu() == -1).sum()) == F, "config 1 frames are expected not to converge"
    step(0)  # the reference result for the checks below: the un-rolled batch
    torch.cuda.synchronize()

    # ---- end to end through the host C ABI: H2D + D2H inside the timed region ------------------------
    def e2e_run(llr_ptr, msg_ptr, tr_ptr, co_ptr, steps):
        def one():
            code.fec_decode_ptr(d.MOD_QPSK, None, None, llr_ptr, F, MAX_TRIALS, d.TERM_PER_FRAME, msg_ptr, tr_ptr, co_ptr)
        for _ in range(2):
            one()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(steps):
            one()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return world * F * steps / float(dt.item())

    e2e_steps = max(3, min(args.steps, 10))
    e2e_value = e2e_pageable = e2e_registered = None
    if not args.device_only:
        e2e_value = e2e_run(h_llr.data_ptr(), h_msg.data_ptr(), h_trials.data_ptr(), h_corr.data_ptr(), e2e_steps)
        assert torch.equal(h_msg, d_msg.cpu()), "host-API result differs from the device-resident path"
        # the same call with PAGEABLE host buffers (what a GNU Radio ring buffer is): the library stages them
        # through its pinned ring
        p_msg = np.empty((F, kb), dtype=np.uint8)
        p_tr, p_co = np.empty(F, dtype=np.int32), np.empty(F, dtype=np.int32)
        e2e_pageable = e2e_run(llr_np.ctypes.data, p_msg.ctypes.data, p_tr.ctypes.data, p_co.ctypes.data, e2e_steps)
        assert np.array_equal(p_msg, h_msg.numpy()), "pageable-buffer result differs from the pinned-buffer one"
        # ... and with the same buffers page-locked in place once (dvbs2b200_host_register: what a block does with its
        # ring buffers in start()); the registration is outside the timed region, as it is outside general_work
        for a in (llr_np, p_msg, p_tr, p_co):
            d.host_register(a)
        p_msg[:] = 0
        e2e_registered = e2e_run(llr_np.ctypes.data, p_msg.ctypes.data, p_tr.ctypes.data, p_co.ctypes.data, e2e_steps)
        for a in (llr_np, p_msg, p_tr, p_co):
            d.host_unregister(a)
        assert np.array_equal(p_msg, h_msg.numpy()), "registered-buffer result differs from the pinned-buffer one"
    # clocks / throttle reasons sampled while the GPU was busy: the timed region and the end-to-end loops after it
    sampler.window(t_load0, time.perf_counter())
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        peak, peak_kind = load_peaks()
        alg_bytes = F * (N + nb)  # soft input read once + packed hard decisions written once
        achieved = alg_bytes / (ldpc_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ldpc_traffic.json")
        if os.path.exists(tpath):
            try:
                with open(tpath) as f:
                    traffic = json.load(f).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        out = {
            "metric": "FECFRAMEs/s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_gpu_per_step": F, "max_trials": MAX_TRIALS,
                       "esn0_db": ESN0_DB, "term": "per-frame",
                       "l2": "inputs larger than L2: 173 MB of LLRs per GPU per step, two device copies used alternately",
                       "sharding": "frames sharded across ranks, tables broadcast once over NCCL"},
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": world * F * N,
                    "d2h_bytes_per_step": world * F * (kb + 8), "buffers": "pinned host"},
            "e2e_pageable": {"value": e2e_pageable, "unit": "frames/s", "buffers": "pageable host (staged through the handle's pinned ring)"},
            "e2e_registered": {"value": e2e_registered, "unit": "frames/s",
                               "buffers": "the same pageable buffers page-locked in place once (dvbs2b200_host_register)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "ldpc_decode_kernel", "achieved": achieved, "peak": peak,
                         "peak_source": peak_kind, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": ldpc_ms,
                         "kernel_share_of_step": ldpc_ms * args.steps / elapsed_ms,
                         "edge_updates_per_s": F * 25 * 226799 / (ldpc_ms * 1e-3)},
            "roofline_issue": _issue_roofline(),
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            frames = 32 * threads * 2
            fps, ms, kind, threads = cpu_reference_run(frames, 3, 1, threads)
            out["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind,
                                   "sample": "%d frames x 3 passes of the same workload" % frames}
        emit(json.dumps(out))
    code.close()
    if world > 1:
        dist.destroy_process_group()


def mixed_batch(d, rank, frames_per_code):
    """This rank's shard of the mixed batch: code ids in a seeded random interleaving, LLRs back to back."""
    from dvbs2rx_b200 import vectors
    per = []
    for c, (std, fs, rate_name, esn0) in enumerate(MIXED_MODCODS):
        # 64 distinct frames per code, repeated: the encoders are numpy and the shard is large
        msg, cw, llr, info = vectors.make_llr_frames(std, fs, d.RATE[rate_name], 64, esn0, seed=7000 + 10 * rank + c)
        reps = (frames_per_code + 63) // 64
        per.append(dict(llr=np.tile(llr, (reps, 1))[:frames_per_code], msg=np.tile(msg, (reps, 1))[:frames_per_code], info=info, fs=fs))
    order = np.repeat(np.arange(len(per), dtype=np.uint8), frames_per_code)
    np.random.default_rng(99 + rank).shuffle(order)
    nxt = [0] * len(per)
    llrs, idx = [], []
    for c in order:
        llrs.append(per[c]["llr"][nxt[c]])
        idx.append(nxt[c])
        nxt[c] += 1
    return per, order, np.concatenate(llrs), np.array(idx)


def run_ours_mixed(args, torch, dist, d, world, rank, local, dev):
    from dvbs2rx_b200 import sharding
    # ---- tables of the five codes: built on rank 0, ONE broadcast of the concatenated blobs ----------
    def builder():
        blobs = [d.build_tables(std, fs, d.RATE[r]) for std, fs, r, _ in MIXED_MODCODS]
        head = np.array([len(blobs)] + [b.size for b in blobs], dtype=np.int64).view(np.uint8)
        return torch.from_numpy(np.concatenate([head] + blobs))
    cat = sharding.broadcast_tables(builder, dev).numpy()
    n = int(cat[:8].view(np.int64)[0])
    sizes = cat[8:8 + 8 * n].view(np.int64)
    off = 8 + 8 * n
    tables = []
    for s in sizes:
        tables.append(np.ascontiguousarray(cat[off:off + int(s)]))
        off += int(s)
    mixed = d.MixedCodes(tables=tables, device=local)

    # the global batch is world x 5 x MIXED_FRAMES_PER_CODE frames; this rank's shard by frame count
    per, order, llr_np, idx = mixed_batch(d, rank, MIXED_FRAMES_PER_CODE)
    F = order.size
    n_in, n_out = mixed.sizes(order)
    (lo, hi), (bi0, bi1), (bo0, bo1) = sharding.shard_mixed(np.tile(n_in, world), np.tile(n_out, world), rank, world)
    assert hi - lo == F and bi1 - bi0 == llr_np.size, "shard_mixed disagrees with the per-rank batch"
    out_total = int(n_out.sum())
    h_llr = torch.from_numpy(llr_np).pin_memory()
    d_llrs = [h_llr.to(dev, non_blocking=True), h_llr.to(dev, non_blocking=True).clone()]
    d_msg = torch.empty(out_total, dtype=torch.uint8, device=dev)
    d_tr = torch.empty(F, dtype=torch.int32, device=dev)
    d_co = torch.empty(F, dtype=torch.int32, device=dev)
    h_msg = torch.empty(out_total, dtype=torch.uint8).pin_memory()
    h_tr = torch.empty(F, dtype=torch.int32).pin_memory()
    h_co = torch.empty(F, dtype=torch.int32).pin_memory()
    stream = torch.cuda.current_stream().cuda_stream

    def step(k):
        mixed.fec_decode_dev(order, d_llrs[k & 1].data_ptr(), MAX_TRIALS, d_msg.data_ptr(), d_tr.data_ptr(), d_co.data_ptr(), stream)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    warm = max(args.warmup, 3)
    for k in range(warm):
        step(k)
    sync_all()
    launches0 = mixed.launches
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_load0 = time.perf_counter()
    t_start.record()
    for k in range(args.steps):
        step(k)
    t_end.record()
    sync_all()
    launches = mixed.launches - launches0
    el = torch.tensor([t_start.elapsed_time(t_end)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
    elapsed_ms = float(el.item())
    value = world * F * args.steps / (elapsed_ms * 1e-3)

    # ---- end to end through the host C ABI (pinned host buffers) ----
    def e2e_step():
        mixed.fec_decode_ptr(order, h_llr.data_ptr(), MAX_TRIALS, h_msg.data_ptr(), h_tr.data_ptr(), h_co.data_ptr())
    for _ in range(2):
        e2e_step()
    assert torch.equal(h_msg, d_msg.cpu()) and torch.equal(h_tr, d_tr.cpu()), "host-API result differs from the device-resident path"
    e2e_steps = max(3, min(args.steps, 10))
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2 = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2, op=dist.ReduceOp.MAX)
    e2e_value = world * F * e2e_steps / float(e2.item())
    sampler.window(t_load0, time.perf_counter())
    clocks = sampler.stop() if rank == 0 else None

    # ---- parity inside the run: a sample of this rank's shard against the per-code oracle (checker only,
    # outside every timed region); reference semantics = five single-code runs re-assembled by frame index
    oracle_lib = _oracle_lib()
    orc = oracle_lib.Oracle()
    msg_np, tr_np, co_np = h_msg.numpy(), h_tr.numpy(), h_co.numpy()
    out_off = np.concatenate([[0], np.cumsum(n_out)])
    mismatches, checked = 0, 0
    for c in range(len(per)):
        pos = np.where(order == c)[0][:args.check_per_code]
        info, fs = per[c]["info"], per[c]["fs"]
        for f in pos:
            llr_f = per[c]["llr"][idx[f]][None, :]
            o_post, o_ret = orc.ldpc_decode(info.table, llr_f, MAX_TRIALS)
            o_msg, o_corr = orc.bch_decode(orc.bch(fs, info.t, info.nbch), orc.pack_hard(o_post, info.nbch))
            got = msg_np[out_off[f]:out_off[f + 1]]
            ok = np.array_equal(got, o_msg[0]) and tr_np[f] == o_ret[0] and co_np[f] == o_corr[0]
            mismatches += 0 if ok else 1
            checked += 1
    # every frame that converged and was BCH-clean must equal what was sent
    sent_bad = 0
    for f in range(F):
        if tr_np[f] >= 0 and co_np[f] >= 0:
            c = order[f]
            if not np.array_equal(msg_np[out_off[f]:out_off[f + 1]], per[c]["msg"][idx[f]]):
                sent_bad += 1
    cnt = torch.tensor([mismatches, checked, sent_bad, int((tr_np >= 0).sum()), F], device=dev, dtype=torch.int64)
    sharding.allreduce_counters(cnt)
    mism, chk, sbad, conv, tot = [int(v) for v in cnt.tolist()]
    if mism or sbad:
        raise SystemExit("mixed workload: %d of %d sampled frames differ from the per-code oracle, %d decoded frames differ from what was sent"
                         % (mism, chk, sbad))
    if rank == 0:
        out = {
            "metric": "FECFRAMEs/s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int8", "data": "synthetic",
            "config": {"workload": MIXED_WORKLOAD, "frames_per_gpu_per_step": F, "max_trials": MAX_TRIALS,
                       "esn0_db": [m[3] for m in MIXED_MODCODS], "term": "per-frame",
                       "l2": "inputs larger than L2: %d MB of LLRs per GPU per step, two device copies used alternately" % (llr_np.size >> 20),
                       "sharding": "batch sharded by frame count (shard_mixed byte ranges), five tables broadcast once over NCCL"},
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": world * int(llr_np.size),
                    "d2h_bytes_per_step": world * (out_total + 8 * F), "buffers": "pinned host"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "parity": {"sampled_frames_vs_per_code_oracle": chk, "mismatches": mism, "converged_frames": conv, "frames": tot,
                       "converged_frames_differing_from_sent": sbad},
        }
        emit(json.dumps(out))
    mixed.close()
    if world > 1:
        dist.destroy_process_group()


_RESULT_FD = None


def emit(line):
    """The one JSON line goes to the process's real stdout; everything else written to fd 1 while the run was on
    (NCCL's version banner, a library's printf) went to stderr."""
    data = (line + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(line + "\n")
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    global _RESULT_FD
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c1", choices=["c1", "mixed"])
    ap.add_argument("--check-per-code", type=int, default=2, help="mixed: frames per code per rank checked against the oracle")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--device-only", action="store_true",
                    help="c1: skip the end-to-end legs (for a launch list under ncu: the host path overlaps copies with a "
                         "persistent kernel, which a profiler that serialises kernels stalls until the kernel's bounded wait gives up)")
    args = ap.parse_args()
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        # the checker libraries only (oracle/ and oracle/_ref); the product library is neither built nor loaded here
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
        if os.path.isdir("/root/reference"):
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
        run_reference(args)
    else:
        import __graft_entry__ as ge
        ge.build()
        run_ours(args)


if __name__ == "__main__":
    main()
