#!/usr/bin/env python3
"""bench.py -- FECFRAMEs/s of the DVB-S2 FEC decode hot path (LDPC -> BCH) on B200.

Workload (BASELINE.json configs[0], the configuration the metric is quoted on): QPSK 1/2 normal
FECFRAMEs (64800 bit), 25 offset-min-sum iterations, AWGN Es/N0 = 1.0 dB (no frame converges, so
every frame runs all 25 iterations and takes the BCH failure path), int8 LLR input, decoded
BBFRAME bytes out.  A "step" is one pass of LDPC+BCH over one batch of frames per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU); frames shard across ranks with no data-path
collective (weak scaling: per-GPU batch fixed); the code tables are built on rank 0 and
broadcast once over NCCL.  --impl reference times the reference's own CPU implementation
(oracle/_ref: its unmodified translation units, AVX2 SIMD, one decoder per host thread).
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
        sm, mx, reasons = [], [], set()
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
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(frames, seed):
    import dvbs2rx_b200 as d
    from dvbs2rx_b200 import vectors
    msg, cw, llr, info = vectors.make_llr_frames(d.STANDARD_DVBS2, d.FECFRAME_NORMAL, d.C1_2, frames, ESN0_DB, seed)
    return msg, llr, info


# --------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own CPU implementation (oracle/_ref)
# --------------------------------------------------------------------------------------------
def cpu_reference_run(frames, steps, warmup, threads):
    """LDPC (AVX2, 32 frames per SIMD batch, one decoder instance per thread) + BCH on the
    resulting bytes, exactly the reference's translation units.  Returns frames/s, ms/step."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    import dvbs2rx_b200 as d
    kind = "reference" if os.path.exists(oracle_lib.REF_PATH) else "port"
    msg, llr, info = make_inputs(frames, seed=1234)
    times = []
    if kind == "reference":
        ref = oracle_lib.Ref()
        name = d.lib().dvbs2b200_table_name(info.table).decode()
        bch = ref.bch(d.FECFRAME_NORMAL, info.t, info.nbch)
        orc = oracle_lib.Oracle()
        for it in range(warmup + steps):
            post, ret, t_ldpc = ref.ldpc_decode_mt(name, llr, MAX_TRIALS, threads)
            hard = np.ascontiguousarray(np.packbits(post[:, :info.nbch] < 0, axis=1))  # untimed glue
            out, corr, t_bch = ref.bch_decode_mt(bch, hard, info.nbch, threads)
            if it >= warmup:
                times.append(t_ldpc + t_bch)
    else:
        orc = oracle_lib.Oracle()
        threads = 1
        bch = orc.bch(d.FECFRAME_NORMAL, info.t, info.nbch)
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            post, ret = orc.ldpc_decode(info.table, llr, MAX_TRIALS, lanes=32)
            out, corr = orc.bch_decode(bch, orc.pack_hard(post, info.nbch))
            if it >= warmup:
                times.append(time.perf_counter() - t0)
    total = float(np.sum(times))
    return frames * len(times) / total, 1e3 * total / len(times), kind, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    frames = 32 * threads * 2  # bounded sample: two SIMD batches per host thread per step
    fps, ms, kind, threads = cpu_reference_run(frames, args.steps, args.warmup, threads)
    sample = "%d frames per step (32-frame AVX2 batches), %d host threads" % (frames, threads)
    print(json.dumps({
        "impl": "reference", "metric": "FECFRAMEs/s", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": frames, "max_trials": MAX_TRIALS,
                   "esn0_db": ESN0_DB, "term": "reference SIMD batch of 32"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def cpu_reference_ldpc_rate(table_name, llr, trials, threads=None):
    """Frames/s of the reference's own AVX2 LDPC decoder (oracle/_ref) on `llr`, one decoder instance per host
    thread; None where oracle/_ref is not built.  The one CPU leg other measurement tools (tools/sweep_configs.py)
    go through, so that nothing outside tests/, smoke() and this file touches oracle/."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
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

    # ---- code tables: built on rank 0, ONE NCCL broadcast, every rank creates from the blob ----
    from dvbs2rx_b200 import sharding
    code = sharding.make_code(d.STANDARD_DVBS2, d.FECFRAME_NORMAL, d.C1_2, local)
    info = code.info
    F = FRAMES_PER_GPU
    N, nb, kb = info.n_ldpc, info.nbch // 8, info.kbch // 8

    # ---- synthetic input: every rank its own frames (shard = contiguous frame range) ------------
    msg, llr_np, _ = make_inputs(F, seed=1000 + rank)
    h_llr = torch.from_numpy(llr_np).pin_memory()
    d_llr = h_llr.to(dev, non_blocking=True)
    d_mid = torch.empty((F, nb), dtype=torch.uint8, device=dev)
    d_msg = torch.empty((F, kb), dtype=torch.uint8, device=dev)
    d_trials = torch.empty(F, dtype=torch.int32, device=dev)
    d_corr = torch.empty(F, dtype=torch.int32, device=dev)
    h_msg = torch.empty((F, kb), dtype=torch.uint8).pin_memory()
    h_trials = torch.empty(F, dtype=torch.int32).pin_memory()
    h_corr = torch.empty(F, dtype=torch.int32).pin_memory()
    l2_flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step(ev=None):
        if ev:
            ev[0].record()
        code.ldpc_decode_dev(d_llr.data_ptr(), F, MAX_TRIALS, d.TERM_PER_FRAME, d.OM_MESSAGE, d_mid.data_ptr(),
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
    for _ in range(max(args.warmup, 3)):
        step()
    sync_all()
    launches0 = code.launches
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    t_load0 = time.perf_counter()
    t_start.record()
    for k in range(args.steps):
        step(evs[k])
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

    # ---- end to end through the host C ABI: pinned host buffers, H2D + D2H inside the timed region ----
    def e2e_step():
        code.fec_decode_ptr(d.MOD_QPSK, None, None, h_llr.data_ptr(), F, MAX_TRIALS, d.TERM_PER_FRAME,
                            h_msg.data_ptr(), h_trials.data_ptr(), h_corr.data_ptr())
    for _ in range(2):
        e2e_step()
    assert torch.equal(h_msg, d_msg.cpu()), "host-API result differs from the device-resident path"
    e2e_steps = max(3, min(args.steps, 10))
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2 = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2, op=dist.ReduceOp.MAX)
    e2e_value = world * F * e2e_steps / float(e2.item())
    # clocks / throttle reasons sampled while the GPU was busy: the timed region and the end-to-end loop after it
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
            "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_gpu_per_step": F, "max_trials": MAX_TRIALS,
                       "esn0_db": ESN0_DB, "term": "per-frame", "l2": "inputs larger than L2 (173 MB per GPU per step)",
                       "sharding": "frames sharded across ranks, tables broadcast once over NCCL"},
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": world * F * N,
                    "d2h_bytes_per_step": world * F * (kb + 8)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "ldpc_decode_kernel", "achieved": achieved, "peak": peak,
                         "peak_source": peak_kind, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": ldpc_ms,
                         "kernel_share_of_step": ldpc_ms * args.steps / elapsed_ms,
                         "edge_updates_per_s": F * 25 * 226799 / (ldpc_ms * 1e-3)},
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            frames = 32 * threads * 2
            fps, ms, kind, threads = cpu_reference_run(frames, 3, 1, threads)
            out["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind,
                                   "sample": "%d frames x 3 passes of the same workload" % frames}
        print(json.dumps(out))
    del l2_flush
    code.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    import __graft_entry__ as ge
    ge.build()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
