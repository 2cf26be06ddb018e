#!/usr/bin/env python3
"""Summarise an ncu report (read here, no GPU needed) into profiles/<name>.md and, for the LDPC
kernel, profiles/ldpc_traffic.json (per-launch DRAM bytes that bench.py copies into `roofline.traffic`).

    python tools/summarize_ncu.py gpurun_out/ldpc_r01c.ncu-rep profiles/r01_ldpc_v21 [--traffic]
"""
import collections
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, outbase = sys.argv[1], sys.argv[2]
    rows = ncu_csv(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    name = m.get("Kernel Name", ("", "?"))[1]
    lines = ["# ncu summary: %s" % name, "", "source report: `%s` (ncu --set full --clock-control none)" % rep, "",
             "| metric | value | unit |", "|---|---|---|"]
    for k in KEYS:
        if k in m:
            lines.append("| %s | %s | %s |" % (k, m[k][1], m[k][0]))
    # SASS-level: executed instructions by opcode, and the top regions
    src = ncu_csv(rep, "source")
    body = [r for r in src[2:] if len(r) > 6]
    ops = collections.Counter()
    total = 0
    for r in body:
        try:
            n = int(r[5])
        except ValueError:
            continue
        parts = r[1].split()
        if not parts:
            continue
        op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
        ops[op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("VIMNMX", "VIADD", "LDS", "STS", "LDG", "STG")) and "." in op else "")] += n
        total += n
    if total:
        lines += ["", "## executed warp instructions by opcode (%.3f G total)" % (total / 1e9), "",
                  "| opcode | share |", "|---|---|"]
        for op, n in ops.most_common(22):
            lines.append("| %s | %.1f %% |" % (op, 100.0 * n / total))
        blk = [op for op in ops if op.startswith(("UTC", "LDTM", "STTM", "UBLKCP", "UTMA", "HMMA"))]
        lines += ["", "TMA / tensor SASS present: %s" % (", ".join(sorted(blk)) or "none")]
    with open(outbase + ".md", "w") as f:
        f.write("\n".join(lines) + "\n")
    if "--traffic" in sys.argv:
        def to_bytes(key):
            u, v = m[key]
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            return float(v) * scale
        t = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
        with open("profiles/ldpc_traffic.json", "w") as f:
            json.dump({"kernel": name, "report": rep, "dram_bytes_per_launch": t,
                       "dram_read": to_bytes("dram__bytes_read.sum"), "dram_write": to_bytes("dram__bytes_write.sum"),
                       "note": "one bench launch = 2664 frames x 68,850 B (soft input + packed hard decisions) = 183,416,400 algorithmic bytes"}, f, indent=1)
        # the issue-side roofline of the LDPC kernel (bench.py copies it into `roofline_issue`): the kernel is bound by
        # instruction issue / the ALU pipe, not by HBM
        def num(key):
            return float(m[key][1].replace(",", "")) if key in m else None
        inst = num("smsp__inst_executed.sum")
        edges = 2664 * 25 * 226799  # the bench launch: frames x iterations x edge updates of DVB-S2 1/2 normal
        with open("profiles/ldpc_issue.json", "w") as f:
            json.dump({"kernel": name, "report": rep, "bound": "issue",
                       "achieved": num("sm__inst_executed.avg.per_cycle_elapsed"), "peak": 4.0, "unit": "warp instructions / cycle / SM",
                       "frac": (num("sm__inst_executed.avg.per_cycle_elapsed") or 0) / 4.0,
                       "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                       "alu_pipe_pct": num("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                       "fma_pipe_pct": num("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                       "warp_instructions_per_launch": inst,
                       "thread_instructions_per_edge_update": inst * 32.0 / edges if inst else None,
                       "stall_barrier_per_issue": num("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
                       "stall_math_pipe_throttle_per_issue": num("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
                       "note": "ncu --set full of one bench-sized launch (2664 frames, 25 iterations); numbers under the profiler, "
                               "the frame rate is bench.py's"}, f, indent=1)
    print("\n".join(lines[:45]))


if __name__ == "__main__":
    main()
