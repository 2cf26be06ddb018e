"""Per-source-line totals of an ncu capture: joins `ncu --page source --csv` (SASS rows: executed instructions and
stall samples, no line numbers in the CSV) with `nvdisasm -g` line annotations of the same kernel, instruction by
instruction.

    python tools/ncu_by_line.py REPORT.ncu-rep OBJECT.o KERNEL_SUBSTRING [top]

The object must be the build the report was captured from (the instruction counts are checked).
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    cols = {n: i for i, n in enumerate(rows[hdr])}
    res = []
    for r in rows[hdr + 1:]:
        if len(r) < len(cols):
            continue
        res.append((r[cols["Source"]].strip(), int(r[cols["Instructions Executed"]] or 0), int(r[cols["# Samples"]] or 0)))
    return res


def line_map(obj, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    lines = txt.split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kernel in l)
    res = []
    cur = ("?", 0)
    for l in lines[start + 1:]:
        if l.startswith("//--------------------- .text") or l.startswith(".text."):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
        if m:
            inl = re.search(r'inlined at "([^"]+)", line (\d+)', m.group(3))
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
        if m:
            res.append((m.group(1).strip(), cur))
    return res


def main():
    rep, obj, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = sass_rows(rep)
    lm = line_map(obj, kernel)
    print("ncu rows %d, nvdisasm instructions %d" % (len(rows), len(lm)))
    n = min(len(rows), len(lm))
    mism = sum(1 for i in range(n) if rows[i][0].split()[0:1] != lm[i][0].replace("@", " @").split()[0:1] and rows[i][0].split()[-1:] != lm[i][0].split()[-1:])
    print("opcode mismatches in the join: %d" % mism)
    inst = collections.Counter()
    samp = collections.Counter()
    for i in range(n):
        inst[lm[i][1]] += rows[i][1]
        samp[lm[i][1]] += rows[i][2]
    ti, ts = sum(inst.values()), sum(samp.values())
    print("total warp instructions %.3f G, samples %d" % (ti / 1e9, ts))
    print("%-22s %6s %10s %7s %7s" % ("file", "line", "inst(M)", "inst%", "samp%"))
    for k, v in sorted(samp.items(), key=lambda kv: -kv[1])[:top]:
        print("%-22s %6d %10.1f %6.2f%% %6.2f%%" % (k[0], k[1], inst[k] / 1e6, 100.0 * inst[k] / ti, 100.0 * v / ts))
    byfile = collections.Counter()
    for k, v in samp.items():
        byfile[k[0]] += v
    print({k: "%.1f%%" % (100.0 * v / ts) for k, v in byfile.items()})


if __name__ == "__main__":
    main()
