"""The LDPC kernel's own thread functions on the CPU, against the oracle (no GPU needed).

gr-dvbs2rx_b200/csrc/ldpc_core.cuh and ldpc_steps.cuh are host/device code: tools/ldpc_emul.cc compiles them with g++
(intrinsics replaced by their scalar definitions) and runs the schedule of code_tables.cc thread by thread -- pair
steps, both forms of the split steps, the compressed check-node state -- on noisy codewords of every LDPC table, and
compares posteriors and return values with oracle/dvbs2_oracle.c bit for bit.  What this cannot see are races
between threads (compute-sanitizer on the GPU box does: profiles/r02_sanitizer_racecheck_synccheck.log)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emulator(tmp_path_factory):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    exe = str(tmp_path_factory.mktemp("emul") / "ldpc_emul")
    csrc = os.path.join(ROOT, "gr-dvbs2rx_b200", "csrc")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", csrc, os.path.join(ROOT, "tools", "ldpc_emul.cc"),
                           os.path.join(csrc, "code_tables.cc"), "-L" + os.path.join(ROOT, "oracle"), "-loracle",
                           "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-o", exe])
    return exe


def _run(exe, tables=(), env=None):
    e = dict(os.environ)
    e.pop("DVBS2B200_CHAIN", None)
    e.update(env or {})
    r = subprocess.run([exe] + [str(t) for t in tables], capture_output=True, text=True, env=e, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("table")]
    assert r.stdout.strip().endswith("all ok"), r.stdout[-2000:]
    return lines


def test_thread_functions_match_oracle_on_every_table(emulator):
    lines = _run(emulator)
    assert len(lines) == 57
    assert all(": ok" in l for l in lines)
    # the run must have exercised split steps (21 DVB-S2 / T2 tables have order-sensitive layers) and converging frames
    assert sum("conflict layers  0" not in l for l in lines) >= 20
    assert any("(1/2 frames converged)" in l or "(2/2 frames converged)" in l for l in lines)


def test_level_form_for_every_split_step(emulator):
    """DVBS2B200_CHAIN=0: the chain-form layers take the level form too (2-link nodes, W = 2 lanes on the GPU)."""
    lines = _run(emulator, tables=(3, 6, 10, 16, 19, 20), env={"DVBS2B200_CHAIN": "0", "EMUL_TRIALS": "8"})
    assert len(lines) == 6 and all(": ok" in l for l in lines)
