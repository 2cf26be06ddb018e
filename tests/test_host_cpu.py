"""CPU tests of the host side: code lookup, schedule, table blob, C-ABI surface, vector generator,
and the N > 1 sharding logic under gloo.  No compute call is made (no GPU here)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built):
    import dvbs2rx_b200 as d
    hdr = open(os.path.join(ROOT, "include", "dvbs2_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(dvbs2b200_\w+)\s*\(", hdr)))
    assert declared == sorted(d.EXPORTED_SYMBOLS)
    lib = C.CDLL(d.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert d.lib().dvbs2b200_version() == 100


def test_lookup_matches_reference_parameters(built):
    """lib/fec_params.cc / lib/ldpc_decoder_bb_impl.cc:104-307 spot checks (Appendix C of SURVEY.md)."""
    import dvbs2rx_b200 as d
    exp = {
        ("C1_2", 1): ("DVB_S2_TABLE_B4", 64800, 32400, 90, 32208, 32400, 12, 16),
        ("C3_4", 1): ("DVB_S2_TABLE_B7", 64800, 48600, 45, 48408, 48600, 12, 16),
        ("C3_5", 1): ("DVB_S2_TABLE_B5", 64800, 38880, 72, 38688, 38880, 12, 16),
        ("C2_3", 0): ("DVB_S2_TABLE_C6", 16200, 10800, 15, 10632, 10800, 12, 14),
        ("C9_10", 1): ("DVB_S2_TABLE_B11", 64800, 58320, 18, 58192, 58320, 8, 16),
    }
    for (rate, fs), (name, n, k, q, kb, nb, t, m) in exp.items():
        i = d.lookup(d.STANDARD_DVBS2, fs, d.RATE[rate])
        assert d.lib().dvbs2b200_table_name(i.table).decode() == name
        assert (i.n_ldpc, i.k_ldpc, i.q, i.kbch, i.nbch, i.t, i.gf_m) == (n, k, q, kb, nb, t, m)
    # the standard selects the T2 table only for 2/3 normal and 3/5 short
    assert d.lib().dvbs2b200_table_name(d.lookup(d.STANDARD_DVBT2, 1, d.C2_3).table) == b"DVB_T2_TABLE_A3"
    assert d.lib().dvbs2b200_table_name(d.lookup(d.STANDARD_DVBT2, 0, d.C3_5).table) == b"DVB_T2_TABLE_B3"
    with pytest.raises(d.Dvbs2Error) as e:
        d.lookup(0, 1, d.C7_8)  # not a DVB-S2 rate (the reference leaves d_ldpc unset)
    assert e.value.code == d.EUNSUPPORTED


def test_circulants_reproduce_links_total(built):
    import dvbs2rx_b200 as d
    for table in range(d.lib().dvbs2b200_num_tables()):
        layer, group, shift = d.table_circulants(table)
        assert (shift < 360).all() and (np.diff(layer) >= 0).all()
        # one circulant = 360 edges; a (layer, group) pair may repeat (conflict layers) but never a triple
        assert len(set(zip(layer.tolist(), group.tolist(), shift.tolist()))) == len(layer)


def test_schedule_statistics_match_survey(built):
    """SURVEY.md Appendix C [probe]: wavefront steps per iteration / deepest layer / conflict layers."""
    import dvbs2rx_b200 as d
    exp = {("C1_2", 1): (162, 33, 8), ("C3_4", 1): (349, 180, 20), ("C3_5", 1): (293, 36, 30),
           ("C2_3", 0): (160, 52, 11), ("C9_10", 1): (197, 90, 18)}
    for (rate, fs), (steps, depth, conf) in exp.items():
        s = d.schedule_stats(d.lookup(0, fs, d.RATE[rate]).table)
        assert (s["steps_per_iter"], s["max_depth"], s["conflict_layers"]) == (steps, depth, conf)


def test_wavefront_schedule_preserves_serial_order(built):
    """Replay the device schedule on the host: every data bit must be touched by check nodes in the
    same relative order as the reference's serial j loop (lib/ldpc_decoder/layered_decoder.hh:50-79)."""
    import dvbs2rx_b200 as d
    for rate, fs in (("C1_2", 1), ("C9_10", 1), ("C2_3", 0), ("C3_4", 1)):
        table = d.lookup(0, fs, d.RATE[rate]).table
        layer, group, shift = d.table_circulants(table)
        for i in np.unique(layer):
            circ = [(g, a) for l, g, a in zip(layer, group, shift) if l == i]
            if len({g for g, _ in circ}) == len(circ):
                continue
            last = {}
            level = np.zeros(360, dtype=int)
            for j in range(360):
                bits = [(g, (j - a) % 360) for g, a in circ]
                level[j] = 1 + max([last.get(b, 0) for b in bits])
                for b in bits:
                    last[b] = level[j]
            # same-level check nodes never share a bit; lower j never sits at a higher level than a
            # later node that shares one of its bits
            for lv in range(1, level.max() + 1):
                seen = set()
                for j in np.nonzero(level == lv)[0]:
                    bits = {(g, (j - a) % 360) for g, a in circ}
                    assert not (bits & seen)
                    seen |= bits


def test_compute_calls_fail_loudly_without_a_gpu(built):
    import dvbs2rx_b200 as d
    if d.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(d.Dvbs2Error) as e:
        d.Code(d.STANDARD_DVBS2, d.FECFRAME_NORMAL, d.C1_2)
    assert e.value.code == d.ECUDA and "no CPU fallback" in str(e.value)
    # the other handle types and the page-locking helper fail the same way: an error code, no crash, no fallback
    for make in (lambda: d.MixedCodes([(0, 1, d.C1_2), (0, 0, d.C2_3)]), lambda: d.MultiCode([0], 0, 1, d.C1_2),
                 lambda: d.PlDescrambler(0, 0), lambda: d.host_register(np.zeros(4096, dtype=np.uint8))):
        with pytest.raises(d.Dvbs2Error) as e:
            make()
        assert e.value.code == d.ECUDA


def test_vector_generator_roundtrips_through_the_oracle(oracle):
    import dvbs2rx_b200 as d
    from dvbs2rx_b200 import vectors
    msg, cw, llr, info = vectors.make_llr_frames(0, 0, d.C2_3, 3, 6.0, seed=9)
    post, ret = oracle.ldpc_decode(info.table, llr, 25)
    hard = oracle.pack_hard(post, info.nbch)
    out, corr = oracle.bch_decode(oracle.bch(0, info.t, info.nbch), hard)
    assert (ret >= 0).all() and (corr >= 0).all()
    assert np.array_equal(out, msg)
    # 8PSK mapper / interleaver against the oracle demapper's hard decisions (no noise)
    msg, cw, info = vectors.encode_frames(0, 1, d.C3_5, 1, np.random.default_rng(2))
    iq = vectors.map_symbols(cw, d.MOD_8PSK, d.C3_5)
    llr = oracle.demap_8psk(iq, 0.1, d.C3_5)
    assert np.array_equal((llr < 0).astype(np.uint8), cw)
    iq = vectors.map_symbols(cw, d.MOD_QPSK, d.C3_5)
    assert np.array_equal((oracle.demap_qpsk(iq, 0.5) < 0).astype(np.uint8), cw)


def test_sharding_under_gloo_world_size_2(built, tmp_path):
    """N > 1 host logic on CPU: contiguous frame shards, table blob broadcast from rank 0, counters
    all-reduced -- the same code path bench.py/run_sharded uses with NCCL."""
    script = os.path.join(ROOT, "tests", "_gloo_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29541")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29541", script, str(tmp_path)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    outs = sorted(os.listdir(tmp_path))
    assert outs == ["rank0.txt", "rank1.txt"]
    a, b = (open(os.path.join(tmp_path, o)).read().split() for o in outs)
    assert a[0] == b[0]            # same table blob digest on both ranks
    assert (a[1], b[1]) == ("0:50", "50:100")  # contiguous shards of 100 frames
    assert a[2] == b[2] == "100"   # all-reduced frame counter


def _split_steps(blob):
    """(layer, delta, depth, flags, work offset) of the split steps of a table blob, and its work[] array."""
    hdr = np.frombuffer(blob[:160].tobytes(), dtype=np.uint32)
    n_steps, step_off, order_off = int(hdr[22]), int(hdr[32]), int(hdr[33])
    raw = blob[step_off:step_off + 8 * n_steps].tobytes()
    steps = np.frombuffer(raw, dtype=np.dtype([("layer", "u1"), ("run_len", "u1"), ("count", "<u2"), ("work_off", "<u4")]))
    work = np.frombuffer(blob[order_off:int(hdr[34])].tobytes(), dtype=np.uint16)
    return [(int(s["layer"]), int(s["run_len"]), int(s["count"]), int(s["work_off"]) >> 24, int(s["work_off"]) & 0xffffff)
            for s in steps if s["count"]], work


def test_split_schedule_covers_every_table(built):
    """Every conflict layer of the 57 tables has at most kMaxSharedLinks (12) shared links.  The level table of a
    split step: levels start at 1, never fall with j (the kernel hands out a level as a RANGE of nodes), and
    first_node[] marks exactly those ranges.  Holds with and without the chain form (DVBS2B200_CHAIN=0)."""
    import dvbs2rx_b200 as d
    from collections import Counter
    worst = 0
    for t in range(d.lib().dvbs2b200_num_tables()):
        lay, grp, sh = d.table_circulants(t)
        for layer in range(int(lay.max()) + 1):
            c = Counter(grp[lay == layer].tolist())
            worst = max(worst, sum(v for v in c.values() if v > 1))
    assert worst == 12
    old = os.environ.get("DVBS2B200_CHAIN")
    try:
        for mode in (None, "0"):
            if mode is None:
                os.environ.pop("DVBS2B200_CHAIN", None)
            else:
                os.environ["DVBS2B200_CHAIN"] = mode
            seen_chain = False
            for fs, rate in ((1, "C1_2"), (1, "C3_4"), (1, "C9_10"), (0, "C8_9"), (0, "C2_3")):
                blob = d.build_tables(0, fs, d.RATE[rate])
                steps, work = _split_steps(blob)
                assert steps
                for layer, delta, depth, flags, off in steps:
                    seen_chain |= bool(flags & 0x80)
                    level = work[off:off + 360].astype(int)
                    first = work[off + 360:off + 360 + depth + 2].astype(int)
                    assert level[0] == 1 and level.max() == depth and (np.diff(level) >= 0).all(), (rate, layer)
                    assert first[depth + 1] == 360
                    for lv in range(1, depth + 1):
                        assert (level[first[lv]:first[lv + 1]] == lv).all() and first[lv] < first[lv + 1], (rate, layer, lv)
            assert seen_chain == (mode is None)
    finally:
        if old is None:
            os.environ.pop("DVBS2B200_CHAIN", None)
        else:
            os.environ["DVBS2B200_CHAIN"] = old


def test_mixed_batch_sharding(built):
    """Mixed-MODCOD batches shard by frame count; byte ranges follow the per-frame sizes and tile the buffers."""
    from dvbs2rx_b200 import sharding
    rng = np.random.default_rng(2)
    sizes_in = rng.choice([16200, 64800], size=37)
    sizes_out = np.where(sizes_in == 16200, 1323, 4026)
    for world in (1, 2, 3, 8):
        prev_hi, prev_in, prev_out = 0, 0, 0
        for r in range(world):
            (lo, hi), (i0, i1), (o0, o1) = sharding.shard_mixed(sizes_in, sizes_out, r, world)
            assert lo == prev_hi and i0 == prev_in and o0 == prev_out
            assert i1 - i0 == sizes_in[lo:hi].sum() and o1 - o0 == sizes_out[lo:hi].sum()
            prev_hi, prev_in, prev_out = hi, i1, o1
        assert prev_hi == 37 and prev_in == sizes_in.sum() and prev_out == sizes_out.sum()


def test_apsk_tables_and_model(built):
    """dvbs2rx_b200.apsk: unit-energy, distinct points; the float64 max-log model reduces to the QPSK formula."""
    from dvbs2rx_b200 import apsk, vectors
    for pts in [apsk.points_16apsk(g) for g in apsk.GAMMA_16APSK.values()] + \
               [apsk.points_32apsk(*g) for g in apsk.GAMMA_32APSK.values()]:
        assert abs(float((pts.astype(np.float64) ** 2).sum(axis=1).mean()) - 1.0) < 1e-6
        assert len({(round(float(x), 5), round(float(y), 5)) for x, y in pts}) == pts.shape[0]
    assert apsk.row_offsets(64800, 4).tolist() == [0, 16200, 32400, 48600]
    assert apsk.row_offsets(64800, 3, (2, 1, 0)).tolist() == [43200, 21600, 0]
    rng = np.random.default_rng(1)
    a = np.float32(np.sqrt(0.5))
    qpsk = np.array([[a, a], [a, -a], [-a, a], [-a, -a]], dtype=np.float32)
    bits = rng.integers(0, 2, size=(2, 64), dtype=np.uint8)
    offs = apsk.row_offsets(64, 2)
    iq = apsk.map_bits(bits, qpsk, offs) + rng.normal(0, 0.3, size=(2, 32, 2)).astype(np.float32)
    n0 = 0.18
    model = apsk.maxlog_llr(iq, qpsk, offs, n0)
    want = np.concatenate([iq[:, :, 0], iq[:, :, 1]], axis=1).astype(np.float64) * (2 * np.sqrt(2.0) / n0)
    assert np.allclose(model, want, rtol=1e-5, atol=1e-5)
