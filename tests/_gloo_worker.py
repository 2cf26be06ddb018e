"""Worker for test_sharding_under_gloo_world_size_2 (launched by torchrun, CPU, gloo)."""
import hashlib
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gr-dvbs2rx_b200"))
from dvbs2rx_b200 import sharding  # noqa: E402

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
blob = sharding.broadcast_tables(lambda: sharding.build_tables_host(0, 1, 3), device="cpu")
lo, hi = sharding.shard_range(100, rank, world, multiple=1)
total = sharding.allreduce_counters(torch.tensor([hi - lo], dtype=torch.int64))
with open(os.path.join(sys.argv[1], "rank%d.txt" % rank), "w") as f:
    f.write("%s %d:%d %d\n" % (hashlib.sha256(blob.numpy().tobytes()).hexdigest(), lo, hi, int(total[0])))
dist.destroy_process_group()
