"""Multi-GPU plumbing: FECFRAMEs are independent, so a batch shards embarrassingly.

One process per GPU (torch.distributed; NCCL on GPUs, gloo in the CPU tests).  The only
collectives are at the edges of a run:
  * init: rank 0 builds the packed code tables once and broadcasts them (ONE broadcast), every
    rank creates its handle from the blob (dvbs2b200_code_create_from_tables);
  * end: frame / error / iteration counters are all-reduced for BER / FER reporting.
There is no data-path collective: rank r decodes frames [lo, hi) of the batch.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import build_tables


def build_tables_host(standard, framesize, rate):
    return torch.from_numpy(build_tables(standard, framesize, rate))


def broadcast_tables(builder, device):
    """rank 0 calls builder() -> uint8 tensor; everyone returns the same bytes (CPU uint8 tensor)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return builder()
    rank = dist.get_rank()
    if rank == 0:
        blob = builder().to(device)
        size = torch.tensor([blob.numel()], dtype=torch.int64, device=device)
    else:
        size = torch.zeros(1, dtype=torch.int64, device=device)
    dist.broadcast(size, 0)
    if rank != 0:
        blob = torch.empty(int(size.item()), dtype=torch.uint8, device=device)
    dist.broadcast(blob, 0)
    return blob.cpu()


def shard_range(frames, rank, world, multiple=32):
    """Contiguous shard [lo, hi) of a batch.  Shard boundaries fall on multiples of `multiple` so a
    batch-coupled termination group (16 / 32 frames) never straddles two GPUs."""
    units = (frames + multiple - 1) // multiple
    per, extra = divmod(units, world)
    lo_u = rank * per + min(rank, extra)
    hi_u = lo_u + per + (1 if rank < extra else 0)
    return min(lo_u * multiple, frames), min(hi_u * multiple, frames)


def allreduce_counters(t):
    """Sum a small int64 tensor of counters over the ranks (no-op for a single process)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def make_code(standard, framesize, rate, local_device):
    """Handle for this rank's GPU from tables built on rank 0 and broadcast once."""
    from . import Code
    dev = torch.device("cuda", local_device)
    blob = broadcast_tables(lambda: build_tables_host(standard, framesize, rate), dev)
    return Code(device=local_device, tables=np.ascontiguousarray(blob.numpy()))


def shard_mixed(frame_in_bytes, frame_out_bytes, rank, world):
    """Contiguous shard of a mixed-MODCOD batch: frames [lo, hi) and the byte ranges of their input and
    output in the back-to-back buffers.  Frames are split by count (they are independent); the byte
    offsets follow from the per-frame sizes."""
    n = len(frame_in_bytes)
    lo, hi = shard_range(n, rank, world, multiple=1)
    cin = np.concatenate([[0], np.cumsum(np.asarray(frame_in_bytes, dtype=np.int64))])
    cout = np.concatenate([[0], np.cumsum(np.asarray(frame_out_bytes, dtype=np.int64))])
    return (lo, hi), (int(cin[lo]), int(cin[hi])), (int(cout[lo]), int(cout[hi]))
