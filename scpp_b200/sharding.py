"""Host-side plumbing for multi-GPU runs: one process (rank) per GPU, the batch of problem instances is partitioned
contiguously; no trajectory or Jacobian data ever crosses GPUs.  The only exchange on the data path is the per-instance
convergence flags once per outer iteration (ncclAllGather inside libscpp_b200; the same protocol is expressed here over
torch.distributed so it can be exercised with the gloo backend on CPU).  bench.py and tools/multi_gpu_check.py bootstrap their ranks with
these helpers (unique-id broadcast, max / sum aggregation of the timings); tests/test_host.py runs them over gloo with world_size 2."""
import numpy as np


def shard_range(n_total, world, rank):
    """contiguous block partition; the first (n_total % world) ranks get one extra instance"""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_unique_id(dist, make_id, rank, src=0):
    """rank `src` creates the NCCL unique id (scpp_b200.comm_unique_id), every rank receives the same 128 bytes"""
    obj = [make_id() if rank == src else None]
    dist.broadcast_object_list(obj, src=src)
    return obj[0]


def global_active(dist, local_flags, pad_to):
    """all-gather of the per-instance flag bytes (0 = still iterating); returns (number of active instances over all
    ranks, gathered flags [world][pad_to]).  Shards are padded to a common length with 'converged' so unequal shards work."""
    import torch
    buf = torch.ones(pad_to, dtype=torch.uint8)
    buf[:len(local_flags)] = torch.as_tensor(np.asarray(local_flags, dtype=np.uint8))
    out = [torch.empty(pad_to, dtype=torch.uint8) for _ in range(dist.get_world_size())]
    dist.all_gather(out, buf)
    allf = torch.stack(out)
    return int((allf == 0).sum()), allf.numpy()


def reduce_timing(dist, seconds, counts, device=None):
    """max over ranks of the timed seconds, sum over ranks of the work counts (bench.py's aggregation); device: "cuda" under the nccl backend"""
    import torch
    t = torch.tensor(list(seconds), dtype=torch.float64, device=device)
    c = torch.tensor(list(counts), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return t.cpu().numpy(), c.cpu().numpy()
