"""Multi-GPU host logic of the path: one process per GPU, every rank owns a
contiguous range of read ids (all windows of a read stay together, output order is kept) and
runs the whole path on it; the only exchange is the sum of the global counter vector
(len(TALLY_FIELDS) int64 values).  The reference's equivalent is multiprocessing.Pool over
shard files (elector/alignment.py:117-119) followed by computeStats' serial accumulation
(computeStats.py:519-675)."""
import numpy as np


def shard_reads(n_reads, rank, world):
    """[lo, hi) read ids of `rank`: contiguous, disjoint, covering, sizes differ by at most 1"""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    return (n_reads * rank) // world, (n_reads * (rank + 1)) // world


def shard_reads_balanced(cost_per_read, rank, world):
    """contiguous ranges balanced by a per-read cost (sum of window cells): rank r takes the reads
    whose cost prefix falls in [r, r+1) * total / world"""
    cost = np.asarray(cost_per_read, dtype=np.float64)
    pre = np.concatenate(([0.0], np.cumsum(cost)))
    total = pre[-1]
    if total <= 0:
        return shard_reads(len(cost), rank, world)
    cuts = np.searchsorted(pre, total * np.arange(world + 1) / world, side="left")
    cuts[0], cuts[-1] = 0, len(cost)
    cuts = np.maximum.accumulate(cuts)
    return int(cuts[rank]), int(cuts[rank + 1])


def slice_windows(d, read_lo, read_hi):
    """sub-workload holding reads [read_lo, read_hi) of the CSR workload d (arrays ref/cor/unc with
    *_off offsets per window, read_first = first window of each read)"""
    rf = d["read_first"]
    w0, w1 = int(rf[read_lo]), int(rf[read_hi])
    out = {}
    for k in ("ref", "cor", "unc"):
        off = d[k + "_off"]
        out[k] = d[k][int(off[w0]):int(off[w1])]
        out[k + "_off"] = off[w0:w1 + 1] - off[w0]
    out["read_first"] = rf[read_lo:read_hi + 1] - w0
    for k in d:
        if k not in out and not isinstance(d[k], np.ndarray):
            out[k] = d[k]
    return out


def reduce_counters(sums, group=None):
    """the one collective of the path: element-wise sum of the per-rank global counters over all ranks
    (NCCL for device tensors, gloo for host tensors); returns the tensor, reduced in place"""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


def gather_counters(per_read, group=None):
    """per-read counter matrices of all ranks, concatenated in rank (= read id) order on every rank"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return per_read
    world = dist.get_world_size(group)
    n = torch.tensor([per_read.shape[0]], dtype=torch.int64, device=per_read.device)
    ns = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(ns, n, group=group)
    mx = int(max(int(x) for x in ns))
    pad = torch.zeros((mx,) + tuple(per_read.shape[1:]), dtype=per_read.dtype, device=per_read.device)
    pad[:per_read.shape[0]] = per_read
    parts = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:int(k)] for p, k in zip(parts, ns)], dim=0)
