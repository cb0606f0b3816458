"""Read sharding across the GPUs of one box.

Reads (or read PAIRS — R1_i and R2_i stay together, detect.py:616-663) are independent units, so
the path shards with no data-path collective: rank g of G takes the contiguous range
[g*ceil(N/G), (g+1)*ceil(N/G)) which keeps "output order = input order" (detect.py:601-614) when the
per-rank label arrays are concatenated.  The only exchange is ONE all-reduce(sum) of the three
int64 label counters the reference logs (detect.py:210-242): 24 bytes over NCCL/NVLink (gloo on CPU
in the tests).  The reference has no counterpart (optional nn.DataParallel, detect.py:95-96)."""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n, rank, world):
    """Contiguous [begin, end) of units owned by `rank`; empty for trailing ranks when n < world."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world: %r/%r" % (rank, world))
    per = -(-int(n) // world)
    b = min(int(n), rank * per)
    return b, min(int(n), b + per)


def shard_reads(seq, off, rank, world):
    """Slice a (bytes, offsets) read set for `rank`; offsets are rebased to start at 0."""
    off = np.asarray(off)
    b, e = shard_bounds(len(off) - 1, rank, world)
    o = off[b:e + 1]
    return np.asarray(seq)[int(o[0]):int(o[-1])], (o - o[0]).astype(np.int64), (b, e)


def allreduce_counts(counts):
    """Sum {non-rRNA, rRNA, unclassified} over ranks, in place; a no-op without a process group.
    `counts` is an int64[3] tensor on the device the backend communicates from (cuda for nccl)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return counts


def gather_labels(labels, n_total, dst=0):
    """Concatenate per-rank label arrays (int8, contiguous shards in rank order) on rank `dst`.
    Convenience for writers that run on one rank; the bench path never calls it (labels leave each
    GPU by its own D2H copy)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return labels
    world, rank = dist.get_world_size(), dist.get_rank()
    per = -(-int(n_total) // world)
    pad = torch.zeros(per, dtype=torch.int8, device=labels.device)
    pad[:labels.numel()] = labels
    out = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, out, dst=dst)
    if rank != dst:
        return None
    return torch.cat(out)[:n_total]
