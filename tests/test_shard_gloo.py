"""CPU, world_size 2 over gloo: the N>1 host logic — contiguous read shards, the int64[3] count
all-reduce (the path's only collective) and ordered label gather.  The classifier on each rank is
the oracle (test infrastructure); on the GPU box bench.py runs the same logic over nccl."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ribodetector_b200 import shard
from ribodetector_b200.utils import synth


def test_shard_bounds_cover_and_order():
    for n in (0, 1, 7, 128, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (b0, e0), (b1, e1) in zip(spans, spans[1:]):
                assert e0 == b1 and b0 <= e0
            assert max(e - b for b, e in spans) == -(-n // world) or n == 0
    with pytest.raises(ValueError):
        shard.shard_bounds(10, 2, 2)


def test_shard_reads_rebases_offsets():
    seq, off = synth.synth_reads(101, 5, 30, 1)
    parts = [shard.shard_reads(seq, off, r, 4) for r in range(4)]
    assert sum(len(p[1]) - 1 for p in parts) == 101
    assert np.array_equal(np.concatenate([p[0] for p in parts]), seq)
    for s, o, (b, e) in parts:
        assert o[0] == 0 and o[-1] == len(s) and len(o) == e - b + 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.model_numpy import NumpyOracle
        from oracle import pairs
        from ribodetector_b200.utils.weights import load_weights
        seq, off = synth.synth_reads(n, 20, 60, 77)
        s, o, (b, e) = shard.shard_reads(seq, off, rank, world)
        reads = synth.to_strings(s, o)
        logits = NumpyOracle(load_weights(), np.float32).logits(reads, 50, "packed") if reads else np.zeros((0, 2))
        labels = pairs.argmax_labels(logits)
        counts = torch.from_numpy(pairs.counts(labels))
        shard.allreduce_counts(counts)
        gathered = shard.gather_labels(torch.from_numpy(labels), n, dst=0)
        q.put((rank, counts.tolist(), None if gathered is None else gathered.numpy().tolist(), (b, e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_counts_and_order_match_single_rank():
    from oracle.model_numpy import NumpyOracle
    from oracle import pairs
    from ribodetector_b200.utils.weights import load_weights
    n = 301                                  # odd: ranks get 151 + 150
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    seq, off = synth.synth_reads(n, 20, 60, 77)
    whole = pairs.argmax_labels(NumpyOracle(load_weights(), np.float32).logits(synth.to_strings(seq, off), 50, "packed"))
    assert res[0][1] == res[1][1] == pairs.counts(whole).tolist()       # both ranks hold the global counts
    assert res[0][2] == whole.tolist() and res[1][2] is None            # ordered gather on rank 0
    assert res[0][3] == (0, 151) and res[1][3] == (151, 301)
