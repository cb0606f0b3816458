"""The `ribodetector_cpu` batch loop as a torch-CPU stand-in (test infrastructure / the CPU arm
that bench.py times; see oracle/__init__.py).

Reference: ``ribodetector/detect_cpu.py``
  * ``-t`` forked worker processes, one intra-op thread each   :88-96, 171-187, 817
  * fixed batches of 1024 reads                                  :596
  * per batch: ``encode_variable_len_read`` → ``[B, L, 4]`` float32 → model → ``argmax``   :695-706
The reference runs the model through onnxruntime, which is not installed in this image (nor on
the GPU box, no network); the session is replaced by the module the ``.onnx`` was exported
from (``model_cpu.SeqModel``, restated in oracle/model_torch.py) with the same weights, i.e.
"ORT unavailable — torch-CPU stand-in for ribodetector_cpu" (BASELINE.md §4).
"""
import multiprocessing as mp
import os
import time

import numpy as np

BATCH = 1024
_STATE = {}


def _init(weights):
    os.environ["OMP_NUM_THREADS"] = "1"
    import torch
    torch.set_num_threads(1)
    from .model_torch import TorchOracle
    _STATE["oracle"] = TorchOracle(weights)


def _work(args):
    seq, off, max_len = args
    b = seq.tobytes()
    base = int(off[0])
    reads = [b[int(off[i]) - base:int(off[i + 1]) - base] for i in range(len(off) - 1)]
    return _STATE["oracle"].logits_padded(reads, max_len, batch=BATCH)


def classify(seq, off, max_len, weights, threads=None):
    """→ (labels int8[n], logits float32[n,2], seconds).  Batches of 1024 reads fanned out over `threads` forked
    single-threaded workers, like detect_cpu.py:283-298; argmax as in detect_cpu.py:705."""
    threads = threads or os.cpu_count() or 1
    n = len(off) - 1
    jobs = []
    for s in range(0, n, BATCH):
        e = min(n, s + BATCH)
        jobs.append((seq[int(off[s]):int(off[e])], off[s:e + 1], max_len))
    ctx = mp.get_context("fork")
    with ctx.Pool(threads, initializer=_init, initargs=(weights,)) as pool:
        pool.map(_work, jobs[:threads])               # warm the workers (model build, page-in)
        t0 = time.perf_counter()
        parts = pool.map(_work, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    logits = np.concatenate(parts) if parts else np.zeros((0, 2), np.float32)
    return np.argmax(logits, axis=1).astype(np.int8), logits, dt
