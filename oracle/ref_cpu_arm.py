"""The CPU arm of bench.py run on the REFERENCE'S OWN code (test infrastructure; see oracle/__init__.py).

When the unmodified reference package can be imported — from ``/root/reference`` (build container) or from
``baseline/_ref`` (``pip install --no-deps --ignore-requires-python --target baseline/_ref /root/reference``;
git-ignored, travels to the GPU box with the snapshot) — the ``ribodetector_cpu`` loop is timed with it:

  * ``-t`` forked worker processes with one intra-op thread each      ``detect_cpu.py:88-96,171-187,817``
  * fixed batches of 1024 reads                                         ``detect_cpu.py:596``
  * per batch ``np.array([SeqEncoder.encode_variable_len_read(read, max_len) ...], dtype=np.float32)`` →
    model → ``np.argmax``                                               ``detect_cpu.py:695-706``

The reference runs the model through ``onnxruntime.InferenceSession`` (``detect_cpu.py:88-96``); onnxruntime is
not installed in this image and cannot be (no network), so the session is replaced by the reference's own
``model_cpu.SeqModel`` — the module its ``.onnx`` was exported from (``convert_onnx.py:16,28-54``) — loaded with the
reference's own ``.pth``.  ``Bio.Seq`` (imported by ``seq_encoder.py:3`` for training helpers only) is stubbed.
Nothing of the reference is copied into this repository: the code runs from where the package is installed.
"""
import multiprocessing as mp
import os
import sys
import time
import types

import numpy as np

BATCH = 1024                     # detect_cpu.py:596
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_STATE = {}


def find_reference():
    """Directory to put on sys.path so that ``import ribodetector`` finds the unmodified reference, or None."""
    for root in ("/root/reference", os.path.join(_ROOT, "baseline", "_ref")):
        if os.path.isfile(os.path.join(root, "ribodetector", "model", "model_cpu.py")):
            return root
    return None


def _import_reference(root):
    if root not in sys.path:
        sys.path.insert(0, root)
    if "Bio.Seq" not in sys.modules:
        bio, bseq = types.ModuleType("Bio"), types.ModuleType("Bio.Seq")
        bseq.Seq = object
        bio.Seq = bseq
        sys.modules.setdefault("Bio", bio)
        sys.modules.setdefault("Bio.Seq", bseq)
    import torch
    from ribodetector.data_loader import seq_encoder as SeqEncoder
    from ribodetector.model import model_cpu
    import json
    pkg = os.path.join(root, "ribodetector")
    with open(os.path.join(pkg, "config.json")) as f:
        cfg = json.load(f)
    args = dict(cfg["arch"]["args"])
    args.pop("pack_seq", None)
    model = model_cpu.SeqModel(**args)
    state = torch.load(os.path.join(pkg, cfg["state_file"]["mcc"]), map_location="cpu")
    model.load_state_dict(state["state_dict"])
    model.eval()
    return torch, SeqEncoder, model


def _init(root):
    # one intra-op thread per worker (detect_cpu.py:88-90) — set BEFORE the first torch op of the forked child: the
    # parent's OpenMP pool does not survive fork(), and a parallel region entered in the child would wait for it
    os.environ["OMP_NUM_THREADS"] = "1"
    import torch
    torch.set_num_threads(1)
    torch, enc, model = _import_reference(root)
    _STATE.update(torch=torch, enc=enc, model=model)


def _batch_logits(reads, max_len):
    torch, enc, model = _STATE["torch"], _STATE["enc"], _STATE["model"]
    x = np.array([enc.encode_variable_len_read(r, max_len=max_len) for r in reads], dtype=np.float32)
    with torch.no_grad():
        return model(torch.from_numpy(x)).numpy()


def _split(seq, off):
    b = seq.tobytes().decode("latin-1")
    base = int(off[0])
    return [b[int(off[i]) - base:int(off[i + 1]) - base] for i in range(len(off) - 1)]


def _work(args):
    seq, off, max_len = args
    return _batch_logits(_split(seq, off), max_len)


def classify(seq, off, max_len, threads=None, root=None):
    """→ (labels int8[n], logits float32[n,2], seconds): batches of 1024 reads over `threads` forked single-threaded
    workers running the reference's encoder and model."""
    root = root or find_reference()
    if root is None:
        raise RuntimeError("the reference package is not importable (neither /root/reference nor baseline/_ref)")
    threads = threads or os.cpu_count() or 1
    n = len(off) - 1
    jobs = []
    for s in range(0, n, BATCH):
        e = min(n, s + BATCH)
        jobs.append((seq[int(off[s]):int(off[e])], off[s:e + 1], max_len))
    ctx = mp.get_context("fork")
    with ctx.Pool(threads, initializer=_init, initargs=(root,)) as pool:
        pool.map(_work, jobs[:threads])               # warm the workers (imports, model build, page-in)
        t0 = time.perf_counter()
        parts = pool.map(_work, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    logits = np.concatenate(parts) if parts else np.zeros((0, 2), np.float32)
    return np.argmax(logits, axis=1).astype(np.int8), logits, dt


def _single_core(args):
    seq, off, max_len, reps = args
    torch, enc, model = _STATE["torch"], _STATE["enc"], _STATE["model"]
    reads = _split(seq, off)
    t0 = time.perf_counter()
    for _ in range(reps):
        x = np.array([enc.encode_variable_len_read(r, max_len=max_len) for r in reads], dtype=np.float32)
    t1 = time.perf_counter()
    xt = torch.from_numpy(x)
    with torch.no_grad():
        model(xt)
        t2 = time.perf_counter()
        for _ in range(reps):
            np.argmax(model(xt).numpy(), axis=1)
    t3 = time.perf_counter()
    n = len(reads) * reps
    return n / (t1 - t0), n / (t3 - t2)


def single_core_split(seq, off, max_len, root=None, reps=2):
    """Encode-only and model-only reads/s of ONE worker (one core, one thread) on one batch of 1024 reads."""
    root = root or find_reference()
    n = min(BATCH, len(off) - 1)
    ctx = mp.get_context("fork")
    with ctx.Pool(1, initializer=_init, initargs=(root,)) as pool:
        enc_rate, model_rate = pool.apply(_single_core, ((seq[int(off[0]):int(off[n])], off[:n + 1], max_len, reps),))
    return {"encode_only_reads_per_s_one_core": enc_rate, "model_only_reads_per_s_one_core": model_rate,
            "reads_per_s_one_core": 1.0 / (1.0 / enc_rate + 1.0 / model_rate)}
