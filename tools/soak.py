"""Stability / determinism soak on the B200 box (one process per GPU; `torchrun --nproc-per-node N tools/soak.py
--minutes M` soaks N GPUs side by side):

  1. repeated create / classify / destroy with every precision (no leaks, identical bits),
  2. for `--minutes` minutes: full-size launches of every tensor-core precision over rotating batches; the logits of
     every launch are reduced ON THE DEVICE to a 64-bit checksum (sum of the raw fp32 bit patterns) and compared with
     the checksum the same batch and precision produced the first time — a bitwise-repeat check of every launch, which
     is what exercises the relaxed cluster-scope mbarrier protocol of the CTA-pair kernels (rd_lstm_tc.cu) for races.

python tools/soak.py [--minutes 10] [--reads 2097152]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ribodetector_b200.model import SeqModel                 # noqa: E402
from ribodetector_b200.utils import synth                    # noqa: E402
from ribodetector_b200.utils.weights import load_weights     # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--minutes", type=float, default=10.0)
ap.add_argument("--reads", type=int, default=1 << 21)
args = ap.parse_args()
rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(rank)
dev = "cuda:%d" % rank
w = load_weights()
PRECS = ("tc_mixed", "tc_exact", "tc_fast", "tc_auto", "tc_mixed_raw")

seq, off = synth.synth_reads(50000, 20, 150, 1)
ref = {}
t0 = time.time()
for i in range(24):
    p = (PRECS + ("fp32",))[i % 6]
    m = SeqModel(precision=p); m.load_state_dict(w); m.to(dev)
    r = m.classify_host(seq, off, 100)
    if p in ref:
        assert np.array_equal(ref[p], r["logits"].numpy()), p
    ref[p] = r["logits"].numpy().copy()
    m.close()
print("[gpu %d] 24 create/classify/destroy cycles ok in %.1f s; free/total %s" % (rank, time.time() - t0, torch.cuda.mem_get_info()), flush=True)

m = SeqModel(); m.load_state_dict(w); m.to(dev)
batches = []
for b, (lo, hi, L) in enumerate(((100, 100, 100), (150, 150, 150), (40, 300, 300))):
    s, o = synth.synth_reads_fixed(args.reads, L, 9 + b) if lo == hi else synth.synth_reads(args.reads // 2, lo, hi, 9 + b)
    batches.append((torch.from_numpy(s).to(dev), torch.from_numpy(o).to(dev), L))


def checksum(logits):
    return int(logits.view(torch.int32).to(torch.int64).sum().item())


first, launches, mismatches = {}, 0, 0
t0 = time.time()
deadline = t0 + 60.0 * args.minutes
report = t0 + 60.0
while time.time() < deadline:
    for b, (s, o, L) in enumerate(batches):
        for p in PRECS:
            cs = checksum(m.classify(s, o, L, precision=p)[0])
            launches += 1
            if first.setdefault((b, p), cs) != cs:
                mismatches += 1
                print("[gpu %d] MISMATCH batch %d precision %s launch %d" % (rank, b, p, launches), flush=True)
    if time.time() > report:
        print("[gpu %d] %5.1f min: %d launch sets, %d mismatches" % (rank, (time.time() - t0) / 60.0, launches, mismatches), flush=True)
        report += 60.0
print("[gpu %d] soak done: %.1f min, %d classify calls (%d distinct batch x precision), %d bitwise mismatches"
      % (rank, (time.time() - t0) / 60.0, launches, len(first), mismatches), flush=True)
m.close()
sys.exit(1 if mismatches else 0)
