#!/usr/bin/env python
"""A small run of the fp32 CUDA-core kernel (rd_lstm_fp32.cu) for compute-sanitizer: ragged reads at hidden sizes with
1, 2, 5 and 8 read subgroups per CTA (named barriers / __syncwarp), checked against the fp64 oracle.
    compute-sanitizer --tool racecheck python tests/fp32_sanity.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ribodetector_b200.model import SeqModel            # noqa: E402
from ribodetector_b200.utils import synth               # noqa: E402
from ribodetector_b200.utils.weights import load_weights  # noqa: E402
from oracle.model_numpy import NumpyOracle              # noqa: E402  (tests/ may use the oracle as the checker)

seq, off = synth.synth_reads(700, 1, 60, 99, n_frac=0.02)
reads = synth.to_strings(seq, off)
for H in (32, 64, 96, 128, 256):
    w = load_weights() if H == 128 else synth.synth_weights(H, 7)
    m = SeqModel(hidden_size=H, precision="fp32")
    m.load_state_dict(w)
    m.to("cuda:0")
    for sem in ("packed", "padded"):
        got = m.classify(seq, off, 50, semantics=sem)[0].cpu().numpy()
        d = np.abs(got - NumpyOracle(w, np.float64).logits(reads, 50, sem)).max()
        print("H = %3d %-6s max|dlogit| vs oracle = %.2e" % (H, sem, d), flush=True)
        assert d < 2e-4
    m.close()
print("ok")
