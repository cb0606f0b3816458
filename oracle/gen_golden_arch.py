#!/usr/bin/env python
"""Golden logits of the REAL reference model at hidden sizes other than the shipped 128 (test infrastructure; build
container only — the GPU box has no /root/reference; the fixture it writes is committed).

The reference builds its model as ``SeqModel(**config['arch']['args'])`` (parse_config.py:43-57, detect.py:93), so any
``hidden_size`` is a legal configuration (model/model.py:11-29) although one checkpoint ships.  For each H this script
instantiates the unmodified ``ribodetector.model.model.SeqModel`` (packed → forward1, model.py:32-37) and
``ribodetector.model.model_cpu.SeqModel`` (padded → forward_last, model_cpu.py:29-37), loads the seeded weights of
``ribodetector_b200.utils.synth.synth_weights(H, seed)`` (regenerated, not stored: PCG64 is stable), runs them on
seeded reads incl. the edge cases of gen_golden.py, pins both oracle restatements against the outputs and writes
tests/golden/arch.npz.     Usage: python oracle/gen_golden_arch.py [--ref /root/reference]
"""
import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import encoders                              # noqa: E402
from oracle.gen_golden import (import_reference, edge_reads, ref_logits_packed, ref_logits_padded)  # noqa: E402
from oracle.model_numpy import NumpyOracle              # noqa: E402
from oracle.model_torch import TorchOracle              # noqa: E402
from ribodetector_b200.utils import synth               # noqa: E402

HIDDEN_SIZES = (32, 64, 96, 192, 256)
WEIGHT_SEED = synth.SEED_BASE + 70
MAX_LEN = 100


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "arch.npz"))
    args = ap.parse_args()
    torch.manual_seed(0)
    torch.set_num_threads(1)
    ref_model, ref_model_cpu, ref_enc, ref_detect, _cfg = import_reference(args.ref)
    rng = np.random.Generator(np.random.PCG64(synth.SEED_BASE + 71))
    s, o = synth.synth_reads(96, 1, 130, synth.SEED_BASE + 72, n_frac=0.01)
    reads = [r for r in edge_reads(rng, MAX_LEN) + synth.to_strings(s, o) if len(r) > 0]
    seq, off = encoders.flatten_reads(reads)
    out = dict(seq=seq, off=off, max_len=np.int64(MAX_LEN), hidden_sizes=np.asarray(HIDDEN_SIZES, np.int64),
               weight_seed=np.int64(WEIGHT_SEED))
    print("%5s %6s %10s %10s %10s %10s" % ("H", "n", "torch/pk", "torch/pad", "np64/pk", "np64/pad"))
    for H in HIDDEN_SIZES:
        w = synth.synth_weights(H, WEIGHT_SEED)
        sd = {k: torch.from_numpy(v) for k, v in w.items()}
        kw = dict(input_size=4, hidden_size=H, num_layers=1, num_classes=2, batch_first=True, bidirectional=True)
        m_packed = ref_model.SeqModel(pack_seq=True, **kw)
        m_packed.load_state_dict(sd)
        m_packed.eval()
        m_padded = ref_model_cpu.SeqModel(pack_seq=False, **kw)
        m_padded.load_state_dict(sd)
        m_padded.eval()
        lp = ref_logits_packed(ref_detect, m_packed, reads, MAX_LEN)
        ld = ref_logits_padded(ref_enc, m_padded, reads, MAX_LEN)
        o_t, o_n = TorchOracle(w, hidden_size=H), NumpyOracle(w, np.float64)
        d = (np.abs(o_t.logits_packed(reads, MAX_LEN) - lp).max(), np.abs(o_t.logits_padded(reads, MAX_LEN) - ld).max(),
             np.abs(o_n.logits(reads, MAX_LEN, "packed") - lp).max(), np.abs(o_n.logits(reads, MAX_LEN, "padded") - ld).max())
        print("%5d %6d %10.2e %10.2e %10.2e %10.2e" % ((H, len(reads)) + d))
        assert d[0] <= 2e-6 and d[1] <= 2e-6 and d[2] <= 5e-5 and d[3] <= 5e-5, (H, d)
        out["logits_packed_h%d" % H] = lp
        out["logits_padded_h%d" % H] = ld
        assert np.abs(lp[:, 1] - lp[:, 0]).std() > 0.05, "degenerate logits"
    np.savez_compressed(args.out, **out)
    print("wrote", args.out, os.path.getsize(args.out), "bytes")


if __name__ == "__main__":
    main()
