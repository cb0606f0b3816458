"""Error of the tensor-core precisions against the on-device fp32 CUDA-core kernel on 2^20 reads per length
(run on the B200 box: python tests/prec_err_big.py [precisions...]).  Sets the stated tolerances of tc_mixed."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ribodetector_b200.model import SeqModel
from ribodetector_b200.utils import synth
from ribodetector_b200.utils.weights import load_weights
precs = sys.argv[1:] or ["tc_mixed", "tc_exact"]
m = SeqModel(); m.load_state_dict(load_weights()); m.to("cuda:0")
n = 1 << 20
def p1(x):
    return 1.0 / (1.0 + np.exp(x[:, 0] - x[:, 1]))
for L, lo, hi in ((100, 100, 100), (150, 150, 150), (300, 300, 300), (300, 40, 300)):
    seq, off = synth.synth_reads_fixed(n, L, 700 + L) if lo == hi else synth.synth_reads(n, lo, hi, 705)
    ref = m.classify(seq, off, L, precision="fp32")[0].cpu().numpy().astype(np.float64)
    margin = np.abs(ref[:, 1] - ref[:, 0])
    for p in precs:
        got = m.classify(seq, off, L, precision=p)[0].cpu().numpy().astype(np.float64)
        d = np.abs(got - ref).max(1)
        dm = np.abs((got[:, 1] - got[:, 0]) - (ref[:, 1] - ref[:, 0]))
        flips = (got[:, 1] > got[:, 0]) != (ref[:, 1] > ref[:, 0])
        print("L=%d (%d-%d) %-8s |dlogit| p50 %.1e p99 %.1e p99.99 %.1e max %.2e | max |dmargin| %.2e | max|dp| %.2e | flips %d, largest |margin| of a flipped read %.2e"
              % (L, lo, hi, p, np.percentile(d, 50), np.percentile(d, 99), np.percentile(d, 99.99), d.max(), dm.max(),
                 np.abs(p1(got) - p1(ref)).max(), int(flips.sum()), float(margin[flips].max()) if flips.any() else 0.0), flush=True)
