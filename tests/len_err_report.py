"""Error of the three precisions against the fp64 oracle as a function of read length (run on the B200 box:
python tests/len_err_report.py).  Lives under tests/ because it uses the oracle."""
import sys, numpy as np, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ribodetector_b200.model import SeqModel
from ribodetector_b200.utils import synth
from ribodetector_b200.utils.weights import load_weights
from oracle.model_numpy import NumpyOracle
from oracle.model_torch import TorchOracle
w = load_weights()
m = SeqModel(); m.load_state_dict(w); m.to("cuda:0")
o64 = NumpyOracle(w, np.float64); ot = TorchOracle(w)
for L in (100, 200, 300):
    seq, off = synth.synth_reads_fixed(3000, L, 900 + L)
    reads = synth.to_strings(seq, off)
    ref = o64.logits(reads, L, "packed")
    tor = ot.logits_packed(reads, L)
    out = {p: m.classify(seq, off, L, precision=p)[0].cpu().numpy().astype(np.float64) for p in ("fp32", "tc_exact", "tc_fast", "tc_mixed", "tc_mixed_raw", "tc_auto")}
    print("L=%d  torch-fp32 vs f64: %.2e | fp32 kernel: %.2e  tc_exact: %.2e  tc_fast: %.2e | tc_exact vs fp32 kernel %.2e  vs torch %.2e" % (
        L, np.abs(tor - ref).max(), np.abs(out["fp32"] - ref).max(), np.abs(out["tc_exact"] - ref).max(), np.abs(out["tc_fast"] - ref).max(),
        np.abs(out["tc_exact"] - out["fp32"]).max(), np.abs(out["tc_exact"] - tor).max()))
    def sm(x):
        return 1.0 / (1.0 + np.exp(x[:, 0] - x[:, 1]))
    for p in ("tc_exact", "tc_mixed", "tc_mixed_raw", "tc_auto", "tc_fast"):
        d = np.abs(out[p] - ref).max(1)
        flips = int(((out[p][:, 1] > out[p][:, 0]) != (ref[:, 1] > ref[:, 0])).sum())
        print("   %-8s |dlogit| percentiles 50/99/99.9/max: %.1e %.1e %.1e %.1e   max|dp| %.1e   label flips %d" % (
            (p,) + tuple(np.percentile(d, [50, 99, 99.9, 100])) + (np.abs(sm(out[p]) - sm(ref)).max(), flips)))
