"""HBM-bound kernels of the path against the measured copy bandwidth (MEASURED_PEAKS.json):
the fp32 one-hot encoders (rd_encode_onehot: the reference's encode_read / encode_variable_len_read
output, 16 B written per base) and the K1 plan/codes stage of rd_classify.  CUDA events, inputs
larger than L2.  python tools/bench_k1.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ribodetector_b200.model import SeqModel              # noqa: E402
from ribodetector_b200.utils import synth                 # noqa: E402
from ribodetector_b200.utils.weights import load_weights  # noqa: E402


def ev_time(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps / 1e3


def main():
    peak = 6541.1
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p)).get("hbm_gbs", peak))
    m = SeqModel()
    m.load_state_dict(load_weights())
    m.to("cuda:0")
    n, L = 1 << 22, 100
    seq, off = synth.synth_reads_fixed(n, L, synth.SEED_BASE + 2)
    s, o = torch.from_numpy(seq).cuda(), torch.from_numpy(off).cuda()
    out = {"hbm_peak_gbs": peak, "reads": n, "read_len": L, "kernels": {}}
    # one-hot, padded layout [n, L, 4] fp32: reads n*L + 8n bytes, writes 16*n*L
    lib, h = m._lib, m._need()
    import ctypes
    dst = torch.empty((n, L, 4), dtype=torch.float32, device="cuda")
    row_off = torch.empty(n + 1, dtype=torch.int64, device="cuda")
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for name, layout in (("onehot_padded", 1), ("onehot_ragged", 0)):
        t = ev_time(lambda: lib.rd_encode_onehot(h, vp(s), vp(o), n, L, layout, vp(dst), vp(row_off) if layout == 0 else None, st))
        nbytes = n * L + 8 * n + 16 * n * L + (16 * n if layout == 0 else 0)
        out["kernels"][name] = {"ms": t * 1e3, "algorithmic_bytes": nbytes, "gbs": nbytes / t / 1e9, "frac_of_peak": nbytes / t / 1e9 / peak}
    # K1 of the classify path (plan + bucket scan + scatter + codes), via the per-stage timers
    m.set_timing(True)
    m.classify(s, o, L)
    m.get_timing(reset=True)
    for _ in range(5):
        m.classify(s, o, L, precision="tc_fast")
    tm = m.get_timing(reset=True)
    t = tm["plan"][0] / tm["plan"][1] / 1e3
    nbytes = 8 * n + n + 12 * n                        # offsets + last base read; plan/perm/splan written (tc modes)
    out["kernels"]["k1_plan_codes"] = {"ms": t * 1e3, "algorithmic_bytes": nbytes, "gbs": nbytes / t / 1e9, "frac_of_peak": nbytes / t / 1e9 / peak}
    t = tm["tail"][0] / tm["tail"][1] / 1e3
    nbytes = 17 * n
    out["kernels"]["k3_tail"] = {"ms": t * 1e3, "algorithmic_bytes": nbytes, "gbs": nbytes / t / 1e9, "frac_of_peak": nbytes / t / 1e9 / peak}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
