"""Cycle breakdown of the tensor-core LSTM kernel roles (needs tools/librd_prof.so built with -DRD_TC_PROFILE)."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ribodetector_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "tools", "librd_prof.so")
from ribodetector_b200.model import SeqModel
from ribodetector_b200.utils import synth
from ribodetector_b200.utils.weights import load_weights
m = SeqModel(); m.load_state_dict(load_weights()); m.to("cuda:0")
n = 148 * 128 * 4
seq, off = synth.synth_reads_fixed(n, 100, 5)
s, o = torch.from_numpy(seq).cuda(), torch.from_numpy(off).cuda()
lib = m._lib
lib.rd_debug_prof.argtypes = [ctypes.c_void_p]
for prec in ("tc_fast", "tc_exact", "tc_mixed"):
    for _ in range(2):
        m.classify(s, o, 100, precision=prec)
    torch.cuda.synchronize()
    out = (ctypes.c_uint64 * 16)()
    lib.rd_debug_prof(ctypes.cast(out, ctypes.c_void_p))
    v = list(out)
    steps = 4 * 100 if prec == "tc_fast" else 8 * 100 // 2 * 1
    tiles = 4
    print(prec, "block 0: tiles/unit=4, steps=400")
    print("  MMA warp : total %d cyc (%.0f/step)  wait acc_empty %d  wait h_ready %d  wait tile %d" % (v[0], v[0] / 400, v[1], v[2], v[3]))
    print("  MMA chunk mc=2 issue->complete: %.0f cycles per chunk (%.1f per MMA)" % (v[15] / 396.0, v[15] / 396.0 / {"tc_fast": 9, "tc_exact": 25, "tc_mixed": 17}[prec]))
    for nm, b in (("epi warp0", 4), ("epi warp5", 10)):
        print("  %s: total %d cyc (%.0f/step)  wait acc_full %d (%.0f/step)  tmem ld %d (%.0f/step)  st+arrive %d (%.0f/step)  [acc_full wait at mc=0: %.0f/step]" %
              (nm, v[b], v[b] / 400, v[b + 1], v[b + 1] / 400, v[b + 2], v[b + 2] / 400, v[b + 3], v[b + 3] / 400, v[b + 4] / 400))
