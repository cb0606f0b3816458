"""Stability check on the B200 box: repeated create / classify / destroy (no leaks, identical bits) and a
long run of full-size launches.  python tools/soak.py"""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ribodetector_b200.model import SeqModel
from ribodetector_b200.utils import synth
from ribodetector_b200.utils.weights import load_weights
w = load_weights()
seq, off = synth.synth_reads(50000, 20, 150, 1)
ref = None
t0 = time.time()
for i in range(25):
    m = SeqModel(precision=("tc_exact", "tc_fast", "tc_auto", "fp32")[i % 4]); m.load_state_dict(w); m.to("cuda:0")
    r = m.classify_host(seq, off, 100)
    if i % 4 == 0:
        if ref is None: ref = r["logits"].numpy().copy()
        assert np.array_equal(ref, r["logits"].numpy())
    m.close()
print("25 create/classify/destroy cycles ok in %.1f s; mem allocated by torch: %d, free/total %s" % (time.time() - t0, torch.cuda.memory_allocated(), torch.cuda.mem_get_info()))
m = SeqModel(); m.load_state_dict(w); m.to("cuda:0")
s, o = synth.synth_reads_fixed(1 << 22, 100, 9)
s, o = torch.from_numpy(s).cuda(), torch.from_numpy(o).cuda()
c = torch.zeros(3, dtype=torch.int64, device="cuda")
t0 = time.time()
for i in range(60):
    m.classify(s, o, 100, counts=c, precision=("tc_exact", "tc_auto")[i % 2])
torch.cuda.synchronize()
print("60 x 4Mi-read launches in %.1f s, counts %s" % (time.time() - t0, c.tolist()))
