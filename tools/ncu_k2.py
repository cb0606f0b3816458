"""One launch set of the LSTM kernel for ncu: python tools/ncu_k2.py [precision] [n_reads] [read_len | lo-hi]."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ribodetector_b200.model import SeqModel
from ribodetector_b200.utils import synth
from ribodetector_b200.utils.weights import load_weights
prec = sys.argv[1] if len(sys.argv) > 1 else "tc_mixed"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
ls = sys.argv[3] if len(sys.argv) > 3 else "100"
H = int(os.environ.get("RD_NCU_HIDDEN", "128"))          # another hidden size: the fp32 CUDA-core kernel on seeded weights
m = SeqModel(hidden_size=H, precision=prec); m.load_state_dict(load_weights() if H == 128 else synth.synth_weights(H, 7)); m.to("cuda:0")
if "-" in ls:
    lo, hi = (int(x) for x in ls.split("-"))
    seq, off = synth.synth_reads(n, lo, hi, synth.SEED_BASE + 5)
    L = hi
else:
    L = int(ls)
    seq, off = synth.synth_reads_fixed(n, L, synth.SEED_BASE + 2)
s, o = torch.from_numpy(seq).cuda(), torch.from_numpy(off).cuda()
for _ in range(3):
    m.classify(s, o, L)
torch.cuda.synchronize()
