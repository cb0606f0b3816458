#!/usr/bin/env python
"""Throughput of the fp32 CUDA-core kernel (rd_lstm_fp32.cu) over the hidden sizes it takes (the shipped 128 included):
device-resident 100 bp reads, CUDA-event timed, 3 warm-up + 5 timed calls per size.
    python tools/bench_fp32.py [n_reads] -> profiles/r2_fp32_hidden.json"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ribodetector_b200.model import SeqModel            # noqa: E402
from ribodetector_b200.utils import synth               # noqa: E402
from ribodetector_b200.utils.weights import load_weights  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 19
    seq, off = synth.synth_reads_fixed(n, 100, 4242)
    d_seq, d_off = torch.from_numpy(seq).cuda(), torch.from_numpy(off).cuda()
    out = {"reads": n, "read_len": 100, "sizes": {}}
    for H in (32, 64, 96, 128, 160, 192, 256):
        m = SeqModel(4, H, 1, 2, precision="fp32")
        m.load_state_dict(load_weights() if H == 128 else synth.synth_weights(H, 7))
        m.to("cuda:0").eval()
        for _ in range(3):
            m.classify(d_seq, d_off, 100)
        torch.cuda.synchronize()
        m.set_timing(True)
        m.get_timing()
        for _ in range(5):
            m.classify(d_seq, d_off, 100)
        torch.cuda.synchronize()
        ms = m.get_timing()["lstm"][0] / 5
        fma = 4.0 * H * H * 100 * n          # 4H x H multiply-adds per read-step
        out["sizes"][str(H)] = {"kernel": "lstm_fp32_kernel",
                                "ms": ms, "reads_per_s": n / ms * 1e3, "fp32_tflops": 2 * fma / ms / 1e9}
        print(H, out["sizes"][str(H)], flush=True)
        m.close()
    with open(os.path.join(ROOT, "profiles", "r2_fp32_hidden.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
