"""K0 (record scan) and K4 (label partition) of the device FASTQ path against the measured copy bandwidth
(MEASURED_PEAKS.json), CUDA events, text larger than L2; plus the streaming form's throughput over host
buffers (rd_fastq_submit / rd_fastq_collect, text already in page-locked memory).
python tools/bench_fastq_dev.py [n_reads] [read_len]"""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from ribodetector_b200.model import SeqModel              # noqa: E402
from ribodetector_b200.utils import synth                 # noqa: E402
from ribodetector_b200.utils.weights import load_weights  # noqa: E402
from bench_k1 import ev_time                               # noqa: E402


fastq_text = synth.fastq_text


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    peak = 6541.1
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p)).get("hbm_gbs", peak))
    m = SeqModel()
    m.load_state_dict(load_weights())
    m.to("cuda:0")
    lib, h = m._lib, m._need()
    text = fastq_text(n, L, synth.SEED_BASE + 9)
    B = text.size
    vp = lambda t: ctypes.c_void_p(t.data_ptr())          # noqa: E731
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    d_text = torch.from_numpy(text).cuda()
    rec = torch.empty((n + 8, 8), dtype=torch.int64, device="cuda")
    info = torch.empty(8, dtype=torch.int64, device="cuda")
    out = {"hbm_peak_gbs": peak, "reads": n, "read_len": L, "text_bytes": int(B), "kernels": {}}

    t = ev_time(lambda: lib.rd_scan_fastq_device(h, vp(d_text), B, 1, n + 8, vp(rec), vp(info), st))
    assert info.cpu().tolist()[1] == n
    nbytes = B + 32 * n + 32 * n + 64 * n + 4 * n         # text read once; line ends written + read; index written; 4 trailing-byte probes
    out["kernels"]["k0_scan"] = {"ms": t * 1e3, "algorithmic_bytes": int(nbytes), "gbs": nbytes / t / 1e9,
                                 "frac_of_peak": nbytes / t / 1e9 / peak, "text_gbs": B / t / 1e9}
    labels = (torch.rand(n, device="cuda") < 0.03).to(torch.int8)
    d_out = torch.empty(B + 16, dtype=torch.uint8, device="cuda")
    sizes = torch.zeros(3, dtype=torch.int64, device="cuda")
    t = ev_time(lambda: lib.rd_partition_records_device(h, vp(d_text), vp(rec), n, vp(labels), vp(d_out), vp(sizes), st))
    assert int(sizes.sum()) == B
    nbytes = 2 * B + 2 * (64 + 1) * n                      # text read + written; index and labels read twice (sizes, copy)
    out["kernels"]["k4_partition"] = {"ms": t * 1e3, "algorithmic_bytes": int(nbytes), "gbs": nbytes / t / 1e9,
                                      "frac_of_peak": nbytes / t / 1e9 / peak, "text_gbs": B / t / 1e9}

    # streaming form over page-locked host buffers: H2D + K0 + K1..K3 + K4 + D2H, two slots
    pin = [torch.from_numpy(text).pin_memory().numpy() for _ in range(2)]
    outs = [torch.empty(B + 32, dtype=torch.uint8, pin_memory=True).numpy() for _ in range(2)]
    for prec in ("tc_exact", "tc_auto"):
        blocks = 8
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for k in range(blocks):
                s = k & 1
                if k >= 2:
                    m.fastq_collect(s)
                got, _, _ = m.fastq_submit(s, [pin[s]], [B], True, n + 8, L, [outs[s]], precision=prec)
                assert got == n
            m.fastq_collect(0)
            m.fastq_collect(1)
            dt = time.perf_counter() - t0
        assert outs[0][:64].tobytes() != b"\0" * 64
        out["stream_" + prec] = {"reads_per_s": blocks * n / dt, "text_gbs": blocks * B / dt / 1e9, "blocks": blocks,
                                 "h2d_bytes_per_block": int(B), "d2h_bytes_per_block": int(B + 1 + 72)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
