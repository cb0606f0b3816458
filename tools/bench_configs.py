"""Throughput of the other BASELINE.json config SHAPES on one B200 (informational; bench.py measures
config 1).  A pair counts as 2 reads.  python tools/bench_configs.py [precision]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ribodetector_b200.model import SeqModel              # noqa: E402
from ribodetector_b200.utils import synth                 # noqa: E402
from ribodetector_b200.utils.weights import load_weights  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else "tc_mixed"
    m = SeqModel(precision=prec)
    m.load_state_dict(load_weights())
    m.to("cuda:0")
    out = {"precision": prec, "configs": {}}
    n = 1 << 21

    def pin(a):
        return torch.from_numpy(a).pin_memory()

    # C3 / C4: paired end, -e rrna, 100 bp and 150 bp
    for name, L in (("C3 100 bp paired, -e rrna", 100), ("C4 150 bp paired, -e rrna", 150)):
        s1, o1 = synth.synth_reads_fixed(n, L, synth.SEED_BASE + 30 + L)
        s2, o2 = synth.synth_reads_fixed(n, L, synth.SEED_BASE + 31 + L)
        hs = [pin(s1), pin(o1), pin(s2), pin(o2)]
        ds = [t.cuda() for t in hs]
        labels = {"labels": torch.empty(n, dtype=torch.int8).pin_memory()}

        def dev():
            l1 = m.classify(ds[0], ds[1], L, want_labels=False)[0]
            l2 = m.classify(ds[2], ds[3], L, want_labels=False)[0]
            m.pair_combine(l1, l2, "rrna")

        def host():
            m.classify_pairs_host(hs[0], hs[1], hs[2], hs[3], L, mode="rrna", out=labels)

        td, th = timed(dev), timed(host)
        out["configs"][name] = {"pairs_per_call": n, "device_reads_per_s": 2 * n / td, "host_reads_per_s": 2 * n / th}
    # C5: 40-300 bp mixed single end, -l 300 and the README-recommended -l 170
    seq, off = synth.synth_reads(n, 40, 300, synth.SEED_BASE + 5)
    hs = [pin(seq), pin(off)]
    ds = [t.cuda() for t in hs]
    labels = {"labels": torch.empty(n, dtype=torch.int8).pin_memory()}
    for L in (300, 170):
        td = timed(lambda: m.classify(ds[0], ds[1], L))
        th = timed(lambda: m.classify_host(hs[0], hs[1], L, want_logits=False, out=labels))
        steps = float(np.minimum(off[1:] - off[:-1], L).mean())
        out["configs"]["C5 40-300 bp mixed single end, -l %d" % L] = {
            "reads_per_call": n, "mean_steps": steps, "device_reads_per_s": n / td, "host_reads_per_s": n / th,
            "device_read_steps_per_s": n * steps / td}
    # C2 shape for reference in the same units
    seq, off = synth.synth_reads_fixed(n, 100, synth.SEED_BASE + 2)
    ds = [torch.from_numpy(seq).cuda(), torch.from_numpy(off).cuda()]
    td = timed(lambda: m.classify(ds[0], ds[1], 100))
    out["configs"]["C2 100 bp single end"] = {"reads_per_call": n, "device_reads_per_s": n / td,
                                              "device_read_steps_per_s": n * 100 / td}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
