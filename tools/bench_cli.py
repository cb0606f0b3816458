"""L-cli measurement: wall clock of the `ribodetector` command line on a synthetic
FASTQ file, parse + classify + write included.  python tools/bench_cli.py [n_reads] [read_len]"""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ribodetector_b200 import detect                     # noqa: E402
from ribodetector_b200.utils import synth                # noqa: E402


def write_fastq(path, n, L, seed):
    rec = synth.fastq_text(n, L, seed)
    rec.tofile(path)
    return rec.size


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    d = tempfile.mkdtemp(prefix="rdcli_")
    inp = os.path.join(d, "in.fq")
    size = write_fastq(inp, n, L, synth.SEED_BASE + 7)
    for rep in range(2):
        for f in ("non.fq", "rrna.fq"):                 # (truncating a multi-GB output of the previous run is not the tool's time)
            if os.path.exists(os.path.join(d, f)):
                os.remove(os.path.join(d, f))
        t0 = time.perf_counter()
        argv = ["-l", str(L), "-i", inp, "-o", os.path.join(d, "non.fq"), "-r", os.path.join(d, "rrna.fq"),
                "-t", str(min(16, os.cpu_count() or 1))] + (["-d", os.environ["RD_CLI_DEVICES"]] if os.environ.get("RD_CLI_DEVICES") else []) \
            + os.environ.get("RD_CLI_EXTRA", "").split()
        args = detect.build_parser(True).parse_args(argv)
        pred = detect.Predictor(detect.ConfigParser.from_json(os.path.join(detect.cd, "config.json")), args)
        pred.load_model()
        t_load = time.perf_counter() - t0
        pred.detect()
        dt = time.perf_counter() - t0
        print("       load_model %.2f s, detect %.2f s" % (t_load, dt - t_load))
        print("run %d: %d reads, %.2f GB FASTQ in %.2f s = %.2f M reads/s (non-rRNA %d, rRNA %d)"
              % (rep, pred.num_seqs, size / 1e9, dt, n / dt / 1e6, pred.num_nonrrna, pred.num_rrna), flush=True)
        print("       stage busy seconds:", {k: round(v, 3) for k, v in pred.stage_seconds.items()},
              "page-locking: %.2f s" % getattr(pred, "setup_seconds", 0.0), flush=True)
    for f in os.listdir(d):
        os.remove(os.path.join(d, f))
    os.rmdir(d)


if __name__ == "__main__":
    main()
