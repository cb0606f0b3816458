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


def write_bgzf(path, text, threads=16):
    """BGZF framing (bgzip / bcl2fastq style: <= 64 KB members with their size in a 'BC' subfield), level 1."""
    import struct
    import zlib
    from concurrent.futures import ThreadPoolExecutor
    mv = memoryview(text)

    def member(i):
        c = mv[i:i + 0xff00]
        z = zlib.compressobj(1, zlib.DEFLATED, -15)
        d = z.compress(c) + z.flush()
        return (b"\x1f\x8b\x08\x04\0\0\0\0\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, 25 + len(d))
                + d + struct.pack("<II", zlib.crc32(c), len(c)))

    with open(path, "wb") as f, ThreadPoolExecutor(threads) as ex:
        for m in ex.map(member, range(0, len(mv), 0xff00), chunksize=64):
            f.write(m)
        f.write(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0\x1b\0\x03\0\0\0\0\0\0\0\0\0")


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    d = tempfile.mkdtemp(prefix="rdcli_")
    fasta = bool(os.environ.get("RD_CLI_FASTA", ""))       # FASTA input (">r%09d" / sequence, one line each) instead of FASTQ
    inp = os.path.join(d, "in.fa" if fasta else "in.fq")
    if fasta:
        seq, _ = synth.synth_reads_fixed(n, L, synth.SEED_BASE + 7)
        rec = np.empty((n, 11 + L + 1), np.uint8)
        rec[:, 0], rec[:, 1] = ord(">"), ord("r")
        idx = np.arange(n)
        for k in range(9):
            rec[:, 2 + 8 - k] = ord("0") + (idx // 10 ** k) % 10
        rec = np.concatenate([rec[:, :11], np.full((n, 1), 10, np.uint8), seq.reshape(n, L), np.full((n, 1), 10, np.uint8)], axis=1)
        rec.tofile(inp)
        size = rec.size
    else:
        size = write_fastq(inp, n, L, synth.SEED_BASE + 7)
    gz = os.environ.get("RD_CLI_GZ", "")                  # "bgzf": BGZF-framed input, "gzip": one plain gzip member
    if gz:
        text = np.fromfile(inp, np.uint8)
        os.remove(inp)
        inp += ".gz"
        if gz == "bgzf":
            write_bgzf(inp, text)
        else:
            import gzip
            with gzip.open(inp, "wb", compresslevel=1) as f:
                f.write(memoryview(text))
        print("input: %s, %.2f GB compressed" % (gz, os.path.getsize(inp) / 1e9), flush=True)
    paired = os.environ.get("RD_CLI_PAIRED", "")          # e.g. "rrna": a second file of mates, -e rrna (BASELINE configs[2] shape)
    if paired:
        inp2 = os.path.join(d, "in2.fq")
        size += write_fastq(inp2, n, L, synth.SEED_BASE + 8)
    for rep in range(2):
        for f in ("non.fq", "rrna.fq", "non2.fq", "rrna2.fq"):                 # (truncating a multi-GB output of the previous run is not the tool's time)
            if os.path.exists(os.path.join(d, f)):
                os.remove(os.path.join(d, f))
        t0 = time.perf_counter()
        argv = ["-l", str(L), "-i", inp, "-o", os.path.join(d, "non.fq"), "-r", os.path.join(d, "rrna.fq"),
                "-t", str(min(16, os.cpu_count() or 1))] + (["-d", os.environ["RD_CLI_DEVICES"]] if os.environ.get("RD_CLI_DEVICES") else []) \
            + os.environ.get("RD_CLI_EXTRA", "").split()
        if paired:
            argv[argv.index("-i") + 2:argv.index("-i") + 2] = [inp2]
            argv[argv.index("-o") + 2:argv.index("-o") + 2] = [os.path.join(d, "non2.fq")]
            argv[argv.index("-r") + 2:argv.index("-r") + 2] = [os.path.join(d, "rrna2.fq")]
            argv += ["-e", paired]
        args = detect.build_parser(True).parse_args(argv)
        pred = detect.Predictor(detect.ConfigParser.from_json(os.path.join(detect.cd, "config.json")), args)
        pred.load_model()
        t_load = time.perf_counter() - t0
        pred.detect()
        dt = time.perf_counter() - t0
        print("       load_model %.2f s, detect %.2f s" % (t_load, dt - t_load))
        k = 2 if paired else 1                            # a pair counts as 2 reads
        print("run %d: %d %s, %.2f GB of text in %.2f s = %.2f M reads/s (non-rRNA %d, rRNA %d)"
              % (rep, pred.num_seqs, "pairs" if paired else "reads", size / 1e9, dt, k * n / dt / 1e6, pred.num_nonrrna, pred.num_rrna), flush=True)
        print("       stage busy seconds:", {k: round(v, 3) for k, v in pred.stage_seconds.items()},
              "page-locking: %.2f s" % getattr(pred, "setup_seconds", 0.0), flush=True)
    for f in os.listdir(d):
        os.remove(os.path.join(d, f))
    os.rmdir(d)


if __name__ == "__main__":
    main()
