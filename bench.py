#!/usr/bin/env python
"""bench.py — reads/sec classified (100 bp) on N B200s, next to the CPU arm.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision P]
  (N>1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

A "step" = one pass of the hot path (K1 plan/encode → K2 LSTM → K3 tail) over one batch of
2^22 synthetic 100 bp single-end reads (419 MB of bases, larger than the 126 MB L2).  The
workload is BASELINE.json configs[1] "100 bp single-end, 50M synthetic reads, 1xB200": 50 M
reads are 12 such batches; K of them are timed.  Per-GPU work is fixed as N grows (weak).

  value  device-resident: sequence bytes + offsets already in HBM, labels stay in HBM; timed with
         CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.
  e2e    the same metric through SeqModel.classify_host (C ABI rd_classify_host): HOST pinned
         buffers in, HOST labels out, H2D and D2H inside the timed region.
  roofline      K2 (the LSTM kernel) timed live with CUDA events inside the timed region
                (rd_set_timing), algorithmic FLOPs per SURVEY.md §8d / DESIGN.md.  The default
                precision is tc_exact (3-pass fp16 split on tcgen05, fp32-grade logits); the
                single-pass tc_fast mode is timed once as well and reported under "fast_mode".
  cpu_baseline  the oracle port of the ribodetector_cpu loop (oracle/cpu_pipeline.py) on the
                host cores, rank 0, N=1 only, bounded sample.
  --impl reference   times that same CPU arm as the line's value.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

READ_LEN = 100
# tensor FLOPs EXECUTED per algorithmic FLOP: K = 128 + 16 input chunk, x3 passes for the fp16 split
EXECUTED_PER_ALGORITHMIC = {"fp32": 1.0, "tc_exact": 25.0 / 8.0, "tc_fast": 9.0 / 8.0, "tc_auto": 9.0 / 8.0,
                            "tc_mixed": 17.0 / 8.0}      # MMA issue slots (an 8-bit K=32 MMA costs what an fp16 K=16 one does)
# dram__bytes_read.sum + dram__bytes_write.sum of one K2 launch (ncu --set full capture of this bench's own launch
# size, 2^22 reads x 100 bp, profiles/r1_ncu_tc_exact_4m_summary.txt: 499.07 MB read + 34.44 MB written)
NCU_DRAM_BYTES_PER_READ = (499.070720e6 + 34.439168e6) / 4194304
MUFU_PER_READ = {"fp32": 10 * 128 * READ_LEN, "tc_exact": 7 * 128 * READ_LEN, "tc_fast": 5 * 128 * READ_LEN,
                 "tc_auto": 5 * 128 * READ_LEN, "tc_mixed": 7 * 128 * READ_LEN}
XU_LANES_PER_CLK_PER_SM = 16          # measured, tools/tc_rate.cu
BATCH_READS = 1 << 22
FLOP_PER_READ = 131072 * READ_LEN + 1024          # n*2*128*512 + 2*256*2
METRIC = "reads/sec classified (100 bp)"


OUT = sys.stdout        # main() points this at the real stdout and sends everything else on fd 1 to stderr


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"tflops": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0))),
                "hbm_gbs": float(d.get("hbm_gbs", 6650.0)), "source": "measured (sustained)"}
    return {"tflops": 1590.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "power_w_max": float(max(pw)) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_arm(weights, threads, batches_per_thread, seed):
    """The oracle port of the ribodetector_cpu loop on a bounded sample → (reads/s, sample str)."""
    from oracle import cpu_pipeline
    from ribodetector_b200.utils import synth
    n = threads * cpu_pipeline.BATCH * batches_per_thread
    seq, off = synth.synth_reads_fixed(n, READ_LEN, seed)
    _labels, dt = cpu_pipeline.classify(seq, off, READ_LEN, weights, threads)
    return n / dt, "%d x %d bp reads (%d batches of 1024 per worker), %.1f s" % (n, READ_LEN, batches_per_thread, dt)


def run_reference(args, rank, world):
    if rank != 0:
        return
    from ribodetector_b200.utils.weights import load_weights
    from ribodetector_b200.utils import synth
    weights = load_weights()
    threads = os.cpu_count() or 1
    vals = []
    sample = ""
    t_all = time.perf_counter()
    for i in range(args.warmup + args.steps):
        v, sample = cpu_arm(weights, threads, 2, synth.SEED_BASE + 100 + i)
        if i >= args.warmup:
            vals.append(v)
    n_step = threads * 1024 * 2
    ms = 1000.0 * n_step / float(np.mean(vals))
    value = n_step * len(vals) / sum(n_step / v for v in vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "100 bp single-end synthetic reads (BASELINE configs[1] shape), bounded "
                               "sample of %d reads per step" % n_step,
                   "note": "ORT unavailable - torch-CPU stand-in for ribodetector_cpu "
                           "(oracle/cpu_pipeline.py: forked 1-thread workers, batches of 1024, padded semantics)"},
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(line), file=OUT, flush=True)


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from ribodetector_b200 import shard
    from ribodetector_b200.model import SeqModel
    from ribodetector_b200.utils import synth
    from ribodetector_b200.utils.weights import load_weights

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    weights = load_weights()
    model = SeqModel(4, 128, 1, 2, pack_seq=True, precision=args.precision)
    model.load_state_dict(weights)
    model.to(dev).eval()

    n = args.reads_per_step
    nbuf = 2                                        # rotate distinct batches between steps
    host, devb = [], []
    for b in range(nbuf):
        seq, off = synth.synth_reads_fixed(n, READ_LEN, synth.SEED_BASE + 2 + 1000 * rank + b)
        hs = torch.from_numpy(seq).pin_memory()
        ho = torch.from_numpy(off).pin_memory()
        host.append((hs, ho))
        devb.append((hs.to(dev), ho.to(dev)))
    counts = torch.zeros(3, dtype=torch.int64, device=dev)
    out_host = {"labels": torch.empty(n, dtype=torch.int8).pin_memory()}
    model._lib.rd_reserve(model._need(), n, READ_LEN)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_device(i):
        s, o = devb[i % nbuf]
        return model.classify(s, o, READ_LEN, counts=counts)

    # ---- device-resident: value ---------------------------------------------------------------------
    for i in range(args.warmup):
        step_device(i)
    barrier()
    model.set_timing(True)
    model.get_timing(reset=True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = model.kernel_launches()
    counts.zero_()
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step_device(i)
    shard.allreduce_counts(counts)                  # the path's only exchange: int64[3] label counts (no-op at N=1)
    e1.record()
    barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = model.kernel_launches() - launches0
    clocks = sampler.stop() if rank == 0 else None
    timing = model.get_timing(reset=True)
    model.set_timing(False)
    total_counts = counts.cpu().tolist()

    # ---- the single-pass mode, device-resident, for information (same timing protocol) -----------------
    def time_mode(prec):
        for i in range(3):
            model.classify(devb[i % nbuf][0], devb[i % nbuf][1], READ_LEN, precision=prec)
        barrier()
        f0 = torch.cuda.Event(enable_timing=True)
        f1 = torch.cuda.Event(enable_timing=True)
        f0.record()
        for i in range(args.steps):
            model.classify(devb[i % nbuf][0], devb[i % nbuf][1], READ_LEN, precision=prec)
        f1.record()
        barrier()
        ms = f0.elapsed_time(f1)
        if world > 1:
            tf = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tf, op=dist.ReduceOp.MAX)
            ms = float(tf.item())
        return {"precision": prec, "value": world * n * args.steps / (ms / 1000.0), "unit": "reads/s",
                "ms_per_step": ms / args.steps}

    fast = auto = None
    if args.precision != "tc_fast" and not args.no_fast:
        fast = time_mode("tc_fast")
        fast["note"] = "single fp16 pass + tanh.approx: |dlogit| <= 5e-2, ~0.01-0.1 % label flips vs fp32"
        auto = time_mode("tc_auto")
        auto["note"] = ("tc_fast over all reads + tc_exact over the ~1 % with |margin| < 0.25: labels identical to "
                        "tc_exact, logits exact-grade only inside the band")

    # ---- end to end: host buffers through the public API ----------------------------------------------
    for i in range(min(args.warmup, 2)):
        model.classify_host(host[i % nbuf][0], host[i % nbuf][1], READ_LEN, want_logits=False, out=out_host)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        r = model.classify_host(host[i % nbuf][0], host[i % nbuf][1], READ_LEN, want_logits=False, out=out_host)
    torch.cuda.synchronize(dev)
    e2e_ms = (time.perf_counter() - t0) * 1000.0
    e2e_counts = r["counts"].tolist()
    # ---- the edges included: FASTQ TEXT in host memory -> label-partitioned record text in host memory ----------
    # (rd_fastq_submit / rd_fastq_collect: H2D, K0 record scan, K1-K3, K4 partition, D2H; two slots in flight)
    fq_n = min(n, 1 << 21)
    fq_blocks = max(2, (args.steps * n) // (2 * fq_n))
    text = torch.from_numpy(synth.fastq_text(fq_n, READ_LEN, synth.SEED_BASE + 5 + rank)).pin_memory().numpy()
    fq_out = [torch.empty(text.size + 32, dtype=torch.uint8).pin_memory().numpy() for _ in range(2)]
    fq_counts = np.zeros(3, np.int64)
    for rep in range(2):                            # the first pass sizes the slots and warms up
        barrier()
        t0 = time.perf_counter()
        for k in range(fq_blocks):
            if k >= 2:
                fq_counts += model.fastq_collect(k & 1)[1]
            got, _, _ = model.fastq_submit(k & 1, [text], [text.size], True, fq_n + 8, READ_LEN, [fq_out[k & 1]])
            assert got == fq_n
        for k in range(2):
            fq_counts += model.fastq_collect(k)[1]
        fq_ms = (time.perf_counter() - t0) * 1000.0
    if world > 1:
        t = torch.tensor([dev_ms, e2e_ms, fq_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms, fq_ms = t.tolist()
    h2d = host[0][0].numel() + host[0][1].numel() * 8
    d2h = n + 24

    if rank == 0:
        peaks = measured_peaks()
        lstm_ms, lstm_n = timing["lstm"]
        lstm_avg_s = (lstm_ms / max(lstm_n, 1)) / 1000.0
        achieved = FLOP_PER_READ * n / lstm_avg_s / 1e12 if lstm_avg_s > 0 else 0.0
        value = world * n * args.steps / (dev_ms / 1000.0)
        e2e = world * n * args.steps / (e2e_ms / 1000.0)
        line = {
            "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32", "tc_exact": "f16x2-split/f32-acc", "tc_fast": "f16/f32-acc",
                      "tc_auto": "f16/f32-acc + f16x2-split/f32-acc on low-margin reads",
                      "tc_mixed": "f16 + e5m2 corrections/f32-acc"}[args.precision],
            "data": "synthetic",
            "config": {"workload": "100 bp single-end, 50M synthetic reads per B200 (BASELINE configs[1]): "
                                   "timed as %d batches of %d reads on each of %d GPU(s)" % (args.steps, n, world),
                       "read_len": READ_LEN, "reads_per_step_per_gpu": n, "precision": args.precision,
                       "semantics": "packed", "l2": "inputs larger than L2 (%.0f MB bases per step, 2 batches rotated)"
                                                    % (n * READ_LEN / 1e6),
                       "parallelism": "reads sharded, %d rank(s), NCCL all-reduce of int64[3] label counts" % world},
            "e2e": {"value": e2e, "unit": "reads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "K2 forward LSTM (%s)" % args.precision,
                         "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["tflops"],
                         "traffic": NCU_DRAM_BYTES_PER_READ * n if args.precision == "tc_exact" and READ_LEN == 100 else None,
                         "traffic_note": "bytes per launch from the ncu --set full capture of a 2^22-read launch (scaled per read for "
                                         "other sizes); algorithmic bytes per launch = %d (sequence + offset + slot plan/perm read, logits written)" % (n * (READ_LEN + 8 + 8 + 8)),
                         "executed_tflops": achieved * EXECUTED_PER_ALGORITHMIC[args.precision],
                         "executed_frac": achieved * EXECUTED_PER_ALGORITHMIC[args.precision] / peaks["tflops"],
                         "peak_source": peaks["source"], "launch_ms": lstm_avg_s * 1000.0,
                         "flop_per_launch": FLOP_PER_READ * n,
                         "share_of_step": lstm_ms / dev_ms if dev_ms else None,
                         "co_bound": {"pipe": "xu (MUFU sigmoid/tanh)", "ops_per_read": MUFU_PER_READ[args.precision],
                                      "achieved_gops": MUFU_PER_READ[args.precision] * n / lstm_avg_s / 1e9 if lstm_avg_s > 0 else 0.0,
                                      "peak_gops": XU_LANES_PER_CLK_PER_SM * 148 * ((clocks or {}).get("sm_mhz") or 1965.0) / 1e3,
                                      "note": "peak = 16 MUFU lanes/clk/SM (measured) x 148 SMs x SM clock under load"}},
            "e2e_fastq": {"value": world * fq_n * fq_blocks / (fq_ms / 1000.0), "unit": "reads/s",
                          "text_gb_per_s_each_way": world * text.size * fq_blocks / (fq_ms / 1000.0) / 1e9,
                          "h2d_bytes_per_block": int(text.size), "d2h_bytes_per_block": int(text.size + 1 + 72),
                          "reads_per_block": fq_n, "blocks": fq_blocks,
                          "note": "FASTQ text in page-locked host memory -> record scan (K0), classify, label partition (K4) on the "
                                  "GPU -> partitioned record text back in host memory (rd_fastq_submit / rd_fastq_collect)"},
            "stage_ms": {k: v[0] / max(v[1], 1) for k, v in timing.items() if v[1]},
            "label_counts": total_counts, "e2e_label_counts_last_step": e2e_counts,
        }
        cb = line["roofline"]["co_bound"]
        cb["frac"] = cb["achieved_gops"] / cb["peak_gops"] if cb["peak_gops"] else None
        if fast:
            line["fast_mode"] = fast
            line["auto_mode"] = auto
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            v, sample = cpu_arm(weights, threads, args.cpu_batches, synth.SEED_BASE + 99)
            line["cpu_baseline"] = {"value": v, "unit": "reads/s", "cores": threads, "kind": "port",
                                    "sample": sample,
                                    "note": "ORT unavailable - torch-CPU stand-in for ribodetector_cpu"}
        print(json.dumps(line), file=OUT, flush=True)
    model.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("RD_BENCH_PRECISION", "tc_exact"),
                    choices=["fp32", "tc_exact", "tc_fast", "tc_auto", "tc_mixed"])
    ap.add_argument("--reads-per-step", type=int, default=BATCH_READS)
    ap.add_argument("--cpu-batches", type=int, default=28, help="1024-read batches per CPU worker in cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fast", action="store_true", help="skip the informational tc_fast timing")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.steps is None:
        args.steps = 12 if args.impl == "ours" else 3      # 12 x 2^22 reads = the 50 M reads of configs[1]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    # stdout carries exactly ONE JSON line: libraries that print to fd 1 (NCCL's version banner, forked
    # workers) are sent to stderr instead
    global OUT
    sys.stdout.flush()
    OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
