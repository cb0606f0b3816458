#!/usr/bin/env python
"""bench.py — reads/sec classified (100 bp) on N B200s, next to the CPU arm.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision P]
  (N>1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

A "step" = one pass of the hot path (K1 plan → K2 LSTM → K3 tail) over one batch of 2^22 synthetic 100 bp
single-end reads (419 MB of bases, larger than the 126 MB L2).  The workload is BASELINE.json configs[1]
"100 bp single-end, 50M synthetic reads, 1xB200": 50 M reads are 12 such batches; K of them are timed.  Per-GPU
work is fixed as N grows (weak scaling).

  value     device-resident: sequence bytes + offsets already in HBM, labels stay in HBM; CUDA events on the
            launching stream, barrier + synchronize on both sides, max over ranks.
  e2e       the same metric through SeqModel.classify_host (C ABI rd_classify_host): HOST pinned buffers in, HOST
            labels out, H2D and D2H inside the timed region.
  roofline  K2 (the LSTM kernel) timed live with CUDA events inside the timed region (rd_set_timing), algorithmic
            FLOPs per SURVEY.md §8d / DESIGN.md; `traffic` from profiles/k2_traffic.json (ncu dram bytes of this
            kernel source, null when the source changed since the capture).
  parity    the headline precision against the CPU arm's logits on the reads the CPU arm classified (a slice of the
            timed batch), outside the timed region; plus against the on-device fp32 CUDA-core kernel.
  configs   BASELINE.json configs[2..4] (C3 100 bp pairs -e rrna, C4 150 bp pairs, C5 40-300 bp mixed -l 300) with
            the same timing protocol, a few steps each: {value, e2e, roofline_frac}.
  strong    a FIXED set of read pairs from ONE source buffer in shared host memory, cut by shard_bounds over the N
            ranks, rd_classify_pairs_host per rank, labels gathered in input order and compared bit for bit with
            the labels one GPU computes for the whole set; counts through the NCCL all-reduce.
  cpu_baseline   the ribodetector_cpu loop on the host cores, rank 0, N=1 only, bounded sample: the reference's own
            encoder + model_cpu.SeqModel when the reference package is importable (oracle/ref_cpu_arm.py, kind
            "reference"), else the oracle port (oracle/cpu_pipeline.py, kind "port").
  --impl reference   times that same CPU arm as the line's value.
"""
import argparse
import hashlib
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
# tensor-core issue slots EXECUTED per algorithmic one: K = 128 needs 8 fp16 MMAs per chunk; + the input/bias chunk;
# x3 passes for the fp16 split; tc_mixed: 8 + 1 + 8 8-bit MMAs (an 8-bit K=32 MMA costs what an fp16 K=16 one does)
EXECUTED_PER_ALGORITHMIC = {"fp32": 1.0, "tc_exact": 25.0 / 8.0, "tc_fast": 9.0 / 8.0, "tc_auto": 9.0 / 8.0,
                            "tc_mixed": 17.0 / 8.0, "tc_mixed_raw": 17.0 / 8.0}
MUFU_PER_UNIT_STEP = {"fp32": 10, "tc_exact": 7, "tc_fast": 5, "tc_auto": 5, "tc_mixed": 7, "tc_mixed_raw": 7}
XU_LANES_PER_CLK_PER_SM = 16          # measured, tools/tc_rate.cu
BATCH_READS = 1 << 22
METRIC = "reads/sec classified (100 bp)"
DTYPE = {"fp32": "f32", "tc_exact": "f16x2-split/f32-acc", "tc_fast": "f16/f32-acc",
         "tc_auto": "f16/f32-acc + f16x2-split/f32-acc on low-margin reads",
         "tc_mixed": "f16 + e5m2 correction pass/f32-acc (+ f16x2-split on low-margin reads)",
         "tc_mixed_raw": "f16 + e5m2 correction pass/f32-acc"}
# stated tolerances at 100 bp (tests/test_gpu_parity.py): max |dlogit|, max |dprob|, label band (labels must match
# the reference's outside it)
TOL = {"fp32": (2e-4, 1e-4, 4e-4), "tc_exact": (2e-4, 1e-4, 4e-4), "tc_mixed": (3e-3, 1e-3, 4e-4),
       "tc_fast": (5e-2, 2e-2, 1e-1), "tc_auto": (5e-2, 2e-2, 4e-4), "tc_mixed_raw": (3e-3, 1e-3, 6e-3)}


def flop_per_read(mean_steps):
    """n * 2*128*512 (forward recurrent GEMM) + 2*256*2 (FC)."""
    return 131072.0 * mean_steps + 1024.0


OUT = sys.stdout        # main() points this at the real stdout and sends everything else on fd 1 to stderr
_T0 = time.perf_counter()


def log(msg):
    """progress on stderr (rank 0): which phase a slow or stuck run is in"""
    if int(os.environ.get("RANK", "0")) == 0:
        print("[bench %6.1fs] %s" % (time.perf_counter() - _T0, msg), file=sys.stderr, flush=True)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"tflops": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0))),
                "hbm_gbs": float(d.get("hbm_gbs", 6650.0)), "source": "measured (sustained)"}
    return {"tflops": 1590.0, "hbm_gbs": 6650.0, "source": "fallback"}


def k2_traffic(precision, read_len):
    """dram bytes per read of one K2 launch from the committed ncu capture — only if it was taken from THIS kernel
    source (sha256 of rd_lstm_tc.cu), else (None, reason)."""
    p = os.path.join(ROOT, "profiles", "k2_traffic.json")
    src = os.path.join(ROOT, "ribodetector_b200", "csrc", "rd_lstm_tc.cu")
    try:
        with open(p) as f:
            d = json.load(f)
        with open(src, "rb") as f:
            sha = hashlib.sha256(f.read()).hexdigest()
    except OSError as e:
        return None, "no capture: %s" % e
    e = d.get("%s_L%d" % (precision, read_len))
    if not e:
        return None, "no ncu capture for %s at %d bp in profiles/k2_traffic.json" % (precision, read_len)
    if e.get("src_sha") != sha:
        return None, "rd_lstm_tc.cu changed since the ncu capture (%s): re-run tools/ncu_summary.py --traffic" % e.get("report")
    return float(e["bytes_per_read"]), "ncu --set full, %s (%d reads per launch)" % (e.get("report"), e.get("reads", 0))


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


# ---- the CPU arm ----------------------------------------------------------------------------------------
def cpu_arm(weights, threads, seq, off, max_len=READ_LEN, split=True):
    """ribodetector_cpu's loop over (seq, off) on `threads` host cores → (info dict, logits, labels).  The
    reference's own code when its package is importable, else the oracle port."""
    from oracle import ref_cpu_arm
    n = len(off) - 1
    root = ref_cpu_arm.find_reference()
    if root:
        labels, logits, dt = ref_cpu_arm.classify(seq, off, max_len, threads, root)
        extra = ref_cpu_arm.single_core_split(seq, off, max_len, root) if split else {}
        kind = "reference"
        note = ("the reference's own encode_variable_len_read + model_cpu.SeqModel (the module its .onnx was exported "
                "from, its own .pth) imported from %s in forked 1-thread workers, batches of 1024 "
                "(detect_cpu.py:686-708); onnxruntime is not installed, so the torch module stands in for the ORT "
                "session" % ("/root/reference" if root == "/root/reference" else "baseline/_ref"))
    else:
        from oracle import cpu_pipeline
        labels, logits, dt = cpu_pipeline.classify(seq, off, max_len, weights, threads)
        extra = {}
        kind = "port"
        note = "reference package not importable: oracle port of the ribodetector_cpu loop (oracle/cpu_pipeline.py)"
    out = {"value": n / dt, "unit": "reads/s", "cores": threads, "kind": kind,
           "sample": "%d x %d bp reads (%d batches of 1024 per worker), %.1f s" % (n, max_len, n // (1024 * threads), dt),
           "reads_per_s_per_core": n / dt / threads, "note": note}
    out.update(extra)
    return out, logits, labels


def parity_block(got, ref, prec, vs=""):
    tol_l, tol_p, band = TOL[prec]
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)

    def p1(x):
        return 1.0 / (1.0 + np.exp(x[:, 0] - x[:, 1]))
    margin = np.abs(ref[:, 1] - ref[:, 0])
    flips = (got[:, 1] > got[:, 0]) != (ref[:, 1] > ref[:, 0])
    d = {"n": int(len(ref)), "vs": vs, "max_dlogit": float(np.abs(got - ref).max()),
         "max_dp": float(np.abs(p1(got) - p1(ref)).max()), "flips": int(flips.sum()),
         "flips_outside_band": int(flips[margin > band].sum()), "reads_inside_band": int((margin <= band).sum()),
         "tolerance": {"dlogit": tol_l, "dp": tol_p, "label_band": band}}
    d["ok"] = bool(d["max_dlogit"] <= tol_l and d["max_dp"] <= tol_p and d["flips_outside_band"] == 0)
    return d


def run_reference(args, rank, world):
    if rank != 0:
        return
    from ribodetector_b200.utils.weights import load_weights
    from ribodetector_b200.utils import synth
    weights = load_weights()
    threads = os.cpu_count() or 1
    vals, info = [], None
    t_all = time.perf_counter()
    n_step = threads * 1024 * 2
    for i in range(args.warmup + args.steps):
        seq, off = synth.synth_reads_fixed(n_step, READ_LEN, synth.SEED_BASE + 100 + i)
        info, _, _ = cpu_arm(weights, threads, seq, off, split=(i == args.warmup + args.steps - 1))
        if i >= args.warmup:
            vals.append(info["value"])
    ms = 1000.0 * n_step / float(np.mean(vals))
    value = n_step * len(vals) / sum(n_step / v for v in vals)
    info["value"] = value
    info["reads_per_s_per_core"] = value / threads
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "100 bp single-end synthetic reads (BASELINE configs[1] shape), bounded "
                               "sample of %d reads per step" % n_step,
                   "note": "ORT unavailable - torch-CPU stand-in for ribodetector_cpu: " + info["note"]},
        "cpu_baseline": info,
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(line), file=OUT, flush=True)


# ---- our arm ----------------------------------------------------------------------------------------------
class Runner:
    def __init__(self, args, rank, world, local_rank):
        import torch
        import torch.distributed as dist
        from ribodetector_b200.model import SeqModel
        from ribodetector_b200.utils.weights import load_weights
        self.torch, self.dist, self.args = torch, dist, args
        self.rank, self.world = rank, world
        torch.cuda.set_device(local_rank)
        self.dev = torch.device("cuda", local_rank)
        if world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.weights = load_weights()
        self.model = SeqModel(4, 128, 1, 2, pack_seq=True, precision=args.precision)
        self.model.load_state_dict(self.weights)
        self.model.to(self.dev).eval()
        self.peaks = measured_peaks()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, *vals):
        if self.world == 1:
            return list(vals)
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def time_device(self, step, steps, warmup):
        """W untimed + K timed calls of step(i) on the current stream → (ms, stage timing, launches)."""
        torch = self.torch
        for i in range(warmup):
            step(i)
        self.barrier()
        self.model.set_timing(True)
        self.model.get_timing(reset=True)
        l0 = self.model.kernel_launches()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step(i)
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1)
        timing = self.model.get_timing(reset=True)
        self.model.set_timing(False)
        return ms, timing, self.model.kernel_launches() - l0

    def time_host(self, call, steps, warmup):
        r = None
        for i in range(warmup):
            call(i)
        self.barrier()
        t0 = time.perf_counter()
        for i in range(steps):
            r = call(i)
        self.torch.cuda.synchronize(self.dev)
        return (time.perf_counter() - t0) * 1000.0, r


def run_configs(R, pin):
    """BASELINE.json configs[2..4] with the headline's timing protocol (3 warm-up + K timed steps, CUDA events, inputs
    larger than L2, max over ranks).  A pair counts as 2 reads."""
    from ribodetector_b200.utils import synth
    torch, model, args, world, rank = R.torch, R.model, R.args, R.world, R.rank
    K = args.config_steps
    out = {}
    n = args.reads_per_step // 2
    for key, L, cfg_idx in (("C3", 100, 2), ("C4", 150, 3)):
        s1, o1 = synth.synth_reads_fixed(n, L, synth.SEED_BASE + 30 + L + 1000 * rank)
        s2, o2 = synth.synth_reads_fixed(n, L, synth.SEED_BASE + 31 + L + 1000 * rank)
        hs = [pin(s1), pin(o1), pin(s2), pin(o2)]
        ds = [t.to(R.dev) for t in hs]
        labels = {"labels": torch.empty(n, dtype=torch.int8).pin_memory()}
        counts = torch.zeros(3, dtype=torch.int64, device=R.dev)

        def dev_step(i):
            l1 = model.classify(ds[0], ds[1], L, want_labels=False)[0]
            l2 = model.classify(ds[2], ds[3], L, want_labels=False)[0]
            model.pair_combine(l1, l2, "rrna", counts=counts)

        ms, timing, _ = R.time_device(dev_step, K, 3)
        hms, _ = R.time_host(lambda i: model.classify_pairs_host(hs[0], hs[1], hs[2], hs[3], L, mode="rrna", out=labels), K, 1)
        ms, hms = R.max_over_ranks(ms, hms)
        lstm_ms, lstm_n = timing["lstm"]
        ach = flop_per_read(L) * n / (lstm_ms / max(lstm_n, 1) / 1e3) / 1e12        # one K2 launch per end
        out[key] = {"workload": "%d bp paired-end, -e rrna (BASELINE configs[%d] shape): %d pairs per step per GPU, %d steps"
                                % (L, cfg_idx, n, K),
                    "value": world * 2 * n * K / (ms / 1e3), "unit": "reads/s", "ms_per_step": ms / K,
                    "e2e": {"value": world * 2 * n * K / (hms / 1e3), "unit": "reads/s",
                            "h2d_bytes_per_step": int(2 * (s1.size + o1.size * 8)), "d2h_bytes_per_step": int(n + 24)},
                    "roofline_frac": ach / R.peaks["tflops"], "k2_tflops": ach, "flop_per_read": flop_per_read(L),
                    "pair_counts": counts.cpu().tolist()}
        del hs, ds
    seq, off = synth.synth_reads(args.reads_per_step // 2, 40, 300, synth.SEED_BASE + 5 + 1000 * rank)
    n5 = len(off) - 1
    hs = [pin(seq), pin(off)]
    ds = [t.to(R.dev) for t in hs]
    labels = {"labels": torch.empty(n5, dtype=torch.int8).pin_memory()}
    mean_steps = float(np.minimum(off[1:] - off[:-1], 300).mean())
    ms, timing, _ = R.time_device(lambda i: model.classify(ds[0], ds[1], 300), K, 3)
    hms, _ = R.time_host(lambda i: model.classify_host(hs[0], hs[1], 300, want_logits=False, out=labels), K, 1)
    ms, hms = R.max_over_ranks(ms, hms)
    lstm_ms, lstm_n = timing["lstm"]
    ach = flop_per_read(mean_steps) * n5 / (lstm_ms / max(lstm_n, 1) / 1e3) / 1e12
    out["C5"] = {"workload": "40-300 bp mixed single-end, -l 300 (BASELINE configs[4] shape, length-bucketed tiles): %d reads "
                             "per step per GPU, mean %.1f steps, %d steps" % (n5, mean_steps, K),
                 "value": world * n5 * K / (ms / 1e3), "unit": "reads/s", "ms_per_step": ms / K,
                 "read_steps_per_s": world * n5 * mean_steps * K / (ms / 1e3),
                 "e2e": {"value": world * n5 * K / (hms / 1e3), "unit": "reads/s",
                         "h2d_bytes_per_step": int(seq.size + off.size * 8), "d2h_bytes_per_step": int(n5 + 24)},
                 "roofline_frac": ach / R.peaks["tflops"], "k2_tflops": ach, "flop_per_read": flop_per_read(mean_steps)}
    # the README-recommended -l 170 for the same reads: longer reads are cut to their first 170 bases
    mean170 = float(np.minimum(off[1:] - off[:-1], 170).mean())
    ms, timing, _ = R.time_device(lambda i: model.classify(ds[0], ds[1], 170), K, 3)
    ms, = R.max_over_ranks(ms)
    lstm_ms, lstm_n = timing["lstm"]
    ach = flop_per_read(mean170) * n5 / (lstm_ms / max(lstm_n, 1) / 1e3) / 1e12
    out["C5_l170"] = {"workload": "the C5 reads with -l 170: mean %.1f steps, %d steps, device-resident only" % (mean170, K),
                      "value": world * n5 * K / (ms / 1e3), "unit": "reads/s", "ms_per_step": ms / K,
                      "roofline_frac": ach / R.peaks["tflops"], "k2_tflops": ach, "flop_per_read": flop_per_read(mean170)}
    return out


def run_strong(R):
    """A fixed set of read pairs in ONE shared host buffer (/dev/shm), cut by shard_bounds over the ranks."""
    from ribodetector_b200 import shard
    from ribodetector_b200.utils import synth
    torch, model, args, world, rank = R.torch, R.model, R.args, R.world, R.rank
    P, L = args.strong_pairs, READ_LEN
    if P <= 0:
        return None
    # the shared source needs 2 * P * L bytes of /dev/shm: shrink the set (every rank takes the same decision) if the box
    # has less, skip the block if it has next to none
    import shutil
    try:
        free = shutil.disk_usage("/dev/shm").free if rank == 0 else 0
    except OSError:
        free = 0
    (free,) = R.max_over_ranks(float(free))
    fit = int((free - (1 << 30)) // (2 * L))
    if fit < (1 << 20):
        return {"skipped": "less than %.1f GB free in /dev/shm for the shared source buffer" % ((2 * L * (1 << 20) + (1 << 30)) / 1e9)} if rank == 0 else None
    P = min(P, fit)
    nbytes = P * L
    tag = "rd_b200_strong_%s" % os.environ.get("MASTER_PORT", str(os.getpid()))
    paths = [os.path.join("/dev/shm", "%s_r%d.npy" % (tag, e)) for e in (1, 2)]
    t_gen = time.perf_counter()
    try:
        if rank == 0:
            for e, p in enumerate(paths):
                m = np.lib.format.open_memmap(p, mode="w+", dtype=np.uint8, shape=(nbytes,))
                synth.fill_bases(m, synth.SEED_BASE + 70 + e, threads=min(16, os.cpu_count() or 1))
                m.flush()
                del m
        R.barrier()
        src = [np.load(p, mmap_mode="r+") for p in paths]
        t_gen = time.perf_counter() - t_gen
        b, e = shard.shard_bounds(P, rank, world)
        cudart = torch.cuda.cudart()

        def view(lo, hi):
            """page-locked tensors over pairs [lo, hi) of the shared buffer + their offsets"""
            ts = []
            for s in src:
                t = torch.from_numpy(s[lo * L:hi * L])
                rc = cudart.cudaHostRegister(t.data_ptr(), t.numel(), 0) if t.numel() else 0
                ts.append((t, int(rc) == 0))
            off = torch.arange(hi - lo + 1, dtype=torch.int64).mul_(L).pin_memory()
            return ts, off

        def unpin(ts):
            for t, ok in ts:
                if ok and t.numel():
                    cudart.cudaHostUnregister(t.data_ptr())

        ts, off = view(b, e)
        m = e - b
        out = {"labels": torch.empty(m, dtype=torch.int8).pin_memory()}
        warm = min(m, 1 << 20)                        # a short untimed pass sizes the library's stage buffers
        if warm:
            model.classify_pairs_host(ts[0][0][:warm * L], off[:warm + 1], ts[1][0][:warm * L], off[:warm + 1], L, mode="rrna")
        R.barrier()
        t0 = time.perf_counter()
        r = model.classify_pairs_host(ts[0][0], off, ts[1][0], off, L, mode="rrna", out=out) if m else None
        torch.cuda.synchronize(R.dev)
        ms = (time.perf_counter() - t0) * 1e3
        my_ms = ms
        (ms,) = R.max_over_ranks(ms)
        (min_ms,) = [-v for v in R.max_over_ranks(-my_ms)]
        counts = (r["counts"] if r is not None else torch.zeros(3, dtype=torch.int64)).to(R.dev)
        shard.allreduce_counts(counts)                                  # NCCL: int64[3]
        gathered = shard.gather_labels(out["labels"].to(R.dev), P) if world > 1 else out["labels"]
        pinned = all(ok for _, ok in ts)
        unpin(ts)
        res = None
        if rank == 0:
            res = {"workload": "%d read pairs of %d bp (-e rrna) from ONE source buffer in shared host memory (/dev/shm, "
                               "generated once by rank 0, page-locked by every rank over its own shard), shard_bounds over "
                               "%d rank(s), rd_classify_pairs_host per rank" % (P, L, world),
                   "pairs": P, "reads_per_s": 2.0 * P / (ms / 1e3), "ms": ms, "n_gpus": world,
                   "host_buffers_page_locked": pinned, "generate_s": t_gen,
                   "pair_counts_allreduced": counts.cpu().tolist()}
            if world > 1:
                # one GPU over the whole set, same box, same run: the N=1 time and the labels to compare with
                ts1, off1 = view(0, P)
                out1 = {"labels": torch.empty(P, dtype=torch.int8).pin_memory()}
                t0 = time.perf_counter()
                r1 = model.classify_pairs_host(ts1[0][0], off1, ts1[1][0], off1, L, mode="rrna", out=out1)
                torch.cuda.synchronize(R.dev)
                ms1 = (time.perf_counter() - t0) * 1e3
                unpin(ts1)
                res.update({"n1_reads_per_s_same_run": 2.0 * P / (ms1 / 1e3), "n1_ms": ms1,
                            "efficiency_vs_n1": ms1 / (world * ms),
                            "rank_ms_min_max": [min_ms, ms], "overhead_ms_vs_n1_over_N": ms - ms1 / world,
                            "limiter": "fixed cost per rd_classify_pairs_host call (first 2^18-read chunk's H2D and the last chunk's "
                                       "D2H are not hidden behind kernels, pipeline fill/drain, last tiles of a shard partly "
                                       "filled): it does not shrink with the shard.  Host feed is far from a limit: %.1f GB/s of "
                                       "H2D over all ranks" % (P * (2 * L + 16) / (ms / 1e3) / 1e9),
                            "labels_equal_n1_bit_for_bit": bool(torch.equal(gathered.cpu(), out1["labels"])),
                            "counts_equal_n1": bool(r1["counts"].tolist() == counts.cpu().tolist())})
            else:
                res.update({"efficiency_vs_n1": 1.0, "labels_equal_n1_bit_for_bit": True,
                            "note": "N = 1: the sharded run IS the one-GPU run"})
        R.barrier()
        return res
    finally:
        if rank == 0:
            time.sleep(0.2)
            for p in paths:
                try:
                    os.unlink(p)
                except OSError:
                    pass


def run_ours(args, rank, world, local_rank):
    R = Runner(args, rank, world, local_rank)
    torch, dist, model, dev = R.torch, R.dist, R.model, R.dev
    from ribodetector_b200 import shard
    from ribodetector_b200.utils import synth

    def pin(a):
        return torch.from_numpy(a).pin_memory()

    n = args.reads_per_step
    nbuf = 2                                        # rotate distinct batches between steps
    host, devb = [], []
    for b in range(nbuf):
        seq, off = synth.synth_reads_fixed(n, READ_LEN, synth.SEED_BASE + 2 + 1000 * rank + b)
        hs, ho = pin(seq), pin(off)
        host.append((hs, ho))
        devb.append((hs.to(dev), ho.to(dev)))
    counts = torch.zeros(3, dtype=torch.int64, device=dev)
    out_host = {"labels": torch.empty(n, dtype=torch.int8).pin_memory()}
    model._lib.rd_reserve(model._need(), n, READ_LEN)

    def step_device(i):
        s, o = devb[i % nbuf]
        return model.classify(s, o, READ_LEN, counts=counts)

    # ---- device-resident: value ---------------------------------------------------------------------
    log("inputs ready; timing the device-resident path (%s)" % args.precision)
    for i in range(args.warmup):
        step_device(i)
    R.barrier()
    model.set_timing(True)
    model.get_timing(reset=True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = model.kernel_launches()
    counts.zero_()
    R.barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step_device(i)
    shard.allreduce_counts(counts)                  # the path's only exchange: int64[3] label counts (no-op at N=1)
    e1.record()
    R.barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = model.kernel_launches() - launches0
    clocks = sampler.stop() if rank == 0 else None
    timing = model.get_timing(reset=True)
    model.set_timing(False)
    total_counts = counts.cpu().tolist()

    # ---- the other precisions, device-resident, for information (same timing protocol) -----------------
    def time_mode(prec):
        ms, _, _ = R.time_device(lambda i: model.classify(devb[i % nbuf][0], devb[i % nbuf][1], READ_LEN, precision=prec),
                                 args.steps, 3)
        (ms,) = R.max_over_ranks(ms)
        return {"precision": prec, "value": world * n * args.steps / (ms / 1000.0), "unit": "reads/s",
                "ms_per_step": ms / args.steps,
                "roofline_frac_of_step": flop_per_read(READ_LEN) * n * args.steps / (ms / 1e3) / 1e12 / R.peaks["tflops"]}

    other = {}
    log("value done (%.1f ms per step); other precisions" % (dev_ms / args.steps))
    if not args.no_fast:
        notes = {"tc_fast": "single fp16 pass + tanh.approx: |dlogit| <= 5e-2, ~0.01-0.1 % label flips vs fp32",
                 "tc_auto": "tc_fast over all reads + tc_exact over the ~1 % with |margin| < 0.25: labels identical to "
                            "tc_exact, logits exact-grade only inside the band",
                 "tc_exact": "3-pass fp16 split (25 MMAs per chunk): |dlogit| <= 2e-4, the arbiter mode",
                 "tc_mixed": "fp16 main pass + e5m2 correction pass (17 MMAs per chunk) + tc_exact over the low-margin band"}
        for prec in ("tc_exact", "tc_mixed", "tc_fast", "tc_auto"):
            if prec != args.precision:
                other[prec] = time_mode(prec)
                other[prec]["note"] = notes[prec]

    # ---- L-kernel as SURVEY 8d words it: five separately timed launches of the K2 stage, best and median ----------
    k2_single = []
    for rep in range(5):
        _, tm1, _ = R.time_device(step_device, 1, 0)
        k2_single.append(tm1["lstm"][0] / max(tm1["lstm"][1], 1))
    k2_single = [R.max_over_ranks(v)[0] for v in k2_single]

    # ---- end to end: host buffers through the public API ----------------------------------------------
    log("e2e (host buffers)")
    e2e_ms, r = R.time_host(lambda i: model.classify_host(host[i % nbuf][0], host[i % nbuf][1], READ_LEN, want_logits=False,
                                                          out=out_host), args.steps, min(args.warmup, 2))
    e2e_counts = r["counts"].tolist()
    # ---- the edges included: FASTQ TEXT in host memory -> label-partitioned record text in host memory ----------
    # (rd_fastq_submit / rd_fastq_collect: H2D, K0 record scan, K1-K3, K4 partition, D2H; two slots in flight)
    log("e2e_fastq (FASTQ text -> partitioned text)")
    fq_n = min(n, 1 << 21)
    fq_blocks = max(2, (args.steps * n) // (2 * fq_n))
    text = torch.from_numpy(synth.fastq_text(fq_n, READ_LEN, synth.SEED_BASE + 5 + rank)).pin_memory().numpy()
    fq_bytes = int(text.size)
    fq_out = [torch.empty(text.size + 32, dtype=torch.uint8).pin_memory().numpy() for _ in range(2)]
    fq_counts = np.zeros(3, np.int64)
    for rep in range(2):                            # the first pass sizes the slots and warms up
        R.barrier()
        t0 = time.perf_counter()
        for k in range(fq_blocks):
            if k >= 2:
                fq_counts += model.fastq_collect(k & 1)[1]
            got, _, _ = model.fastq_submit(k & 1, [text], [text.size], True, fq_n + 8, READ_LEN, [fq_out[k & 1]])
            assert got == fq_n
        for k in range(2):
            fq_counts += model.fastq_collect(k)[1]
        fq_ms = (time.perf_counter() - t0) * 1000.0
    dev_ms, e2e_ms, fq_ms = R.max_over_ranks(dev_ms, e2e_ms, fq_ms)
    h2d = host[0][0].numel() + host[0][1].numel() * 8
    d2h = n + 24
    del text, fq_out

    # ---- parity of the headline precision, outside the timed region ---------------------------------------
    log("parity vs the fp32 kernel")
    n_par = min(n, 8192)
    ps, po = host[0][0][:n_par * READ_LEN], host[0][1][:n_par + 1]
    got_par = model.classify(ps, po, READ_LEN)[0].cpu().numpy()
    ref32 = model.classify(ps, po, READ_LEN, precision="fp32")[0].cpu().numpy()
    parity_fp32 = parity_block(got_par, ref32, args.precision, vs="the on-device fp32 CUDA-core kernel (packed semantics) on the "
                               "first %d reads of the timed batch" % n_par)

    def guarded(name, fn):
        """the extra blocks must not cost the headline line: an exception is reported in the block's place"""
        try:
            return fn()
        except Exception as ex:             # noqa: BLE001
            import traceback
            traceback.print_exc(file=sys.stderr)
            return {"error": "%s: %s" % (type(ex).__name__, ex)}

    log("configs C3/C4/C5")
    configs = guarded("configs", lambda: run_configs(R, pin)) if not args.no_configs else None
    log("strong (fixed single-source set)")
    strong = guarded("strong", lambda: run_strong(R)) if not args.no_strong else None
    log("GPU phases done")

    if rank == 0:
        peaks = R.peaks
        lstm_ms, lstm_n = timing["lstm"]
        lstm_avg_s = (lstm_ms / max(lstm_n, 1)) / 1000.0
        fpr = flop_per_read(READ_LEN)
        achieved = fpr * n / lstm_avg_s / 1e12 if lstm_avg_s > 0 else 0.0
        value = world * n * args.steps / (dev_ms / 1000.0)
        e2e = world * n * args.steps / (e2e_ms / 1000.0)
        bytes_per_read, traffic_note = k2_traffic(args.precision, READ_LEN)
        mufu = MUFU_PER_UNIT_STEP[args.precision] * 128 * READ_LEN
        line = {
            "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": DTYPE[args.precision], "data": "synthetic",
            "config": {"workload": "100 bp single-end, 50M synthetic reads per B200 (BASELINE configs[1]): "
                                   "timed as %d batches of %d reads on each of %d GPU(s)" % (args.steps, n, world),
                       "read_len": READ_LEN, "reads_per_step_per_gpu": n, "precision": args.precision,
                       "tolerance": "|dlogit| <= %.0e, |dprob| <= %.0e, labels identical outside |margin| <= %.0e "
                                    "(tests/test_gpu_parity.py)" % TOL[args.precision],
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
                         "traffic": bytes_per_read * n if bytes_per_read is not None else None,
                         "traffic_note": traffic_note + "; algorithmic bytes per launch = %d (sequence + offset + slot "
                                         "plan/perm read, logits written)" % (n * (READ_LEN + 8 + 8 + 8)),
                         "executed_tflops": achieved * EXECUTED_PER_ALGORITHMIC[args.precision],
                         "executed_frac": achieved * EXECUTED_PER_ALGORITHMIC[args.precision] / peaks["tflops"],
                         "peak_source": peaks["source"], "launch_ms": lstm_avg_s * 1000.0,
                         "flop_per_launch": fpr * n,
                         "launch_ms_best_of_5": min(k2_single), "launch_ms_median_of_5": float(np.median(k2_single)),
                         "frac_best_of_5": fpr * n / (min(k2_single) / 1e3) / 1e12 / peaks["tflops"],
                         "share_of_step": lstm_ms / dev_ms if dev_ms else None,
                         "co_bound": {"pipe": "xu (MUFU ex2/rcp/tanh)", "ops_per_read": mufu,
                                      "achieved_gops": mufu * n / lstm_avg_s / 1e9 if lstm_avg_s > 0 else 0.0,
                                      "peak_gops": XU_LANES_PER_CLK_PER_SM * 148 * ((clocks or {}).get("sm_mhz") or 1965.0) / 1e3,
                                      "note": "peak = 16 MUFU lanes/clk/SM (measured) x 148 SMs x SM clock under load"}},
            "e2e_fastq": {"value": world * fq_n * fq_blocks / (fq_ms / 1000.0), "unit": "reads/s",
                          "text_gb_per_s_each_way": world * fq_bytes * fq_blocks / (fq_ms / 1000.0) / 1e9,
                          "h2d_bytes_per_block": fq_bytes, "d2h_bytes_per_block": fq_bytes + 1 + 72,
                          "reads_per_block": fq_n, "blocks": fq_blocks,
                          "note": "FASTQ text in page-locked host memory -> record scan (K0), classify, label partition (K4) on the "
                                  "GPU -> partitioned record text back in host memory (rd_fastq_submit / rd_fastq_collect)"},
            "stage_ms": {k: v[0] / max(v[1], 1) for k, v in timing.items() if v[1]},
            "label_counts": total_counts, "e2e_label_counts_last_step": e2e_counts,
            "parity_fp32_kernel": parity_fp32,
        }
        cb = line["roofline"]["co_bound"]
        cb["frac"] = cb["achieved_gops"] / cb["peak_gops"] if cb["peak_gops"] else None
        if other:
            line["other_precisions"] = other
            line["fast_mode"] = other.get("tc_fast")
            line["auto_mode"] = other.get("tc_auto")
        if configs:
            line["configs"] = configs
        if strong:
            line["strong"] = strong
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            n_cpu = min(n, threads * 1024 * args.cpu_batches)
            cs, co = host[0][0][:n_cpu * READ_LEN].numpy(), host[0][1][:n_cpu + 1].numpy()
            log("cpu_baseline on %d cores, %d reads" % (threads, n_cpu))
            info, cpu_logits, _ = cpu_arm(R.weights, threads, cs, co)
            log("cpu_baseline done")
            info["sample"] = "the first " + info["sample"] + " of the timed batch"
            line["cpu_baseline"] = info
            # ribodetector_cpu = padded semantics (model_cpu.py:57-62): compare like with like
            got = model.classify(cs, co, READ_LEN, semantics="padded")[0].cpu().numpy()
            line["parity"] = parity_block(got, cpu_logits, args.precision,
                                          vs="the CPU arm's fp32 logits (kind %s; padded semantics = ribodetector_cpu) on the reads "
                                             "it classified" % info["kind"])
        else:
            line["parity"] = dict(parity_fp32, note="no CPU arm in this run (N > 1 or --no-cpu-baseline): this is parity_fp32_kernel")
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
    ap.add_argument("--precision", default=os.environ.get("RD_BENCH_PRECISION", "tc_mixed"),
                    choices=["fp32", "tc_exact", "tc_fast", "tc_auto", "tc_mixed", "tc_mixed_raw"])
    ap.add_argument("--reads-per-step", type=int, default=BATCH_READS)
    ap.add_argument("--cpu-batches", type=int, default=20, help="1024-read batches per CPU worker in cpu_baseline")
    ap.add_argument("--config-steps", type=int, default=3, help="timed steps of each BASELINE configs[2..4] entry")
    ap.add_argument("--strong-pairs", type=int, default=25 << 20, help="read pairs of the fixed single-source set")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fast", action="store_true", help="skip the informational timing of the other precisions")
    ap.add_argument("--no-configs", action="store_true")
    ap.add_argument("--no-strong", action="store_true")
    ap.add_argument("--watchdog", type=int, default=1500, help="seconds after which every thread's stack is dumped to "
                                                                "stderr and the run exits (0 = off)")
    args = ap.parse_args()
    if args.watchdog > 0:
        import faulthandler
        faulthandler.dump_traceback_later(args.watchdog, exit=True)
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
