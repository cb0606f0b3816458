"""Summarise an .ncu-rep (ncu --set full) into the text form kept under profiles/:
    python tools/ncu_summary.py report.ncu-rep [kernel-name-substring] > profiles/<name>_summary.txt
and record the DRAM traffic of the LSTM kernel for bench.py's roofline.traffic:
    python tools/ncu_summary.py report.ncu-rep lstm_tc --traffic tc_mixed 100 4194304
adds {"<precision>_L<read_len>": {bytes_per_read, reads, report, src_sha}} to profiles/k2_traffic.json, where src_sha is
the sha256 of csrc/rd_lstm_tc.cu at capture time — bench.py reports traffic = null once the kernel source has changed.
(Runs where ncu is installed; no GPU needed.)"""
import csv
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

METRICS = [
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__time_duration.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__block_size",
    "launch__cluster_size", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "sm__cycles_elapsed.avg.per_second", "sm__cycles_elapsed.max", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


def record_traffic(rows, hdr, want, rep, precision, read_len, reads):
    """dram bytes of the LAST matching launch (the first ones of a run may be short warm-up launches)."""
    last = None
    for r in rows[2:]:
        if want in r[hdr.index("Kernel Name")]:
            last = r
    if last is None:
        raise SystemExit("no kernel matching %r in %s" % (want, rep))
    units = rows[1]

    def mbytes(metric):
        i = hdr.index(metric)
        v = float(last[i].replace(",", ""))
        return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
    total = mbytes("dram__bytes_read.sum") + mbytes("dram__bytes_write.sum")
    with open(os.path.join(ROOT, "ribodetector_b200", "csrc", "rd_lstm_tc.cu"), "rb") as f:
        sha = hashlib.sha256(f.read()).hexdigest()
    path = os.path.join(ROOT, "profiles", "k2_traffic.json")
    d = {}
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
    d["%s_L%d" % (precision, read_len)] = {
        "bytes_per_read": total / reads, "dram_bytes_read": mbytes("dram__bytes_read.sum"),
        "dram_bytes_written": mbytes("dram__bytes_write.sum"), "reads": reads, "read_len": read_len,
        "precision": precision, "report": os.path.basename(rep), "src_sha": sha}
    with open(path, "w") as f:
        json.dump(d, f, indent=1, sort_keys=True)
        f.write("\n")
    print("profiles/k2_traffic.json: %s_L%d = %.1f B/read (%.1f MB per launch of %d reads)"
          % (precision, read_len, total / reads, total / 1e6, reads), file=sys.stderr)


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else ""
    text = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(text.splitlines()))
    hdr, units = rows[0], rows[1]
    if "--traffic" in sys.argv:
        i = sys.argv.index("--traffic")
        record_traffic(rows, hdr, want, rep, sys.argv[i + 1], int(sys.argv[i + 2]), int(sys.argv[i + 3]))
        return
    seen = set()
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if want not in name or name in seen:
            continue
        seen.add(name)
        print("Kernel Name ", name)
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(m, units[i], r[i])
        print()


if __name__ == "__main__":
    main()
