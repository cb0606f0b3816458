"""Summarise an .ncu-rep (ncu --set full) into the text form kept under profiles/:
python tools/ncu_summary.py report.ncu-rep [kernel-name-substring] > profiles/<name>_summary.txt"""
import csv
import subprocess
import sys

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


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    text = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(text.splitlines()))
    hdr, units = rows[0], rows[1]
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
