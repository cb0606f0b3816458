#!/usr/bin/env bash
# Round-2 evidence run on ONE B200 (gpurun): ncu captures of the LSTM kernel variants, the launch list of the bench's
# own step, compute-sanitizer synccheck / racecheck over smoke().  Outputs under gpurun_out/.
set -u
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:lstm_tc -s 2 -c 1 -f -o $O/r2_ncu_mixed_4m   python tools/ncu_k2.py tc_mixed_raw 4194304 100   > $O/r2_ncu_a.log 2>&1
$NCU -k regex:lstm_tc -s 2 -c 1 -f -o $O/r2_ncu_mixed_150  python tools/ncu_k2.py tc_mixed_raw 2097152 150   > $O/r2_ncu_b.log 2>&1
$NCU -k regex:lstm_tc -s 2 -c 1 -f -o $O/r2_ncu_mixed_40_300 python tools/ncu_k2.py tc_mixed_raw 2097152 40-300 > $O/r2_ncu_c.log 2>&1
$NCU -k regex:lstm_tc -s 5 -c 1 -f -o $O/r2_ncu_mixed_pass2 python tools/ncu_k2.py tc_mixed 4194304 100       > $O/r2_ncu_d.log 2>&1
$NCU -k regex:lstm_tc -s 5 -c 1 -f -o $O/r2_ncu_auto_pass2  python tools/ncu_k2.py tc_auto 4194304 100        > $O/r2_ncu_e.log 2>&1
$NCU -k regex:lstm_tc -s 2 -c 1 -f -o $O/r2_ncu_exact_4m   python tools/ncu_k2.py tc_exact 4194304 100        > $O/r2_ncu_f.log 2>&1
$NCU -k regex:lstm_tc -s 2 -c 1 -f -o $O/r2_ncu_fast_4m    python tools/ncu_k2.py tc_fast 4194304 100         > $O/r2_ncu_g.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_tc_mixed.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fast --no-configs --no-strong > $O/r2_launch_bench.log 2>&1
timeout 600 compute-sanitizer --tool synccheck python __graft_entry__.py smoke > $O/r2_sanitizer_synccheck.log 2>&1
timeout 600 compute-sanitizer --tool racecheck python __graft_entry__.py smoke > $O/r2_sanitizer_racecheck.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python __graft_entry__.py smoke > $O/r2_sanitizer_memcheck.log 2>&1
tail -3 $O/r2_sanitizer_synccheck.log $O/r2_sanitizer_racecheck.log $O/r2_sanitizer_memcheck.log
ls -la $O/*.ncu-rep
