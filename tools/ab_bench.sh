#!/usr/bin/env bash
# Same-box A/B of library builds: tools/ab_bench.sh name1 name2 ... (tools/librd_ab_<name>.so), two interleaved rounds.
# Prints value / fast / auto reads/s and the SM clock of every run.
set -u
for round in 1 2; do
  for name in "$@"; do
    RD_B200_LIB=$PWD/tools/librd_ab_$name.so python bench.py --no-cpu-baseline --no-configs --no-strong --steps 12 2>/dev/null | tail -1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$name', d['config']['precision'], '%.2fM' % (d['value']/1e6), 'fast %.2fM' % (d['fast_mode']['value']/1e6), 'auto %.2fM' % (d['auto_mode']['value']/1e6), 'sm_mhz', d['clocks']['sm_mhz'], 'W', d['clocks']['power_w_max'])"
  done
done
