#!/usr/bin/env bash
# Build librd_b200.so in-tree for sm_100a.  Usage: csrc/build.sh [extra nvcc flags]
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../librd_b200.so
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
      -Xcompiler -fPIC -shared -Xptxas -v "$@" \
      rd_api.cu rd_plan.cu rd_lstm_fp32.cu rd_lstm_tc.cu rd_tail.cu rd_fastx.cu rd_fastq_dev.cu -o $OUT -lz 2> build.log || { cat build.log; exit 1; }
grep -E "error|warning|spill|registers" build.log | grep -v "^$" | head -60 || true
ls -la $OUT
