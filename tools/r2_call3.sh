#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2d
mkdir -p "$OUT"
for pe in 0 3 6; do echo "=== poly $pe"; LCA_B200_POLY_EVERY=$pe S=32768 timeout 100 python tools/gpu_time_passes.py 2>&1 | head -n 1; done
S=32768 OUT=$OUT/ncu bash tools/ncu_capture.sh 2>&1 | tail -n 4
