#!/usr/bin/env bash
# quick one-GPU check after a kernel change: native GPU tests + per-pass timings
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/quick
mkdir -p "$OUT"
timeout 900 python -m pytest tests -q -m gpu -rA -p no:cacheprovider > "$OUT/pytest_gpu.log" 2>&1; tail -n 3 "$OUT/pytest_gpu.log"; grep -E "^(FAILED|ERROR)" "$OUT/pytest_gpu.log" | head
: > "$OUT/passes.jsonl"
for cfg in "S=32768" "S=32768 D=64 H=16" "S=8192 H=32" "S=32768 CAUSAL=0 H=4" "S=131072 N=3" ${EXTRA_CFGS:-}; do
  echo "=== perf $cfg"; echo "# $cfg" >> "$OUT/passes.jsonl"; env $cfg timeout 200 python tools/gpu_time_passes.py 2>&1 | tail -n 3 | tee -a "$OUT/passes.jsonl" | sed 's/^/   /'
done
