#!/usr/bin/env bash
# quick one-GPU check after a kernel change: native GPU tests + per-pass timings
set -u
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_native_gpu.py tests/test_dropout_gpu.py tests/test_cuda_graph_gpu.py -x -q -m gpu 2>&1 | tail -n 5
for cfg in "S=32768" "S=32768 LCA_B200_DKV_BY=64" "S=32768 D=64 H=16" "S=32768 D=64 H=16 LCA_B200_DKV_BY=64" "S=8192 H=32" "S=32768 CAUSAL=0 H=4"; do
  echo "=== perf $cfg"; env $cfg timeout 200 python tools/gpu_time_passes.py 2>&1 | tail -n 3 | sed 's/^/   /'
done
