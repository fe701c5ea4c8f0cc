#!/usr/bin/env bash
# Round-2 one-GPU validation of the final kernel set: every single-GPU test (named, -rA), per-pass timings, fp8 forward,
# library peers (cuDNN SDPA, flash-attn 2), headline bench at N=1, ncu captures (forward, dQ pass, dK/dV pass, fused push).
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2c
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader > "$OUT/gpu.txt" 2>&1
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu -rA -p no:cacheprovider > "$OUT/pytest_gpu.log" 2>&1; tail -n 3 "$OUT/pytest_gpu.log"
echo "=== fp8 tests"; LCA_B200_EXPERIMENTAL_FP8=1 timeout 300 python -m pytest tests/test_fp8.py -q -m gpu -rA -p no:cacheprovider > "$OUT/pytest_fp8.log" 2>&1; tail -n 2 "$OUT/pytest_fp8.log"
: > "$OUT/passes.jsonl"
for cfg in "S=32768" "S=32768 D=64 H=16" "S=131072 N=3" "S=8192 H=32" "S=32768 CAUSAL=0 H=4" "S=4096 H=32 N=20"; do
  echo "=== perf $cfg"; echo "# $cfg" >> "$OUT/passes.jsonl"; env $cfg timeout 200 python tools/gpu_time_passes.py 2>&1 | tail -n 3 | tee -a "$OUT/passes.jsonl"
done
echo "=== fp8"; timeout 200 python tools/gpu_time_fp8.py > "$OUT/fp8.jsonl" 2>&1; tail -n 8 "$OUT/fp8.jsonl"
echo "=== peers"; timeout 300 python tools/gpu_time_sdpa_peers.py > "$OUT/peers.jsonl" 2>&1; tail -n 8 "$OUT/peers.jsonl"
D=64 H=16 timeout 300 python tools/gpu_time_sdpa_peers.py > "$OUT/peers_d64.jsonl" 2>&1; tail -n 6 "$OUT/peers_d64.jsonl"
echo "=== bench N=1"; timeout 400 python bench.py --steps 5 --warmup 3 > "$OUT/bench_n1.json" 2> "$OUT/bench_n1.err"; cut -c1-1500 "$OUT/bench_n1.json"; tail -n 3 "$OUT/bench_n1.err"
echo "=== bench N=1 fwd"; timeout 400 python bench.py --steps 5 --warmup 3 --mode fwd > "$OUT/bench_n1_fwd.json" 2>> "$OUT/bench_n1.err"; cut -c1-600 "$OUT/bench_n1_fwd.json"
if [ "${NCU:-1}" = "1" ]; then
  echo "=== ncu"; S=32768 OUT=$OUT/ncu bash tools/ncu_capture.sh 2>&1 | tail -n 12
fi
