#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2e
mkdir -p "$OUT"
timeout 600 python -m pytest tests/test_fp8.py tests/test_native_gpu.py -q -m gpu -rA -p no:cacheprovider > "$OUT/pytest.log" 2>&1; tail -n 2 "$OUT/pytest.log"; grep -E "^(FAILED|ERROR)" "$OUT/pytest.log" | head
echo "=== fp8"; timeout 200 python tools/gpu_time_fp8.py 2>&1 | tee "$OUT/fp8.jsonl" | tail -n 5
for pe in 2 3 4; do echo "=== D=64 poly $pe"; LCA_B200_POLY_EVERY=$pe S=32768 D=64 H=16 timeout 100 python tools/gpu_time_passes.py 2>&1 | head -n 1; done
echo "=== peers"; timeout 300 python tools/gpu_time_sdpa_peers.py 2>&1 | tee "$OUT/peers.jsonl" | head -n 4
echo "=== bench N=1"; timeout 400 python bench.py --steps 5 --warmup 3 > "$OUT/bench_n1.json" 2> "$OUT/bench_n1.err"; cut -c1-400 "$OUT/bench_n1.json"; tail -n 3 "$OUT/bench_n1.err"
