#!/usr/bin/env bash
# Round-2 one-GPU pass over the pruned kernels (packed arithmetic everywhere, x_empty count-9 everywhere): GPU tests,
# per-pass timings, fp8 forward, ncu captures of forward / dQ pass / dK-dV pass / a fused launch with push CTAs.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_n1
mkdir -p "$OUT"
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu -rA > "$OUT/pytest_gpu.log" 2>&1; tail -n 4 "$OUT/pytest_gpu.log"
grep -c PASSED "$OUT/pytest_gpu.log"
echo "=== fp8 tests"; LCA_B200_EXPERIMENTAL_FP8=1 timeout 300 python -m pytest tests/test_fp8.py -q -m gpu 2>&1 | tail -n 2
for cfg in "S=32768" "S=32768 D=64 H=16" "S=32768 LCA_B200_DYN_SCHED=1" "S=131072 N=3" "S=8192 H=32"; do
  echo "=== perf $cfg"; env $cfg timeout 200 python tools/gpu_time_passes.py 2>&1 | tail -n 3
done
echo "=== fp8"; timeout 200 python tools/gpu_time_fp8.py 2>&1 | tail -n 5
echo "=== bench N=1"; timeout 400 python bench.py --steps 5 --warmup 3 2>&1 | tail -n 1 | cut -c1-1200
if [ "${NCU:-1}" = "1" ]; then
  export S=16384 H=8 N=1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fmha_fwd -s 2 -c 1 -f -o $OUT/prof_fwd python tools/gpu_time_passes.py > $OUT/ncu_fwd.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fmha_bwd -s 4 -c 2 -f -o $OUT/prof_bwd python tools/gpu_time_passes.py > $OUT/ncu_bwd.log 2>&1
  S=8192 H=8 NCOMM=16 ITERS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fmha_fwd -s 1 -c 1 -f -o $OUT/prof_fused_push python tools/debug_push_1gpu.py > $OUT/ncu_push.log 2>&1
  ls -la $OUT | tail -6
fi
