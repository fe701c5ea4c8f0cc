#!/usr/bin/env bash
# First GPU call of a round: validate the opt-in (template-gated) kernel variants that were written without hardware
# and measure them against the default build.  One GPU, ~30 minutes.  Everything lands in gpurun_out/experimental/.
#
#   gpurun --timeout 2700 -- 'bash tools/validate_experimental.sh'
#
# Each step runs under its own timeout so a hung kernel cannot eat the box.  Exit code = number of failed steps.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/experimental
mkdir -p "$OUT"
fail=0
step() {   # step <name> <timeout_s> <env...> -- <cmd...>
  local name=$1 tmo=$2; shift 2
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  echo "=== $name (${envs[*]:-default})"
  if env ${envs[@]+"${envs[@]}"} timeout "$tmo" "$@" > "$OUT/$name.log" 2>&1; then
    echo "    ok"; tail -n 3 "$OUT/$name.log" | sed 's/^/    /'
  else
    echo "    FAILED (exit $?)"; tail -n 15 "$OUT/$name.log" | sed 's/^/    /'; fail=$((fail + 1))
  fi
}

# 0. reference point: default build
step perf_default 120 S=32768 -- python tools/gpu_time_passes.py

# 0b. the round-1 hang and its fix (dQ pass, empty work items): the fixed kernel must finish and match the oracle;
#     (the pre-fix reproduction, which is EXPECTED to hang, runs as the very last step of this script)
step xfix_fixed 180 -- python tools/gpu_repro_xfix.py
# 1. packed fp32x2 softmax / dS arithmetic (FFMA2 / FADD2 / FMUL2)
step tests_f32x2 420 LCA_B200_F32X2=1 -- python -m pytest tests/test_native_gpu.py -x -q -m gpu
step perf_f32x2 120 LCA_B200_F32X2=1 S=32768 -- python tools/gpu_time_passes.py
for pe in 2 3 4; do
  step perf_f32x2_poly$pe 120 LCA_B200_F32X2=1 LCA_B200_POLY_EVERY=$pe S=32768 -- python tools/gpu_time_passes.py
done
for pe in 2 3 6; do          # head_dim 64: exponentials outnumber the tensor work 2:1, needs the heaviest offload
  step perf_d64_f32x2_poly$pe 120 LCA_B200_F32X2=1 LCA_B200_POLY_EVERY=$pe S=32768 D=64 H=16 -- python tools/gpu_time_passes.py
done
step perf_d64_default 120 S=32768 D=64 H=16 -- python tools/gpu_time_passes.py

# 1a. forward with 64-row K/V tiles and double-buffered scores (separate kernel file, scalar softmax arithmetic)
step tests_bn64 420 LCA_B200_FWD_BN64=1 -- python -m pytest tests/test_native_gpu.py -x -q -m gpu -k "fwd or module or padded or varlen or util"
step perf_bn64 120 LCA_B200_FWD_BN64=1 S=32768 -- python tools/gpu_time_passes.py
step perf_bn64_poly3 120 LCA_B200_FWD_BN64=1 LCA_B200_POLY_EVERY=3 S=32768 -- python tools/gpu_time_passes.py
step tests_bn64_f32x2 420 LCA_B200_FWD_BN64=1 LCA_B200_F32X2=1 -- python -m pytest tests/test_native_gpu.py -x -q -m gpu -k "fwd or module or padded or varlen"
step perf_bn64_f32x2 120 LCA_B200_FWD_BN64=1 LCA_B200_F32X2=1 S=32768 -- python tools/gpu_time_passes.py
step perf_bn64_f32x2_poly3 120 LCA_B200_FWD_BN64=1 LCA_B200_F32X2=1 LCA_B200_POLY_EVERY=3 S=32768 -- python tools/gpu_time_passes.py
step perf_bn64_d64 120 LCA_B200_FWD_BN64=1 LCA_B200_POLY_EVERY=3 S=32768 D=64 H=16 -- python tools/gpu_time_passes.py

# 1b. backward with both element-wise warpgroups on every streamed tile (halves the per-tile critical path)
step tests_split 420 LCA_B200_BWD_SPLIT=1 -- python -m pytest tests/test_native_gpu.py -x -q -m gpu -k "bwd or backward or padded or module"
step perf_split 120 LCA_B200_BWD_SPLIT=1 S=32768 -- python tools/gpu_time_passes.py
step tests_split_f32x2 420 LCA_B200_BWD_SPLIT=1 LCA_B200_F32X2=1 -- python -m pytest tests/test_native_gpu.py -x -q -m gpu -k "bwd or backward or padded or module"
step perf_split_f32x2 120 LCA_B200_BWD_SPLIT=1 LCA_B200_F32X2=1 S=32768 -- python tools/gpu_time_passes.py

# 2. dynamic tile scheduler
step tests_dyn 420 LCA_B200_DYN_SCHED=1 -- python -m pytest tests/test_native_gpu.py -x -q -m gpu
step perf_dyn 120 LCA_B200_DYN_SCHED=1 S=32768 -- python tools/gpu_time_passes.py

# 3. fp8 forward (tcgen05 kind::f8f6f4)
step tests_fp8 300 LCA_B200_EXPERIMENTAL_FP8=1 -- python -m pytest tests/test_fp8.py -x -q -m gpu

# 4. native dropout instantiations (coordinate-keyed keep mask regenerated in registers)
step tests_dropout 300 LCA_B200_NATIVE_DROPOUT=1 -- python -m pytest tests/test_dropout_gpu.py -x -q -m gpu

# 5. CUDA-graph capture of the single-GPU forward + backward
step tests_graphs 240 LCA_B200_TEST_GRAPHS=1 -- python -m pytest tests/test_cuda_graph_gpu.py -x -q -m gpu

grep -h '"name"' "$OUT"/perf_*.log 2>/dev/null | sed 's/^/  /' > "$OUT/summary.txt"
for f in "$OUT"/perf_*.log; do echo "$(basename "$f" .log): $(grep -h '"name"' "$f" | python -c '
import sys, json
print("  ".join("%s %.3f ms" % (d["name"].split("(")[0], d["ms"]) for d in map(json.loads, sys.stdin)))')"; done | tee -a "$OUT/summary.txt"
# last on purpose: a kernel that is expected to hang is killed after 40 s; nothing else depends on the GPU afterwards
echo "=== xfix_prefix_repro (LCA_B200_NO_XFIX=1; a timeout here CONFIRMS the diagnosis)"
if LCA_B200_NO_XFIX=1 timeout -s KILL 40 python tools/gpu_repro_xfix.py --no-check > "$OUT/xfix_prefix_repro.log" 2>&1; then
  echo "    pre-fix kernel finished this time (the hazard is timing dependent)"
else
  echo "    pre-fix kernel did not finish (exit $?): $(tail -n 1 "$OUT/xfix_prefix_repro.log")"
fi
echo "failed steps: $fail"
exit $fail
