#!/usr/bin/env bash
# Second GPU call of a round (N GPUs, default 2): multi-GPU validation of what cannot be checked on one GPU.
#
#   gpurun --gpus 2 --timeout 1500 -- 'bash tools/validate_multigpu.sh 2'
#   gpurun --gpus 8 --timeout 900 -- 'QUICK=1 FLAGS="..." bash tools/validate_multigpu.sh 8'     (charged 8x: keep it short)
#
# 1. fused test-suite for this GPU count (default build)                      -> tests_default.log
# 2. the same with the opt-in kernel variants that passed tools/validate_experimental.sh (pass them as FLAGS=...)
# 3. LCA_B200_SLAB=vmm (torch symmetric memory / CUDA VMM slab): tests, NVLS broadcast push on top of it
#    (LCA_B200_NVLS=1, pure-ring meshes), then the combination that hung once in
#    round 1 (fused forward + NCCL backward) under a short timeout
# 4. bench.py fwd+bwd and fwd at N GPUs, both arms, with the exposed-communication probe
# Everything lands in gpurun_out/multigpu_n$N/.  Every step has its own timeout.
set -u
cd "$(dirname "$0")/.."
N=${1:-2}
FLAGS=${FLAGS:-}                      # e.g. FLAGS="LCA_B200_F32X2=1 LCA_B200_BWD_SPLIT=1"
OUT=gpurun_out/multigpu_n$N
mkdir -p "$OUT"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
SEQ=$((N == 2 ? 131072 : 262144))
fail=0
step() {   # step <name> <timeout_s> <env...> -- <cmd...>
  local name=$1 tmo=$2; shift 2
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  echo "=== $name (${envs[*]:-default})"
  if env ${envs[@]+"${envs[@]}"} timeout "$tmo" "$@" > "$OUT/$name.log" 2>&1; then
    echo "    ok"; tail -n 2 "$OUT/$name.log" | cut -c1-400 | sed 's/^/    /'
  else
    echo "    FAILED (exit $?)"; tail -n 12 "$OUT/$name.log" | cut -c1-300 | sed 's/^/    /'; fail=$((fail + 1))
  fi
}

step tests_default 420 -- python -m pytest tests/test_fused_multigpu.py -x -q -k "${N}gpu or collective"
if [ -n "$FLAGS" ]; then
  # shellcheck disable=SC2086
  step tests_flags 420 $FLAGS -- python -m pytest tests/test_fused_multigpu.py -x -q -k "${N}gpu"
fi
if [ "${QUICK:-0}" != "1" ]; then     # QUICK=1 (use it at 8 GPUs: charged 8x): default tests + the bench arms only
step tests_fused_dropout 300 LCA_B200_NATIVE_DROPOUT=1 -- python -m pytest tests/test_fused_multigpu.py -x -q -k "dropout"
step tests_vmm 420 LCA_B200_SLAB=vmm -- python -m pytest tests/test_fused_multigpu.py -x -q -k "${N}gpu"
step tests_fastpush 420 LCA_B200_FAST_PUSH=1 -- python -m pytest tests/test_fused_multigpu.py -x -q -k "${N}gpu"
step bench_ours_fb_fastpush 300 LCA_B200_FAST_PUSH=1 -- $TR --master-port 29618 bench.py --gpus "$N" --steps 5 --warmup 3 --seq "$SEQ"
step bench_ours_fwd_fastpush_2cta 300 LCA_B200_FAST_PUSH=1 LCA_B200_COMM_CTAS=2 -- $TR --master-port 29619 bench.py --gpus "$N" --steps 5 --warmup 3 --seq "$SEQ" --mode fwd
step bench_ulysses_default 300 -- $TR --master-port 29620 bench.py --gpus "$N" --steps 10 --warmup 3 --seq 32768 --heads 32 --ulysses "$N" --ring-impl basic --mode fwd
step bench_ulysses_fastpush 300 LCA_B200_FAST_PUSH=1 -- $TR --master-port 29621 bench.py --gpus "$N" --steps 10 --warmup 3 --seq 32768 --heads 32 --ulysses "$N" --ring-impl basic --mode fwd
step tests_nvls 420 LCA_B200_SLAB=vmm LCA_B200_NVLS=1 -- python -m pytest tests/test_fused_multigpu.py -x -q -k "${N}gpu"
step bench_ours_fb_nvls 300 LCA_B200_SLAB=vmm LCA_B200_NVLS=1 -- $TR --master-port 29617 bench.py --gpus "$N" --steps 5 --warmup 3 --seq "$SEQ"
step hang_repro_ipc 150 LCA_B200_FUSED_BWD=0 -- $TR --master-port 29611 bench.py --gpus "$N" --steps 3 --warmup 3 --seq "$SEQ" --no-comm-probe
step hang_repro_vmm 150 LCA_B200_FUSED_BWD=0 LCA_B200_SLAB=vmm -- $TR --master-port 29612 bench.py --gpus "$N" --steps 3 --warmup 3 --seq "$SEQ" --no-comm-probe
fi
step bench_ref_fb 400 -- $TR --master-port 29613 bench.py --gpus "$N" --steps 3 --warmup 3 --seq "$SEQ" --impl reference
step bench_ours_fb 300 -- $TR --master-port 29614 bench.py --gpus "$N" --steps 5 --warmup 3 --seq "$SEQ"
step bench_ours_fwd 300 -- $TR --master-port 29615 bench.py --gpus "$N" --steps 5 --warmup 3 --seq "$SEQ" --mode fwd
if [ -n "$FLAGS" ]; then
  # shellcheck disable=SC2086
  step bench_ours_fb_flags 300 $FLAGS -- $TR --master-port 29616 bench.py --gpus "$N" --steps 5 --warmup 3 --seq "$SEQ"
fi
for f in "$OUT"/bench_*.log; do echo "$(basename "$f" .log): $(grep -h '^{' "$f" | tail -1 | python -c '
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print(d.get("value"), d.get("unit"), d.get("ms_per_step"), "ms  e2e", (d.get("e2e") or {}).get("value"), " comm", d.get("comm"))
except Exception as e:
    print("no json:", e)')"; done | tee "$OUT/summary.txt"
echo "failed steps: $fail"
exit $fail
