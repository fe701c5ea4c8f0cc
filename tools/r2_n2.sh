#!/usr/bin/env bash
# Round-2 two-GPU validation of the bulk push engine: test matrix under the new default, under the dynamic scheduler
# (push CTAs rejoin the compute pool), and with the legacy scalar push; then the bench A/B grid.
#   gpurun --gpus 2 --timeout 1500 -- 'bash tools/r2_n2.sh'
set -u
cd "$(dirname "$0")/.."
N=${1:-2}
OUT=gpurun_out/r2_n$N
mkdir -p "$OUT"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
SEQ=${SEQ:-131072}
fail=0
port=29700
step() {   # step <name> <timeout_s> <env...> -- <cmd...>
  local name=$1 tmo=$2; shift 2
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  echo "=== $name (${envs[*]:-default})"
  if env ${envs[@]+"${envs[@]}"} timeout "$tmo" "$@" > "$OUT/$name.log" 2>&1; then
    echo "    ok"; grep -h "PASS\|FAIL\|passed\|failed" "$OUT/$name.log" | tail -n 40 | cut -c1-200 | sed 's/^/    /'
  else
    echo "    FAILED (exit $?)"; grep -h "PASS\|FAIL" "$OUT/$name.log" | cut -c1-200 | sed 's/^/    /'; tail -n 25 "$OUT/$name.log" | cut -c1-300 | sed 's/^/    /'; fail=$((fail + 1))
  fi
}
bench() {  # bench <name> <env...> -- <bench args...>
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  port=$((port + 1))
  echo "=== $name (${envs[*]:-default}) $*"
  if env ${envs[@]+"${envs[@]}"} timeout 300 $TR --master-port $port bench.py --gpus "$N" "$@" > "$OUT/$name.log" 2>&1; then
    grep -h '^{' "$OUT/$name.log" | tail -1 | python -c '
import sys, json
d = json.loads(sys.stdin.read())
print("    ", d.get("value"), d.get("unit"), d.get("ms_per_step"), "ms  e2e", (d.get("e2e") or {}).get("value"), " comm", d.get("comm"), " check", d.get("check"))'
  else
    echo "    FAILED (exit $?)"; tail -n 15 "$OUT/$name.log" | cut -c1-300 | sed 's/^/    /'; fail=$((fail + 1))
  fi
}

step tests_bulk 600 -- python -m pytest tests/test_fused_multigpu.py -x -q -rA -k "matrix and $N or collective"
step tests_bulk_dyn 600 LCA_B200_DYN_SCHED=1 -- python -m pytest tests/test_fused_multigpu.py -x -q -rA -k "matrix and $N"
step tests_scalar 300 LCA_B200_PUSH=scalar LCA_B200_TEST_CASES=u2_basic,r2_zigzag_gqa,u2r2_zigzag,r8_zigzag,u8_basic -- python -m pytest tests/test_fused_multigpu.py -x -q -rA -k "matrix and $N"

A="--steps 5 --warmup 3 --seq $SEQ"
bench fb_scalar8 LCA_B200_PUSH=scalar -- $A
bench fb_bulk8 -- $A
bench fb_bulk4 LCA_B200_COMM_CTAS=4 -- $A
bench fb_bulk2 LCA_B200_COMM_CTAS=2 -- $A
bench fb_bulk8_dyn LCA_B200_DYN_SCHED=1 -- $A
bench fb_bulk4_dyn LCA_B200_DYN_SCHED=1 LCA_B200_COMM_CTAS=4 -- $A
bench fwd_bulk4 LCA_B200_COMM_CTAS=4 -- $A --mode fwd
bench fwd_bulk4_dyn LCA_B200_DYN_SCHED=1 LCA_B200_COMM_CTAS=4 -- $A --mode fwd
# BASELINE config 2 shape (pure Ulysses, S=32K, h=32) at this GPU count
U="--steps 10 --warmup 3 --seq 32768 --heads 32 --ulysses $N --ring-impl basic"
bench uly_fwd_scalar LCA_B200_PUSH=scalar -- $U --mode fwd
bench uly_fwd_bulk4 LCA_B200_COMM_CTAS=4 -- $U --mode fwd
bench uly_fb_bulk4 LCA_B200_COMM_CTAS=4 -- $U
bench uly_fb_bulk4_dyn LCA_B200_DYN_SCHED=1 LCA_B200_COMM_CTAS=4 -- $U
bench uly_fb_ref -- $U --impl reference
bench fb_ref -- --steps 3 --warmup 3 --seq $SEQ --impl reference
echo "failed steps: $fail"
exit $fail
