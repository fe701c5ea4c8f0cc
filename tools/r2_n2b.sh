#!/usr/bin/env bash
# second 2-GPU pass: dynamic scheduler (push CTAs rejoin the compute pool) after the shared-memory race fix, push-CTA count sweep
set -u
cd "$(dirname "$0")/.."
N=${1:-2}
OUT=gpurun_out/r2_n${N}b
mkdir -p "$OUT"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
SEQ=${SEQ:-131072}
port=29800
bench() {
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
c = d.get("comm") or {}
print("    ", d.get("value"), d.get("unit"), d.get("ms_per_step"), "ms  e2e", (d.get("e2e") or {}).get("value"), " compute_only", c.get("compute_only_ms"), "exposed", c.get("exposed_comm_ms"), " check_ok", (d.get("check") or {}).get("ok"))'
  else
    echo "    FAILED (exit $?)"; tail -n 12 "$OUT/$name.log" | cut -c1-300 | sed 's/^/    /'
  fi
}
echo "=== tests_dyn"
LCA_B200_DYN_SCHED=1 timeout 600 python -m pytest tests/test_fused_multigpu.py -x -q -rA -k "matrix and $N" > "$OUT/tests_dyn.log" 2>&1; grep -h "PASS\|FAIL\|passed\|failed" "$OUT/tests_dyn.log" | tail -n 25 | cut -c1-160
A="--steps 5 --warmup 3 --seq $SEQ --no-check"
for nc in 8 12 16 24; do
  bench fb_dyn$nc LCA_B200_DYN_SCHED=1 LCA_B200_COMM_CTAS=$nc -- $A
done
bench fb_static12 LCA_B200_COMM_CTAS=12 -- $A
for nc in 8 16; do
  bench fwd_dyn$nc LCA_B200_DYN_SCHED=1 LCA_B200_COMM_CTAS=$nc -- $A --mode fwd
done
U="--steps 10 --warmup 3 --seq 32768 --heads 32 --ulysses $N --ring-impl basic --no-check"
for nc in 8 16; do
  bench uly_fwd_dyn$nc LCA_B200_DYN_SCHED=1 LCA_B200_COMM_CTAS=$nc -- $U --mode fwd
  bench uly_fb_dyn$nc LCA_B200_DYN_SCHED=1 LCA_B200_COMM_CTAS=$nc -- $U
done
bench uly_fb_static8 -- $U
