#!/usr/bin/env bash
# Round-2 eight-GPU pass (charged 8x: keep it short): fused test matrix, every BASELINE GPU config in both arms from ONE
# launch per arm, the headline config with the old push engine for the A/B, NVLink byte counters around one run.
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/r2_n8.sh 8'
set -u
cd "$(dirname "$0")/.."
N=${1:-8}
OUT=gpurun_out/r2_n$N
mkdir -p "$OUT"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
summ() { grep -h '^{' "$1" | python -c '
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    c = d.get("comm") or {}; k = d.get("check") or {}; s = d.get("staging") or {}
    print("    cfg", d.get("config_id"), d.get("impl"), (d.get("config") or {}).get("mode"), (d.get("config") or {}).get("parallelism"), "|", d.get("value"), "TFLOPS", d.get("ms_per_step"), "ms | e2e", (d.get("e2e") or {}).get("value"), "| compute_only", c.get("compute_only_ms"), "exposed", c.get("exposed_comm_ms"), "| check", k.get("ok"), k.get("max_err_out"), k.get("max_rel_err_dq"), k.get("max_rel_err_dk"), "| slab", s.get("slab_bytes_per_rank"))'; }
echo "=== tests n=$N"
timeout 420 python -m pytest tests/test_fused_multigpu.py -x -q -rA -k "matrix and $N" > "$OUT/tests.log" 2>&1; grep -h "PASS\|FAIL\|passed\|failed\|Error" "$OUT/tests.log" | tail -n 16 | cut -c1-200
nvidia-smi nvlink -gt d -i 0 > "$OUT/nvlink_before.txt" 2>&1
echo "=== ours: configs 3,2,4,5 x fwdbwd,fwd"
timeout 420 $TR --master-port 29901 bench.py --gpus $N --steps 10 --warmup 3 --configs 3,2,4,5 --modes fwdbwd,fwd > "$OUT/ours_all.log" 2>&1 || tail -n 20 "$OUT/ours_all.log" | cut -c1-300
summ "$OUT/ours_all.log"
nvidia-smi nvlink -gt d -i 0 > "$OUT/nvlink_after.txt" 2>&1
echo "=== reference: configs 3,2,4,5 x fwdbwd,fwd"
timeout 600 $TR --master-port 29902 bench.py --gpus $N --steps 5 --warmup 3 --configs 3,2,4,5 --modes fwdbwd,fwd --impl reference > "$OUT/ref_all.log" 2>&1 || tail -n 20 "$OUT/ref_all.log" | cut -c1-300
summ "$OUT/ref_all.log"
echo "=== ours config 3 A/B: scalar push + static schedule + 8 push CTAs (the round-1 engine)"
LCA_B200_PUSH=scalar LCA_B200_DYN_SCHED=0 LCA_B200_COMM_CTAS=8 timeout 300 $TR --master-port 29903 bench.py --gpus $N --steps 5 --warmup 3 --no-check > "$OUT/ours_c3_r1engine.log" 2>&1 || tail -n 20 "$OUT/ours_c3_r1engine.log" | cut -c1-300
summ "$OUT/ours_c3_r1engine.log"
if [ "${EXTRA:-1}" = "1" ]; then
echo "=== ours config 3: 24 push CTAs / bulk+static 8"
LCA_B200_COMM_CTAS=24 timeout 300 $TR --master-port 29904 bench.py --gpus $N --steps 5 --warmup 3 --no-check --no-comm-probe > "$OUT/ours_c3_nc24.log" 2>&1; summ "$OUT/ours_c3_nc24.log"
LCA_B200_DYN_SCHED=0 LCA_B200_COMM_CTAS=8 timeout 300 $TR --master-port 29905 bench.py --gpus $N --steps 5 --warmup 3 --no-check --no-comm-probe > "$OUT/ours_c3_static8.log" 2>&1; summ "$OUT/ours_c3_static8.log"
fi
