#!/usr/bin/env bash
# Round-2 two/four-GPU re-validation after the kernel-role slimming: fused test matrix (named cases), collective backend,
# headline bench (config 3) and the Ulysses shape (config 2) at this GPU count.
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/r2_n2c.sh 2'
set -u
cd "$(dirname "$0")/.."
N=${1:-2}
OUT=gpurun_out/r2_n${N}c
mkdir -p "$OUT"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
summ() { grep -h '^{' "$1" | python -c '
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    c = d.get("comm") or {}; k = d.get("check") or {}; s = d.get("staging") or {}
    print("    cfg", d.get("config_id"), d.get("impl"), (d.get("config") or {}).get("mode"), (d.get("config") or {}).get("parallelism"), "|", d.get("value"), "TFLOPS", d.get("ms_per_step"), "ms | e2e", (d.get("e2e") or {}).get("value"), "| compute_only", c.get("compute_only_ms"), "exposed", c.get("exposed_comm_ms"), "| check", k.get("ok"), k.get("max_err_out"), k.get("max_rel_err_dq"), k.get("max_rel_err_dk"), "| slab", s.get("slab_bytes_per_rank"), "| clk", (d.get("clocks") or {}).get("sm_mhz"))'; }
echo "=== tests n=$N"
timeout 600 python -m pytest tests/test_fused_multigpu.py -q -rA -p no:cacheprovider -k "(matrix and $N) or collective" > "$OUT/tests.log" 2>&1; grep -h "PASS\|FAIL\|passed\|failed\|Error" "$OUT/tests.log" | tail -n 30 | cut -c1-200
echo "=== ours: configs ${CFGS:-3,2} x fwdbwd,fwd"
timeout 420 $TR --master-port 29911 bench.py --gpus $N --steps ${STEPS:-5} --warmup 3 --configs ${CFGS:-3,2} --modes fwdbwd,fwd > "$OUT/ours.log" 2>&1 || tail -n 20 "$OUT/ours.log" | cut -c1-300
summ "$OUT/ours.log"
