#!/usr/bin/env bash
# Lean eight-GPU pass (charged 8x): fused test matrix n=8, every BASELINE GPU config (ours) from ONE launch.
#   gpurun --gpus 8 --timeout 600 -- 'bash tools/r2_n8c.sh'
set -u
cd "$(dirname "$0")/.."
N=8
OUT=gpurun_out/r2_n8c
mkdir -p "$OUT"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
summ() { grep -h '^{' "$1" | python -c '
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    c = d.get("comm") or {}; k = d.get("check") or {}; s = d.get("staging") or {}
    print("    cfg", d.get("config_id"), d.get("impl"), (d.get("config") or {}).get("mode"), (d.get("config") or {}).get("parallelism"), "|", d.get("value"), "TFLOPS", d.get("ms_per_step"), "ms | e2e", (d.get("e2e") or {}).get("value"), "| compute_only", c.get("compute_only_ms"), "exposed", c.get("exposed_comm_ms"), "| check", k.get("ok"), k.get("max_err_out"), k.get("max_rel_err_dq"), k.get("max_rel_err_dk"), "| slab", s.get("slab_bytes_per_rank"), "| clk", (d.get("clocks") or {}).get("sm_mhz"))'; }
echo "=== tests n=8"
timeout 300 python -m pytest tests/test_fused_multigpu.py -q -rA -p no:cacheprovider -k "matrix and 8" > "$OUT/tests.log" 2>&1; grep -h "PASS\|FAIL\|passed\|failed\|Error" "$OUT/tests.log" | tail -n 16 | cut -c1-200
echo "=== ours: configs 3,2,4,5 x fwdbwd,fwd"
timeout 300 $TR --master-port 29921 bench.py --gpus $N --steps 10 --warmup 3 --configs 3,2,4,5 --modes fwdbwd,fwd > "$OUT/ours_all.log" 2>&1 || tail -n 20 "$OUT/ours_all.log" | cut -c1-300
summ "$OUT/ours_all.log"
