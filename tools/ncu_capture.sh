#!/usr/bin/env bash
# One-GPU ncu captures (--set full, no clock control) of the forward, the dQ pass, the dK/dV pass and one fused launch
# with push CTAs (single-GPU harness: pushes into this GPU's own slab).  gpurun merges at most 64 MiB back, so the
# reports are exported ON THE BOX as CSV (raw metrics page + per-instruction source page, gzip'ed) and then deleted.
#   gpurun --timeout 1500 -- 'bash tools/ncu_capture.sh'      then here:  python tools/ncu_summary.py
set -u
cd "$(dirname "$0")/.."
OUT=${OUT:-gpurun_out/ncu_r2}
mkdir -p "$OUT" /tmp/ncu
export S=${S:-16384} H=${H:-8} N=1
cap() {  # cap <name> <kernel regex> <skip> <count> <cmd...>
  local name=$1 re=$2 skip=$3 cnt=$4; shift 4
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$re -s $skip -c $cnt -f -o /tmp/ncu/$name "$@" > "$OUT/$name.log" 2>&1
  ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv > "$OUT/$name.raw.csv" 2>/dev/null
  ncu -i /tmp/ncu/$name.ncu-rep --page source --csv 2>/dev/null | gzip -9 > "$OUT/$name.source.csv.gz"
  ncu -i /tmp/ncu/$name.ncu-rep --page details --csv 2>/dev/null | gzip -9 > "$OUT/$name.details.csv.gz"
  ls -la /tmp/ncu/$name.ncu-rep "$OUT/$name".* | awk '{print $5, $9}'
}
cap fwd fmha_fwd 2 1 python tools/gpu_time_passes.py
cap bwd fmha_bwd 2 2 python tools/gpu_time_passes.py          # N=1: launches 1-3 = dQ pass (2 warm-ups + timed), 4-6 = dK/dV pass -> captures #3 (dQ) and #4 (dK/dV)
S=8192 H=8 NCOMM=16 ITERS=2 cap fused_push fmha_fwd 1 1 python tools/debug_push_1gpu.py
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file "$OUT/launches.csv" python tools/gpu_time_passes.py > /dev/null 2>&1
du -sh "$OUT"
