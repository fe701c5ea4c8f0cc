#!/usr/bin/env bash
# One-GPU ncu captures of the three hot kernels (run under gpurun).  Reports land in gpurun_out/.
set -x
export S=${S:-16384} H=${H:-8} N=1
ncu --set full --clock-control none --import-source on -k regex:fmha_fwd -s 2 -c 1 -o gpurun_out/prof_fwd python tools/gpu_time_passes.py > gpurun_out/ncu_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fmha_bwd -s 4 -c 2 -o gpurun_out/prof_bwd2 python tools/gpu_time_passes.py > gpurun_out/ncu_bwd2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python tools/gpu_time_passes.py > /dev/null 2>&1
ls -la gpurun_out | tail -8
