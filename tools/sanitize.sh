#!/usr/bin/env bash
# Race / memory checking of the single-GPU kernels with compute-sanitizer (run under gpurun, 1 GPU).
# The reference has no sanitizer story at all.
set -x
cd "$(dirname "$0")/.."
export S=${S:-1024} H=${H:-2} N=1
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 7 python tools/gpu_time_passes.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool exit=$?" | tee -a gpurun_out/sanitize_summary.txt
  tail -3 gpurun_out/sanitize_$tool.log
done
