#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/dbg1
echo "== bulk";   timeout 120 python tools/debug_push_1gpu.py 2>&1 | tail -6
echo "== bulk B=2 strided (GQA slice), D=64"; B=2 H=8 HKV=2 D=64 S=4096 NCOMM=3 timeout 120 python tools/debug_push_1gpu.py 2>&1 | tail -4
echo "== bulk wide rows"; H=32 HKV=32 S=1024 NCOMM=8 timeout 120 python tools/debug_push_1gpu.py 2>&1 | tail -4
echo "== bulk under memcheck"
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/debug_push_1gpu.py > gpurun_out/dbg1/memcheck.log 2>&1; tail -n 6 gpurun_out/dbg1/memcheck.log | cut -c1-220
