#!/usr/bin/env bash
# Non-causal (DiT-style) attention sweep (reference: scripts/run_dit.sh).
set -euo pipefail
GPUS=${GPUS:-8}; cd "$(dirname "$0")/.."
for U in 8 4 2 1; do
  torchrun --standalone --local-addr 127.0.0.1 --nproc_per_node "$GPUS" benchmark/benchmark_longctx.py \
    --nheads 24 --head_size 128 --seq_len 16384 --batch_size 1 --ulysses_degree "$U" --ring_impl_type basic \
    --fwd_only --no_causal
done
