#!/usr/bin/env bash
# GQA sweep (reference: scripts/run_gqa.sh): H=64, group_num=8, fwd and fwd+bwd, all ulysses degrees.
set -euo pipefail
GPUS=${GPUS:-8}; cd "$(dirname "$0")/.."
for U in 8 4 2 1; do
  for RING in basic zigzag strip; do
    for MODE in "--fwd_only" ""; do
      torchrun --standalone --local-addr 127.0.0.1 --nproc_per_node "$GPUS" benchmark/benchmark_longctx.py \
        --nheads 64 --group_num 8 --head_size 128 --seq_len 4096 --batch_size 2 --ulysses_degree "$U" \
        --ring_impl_type "$RING" $MODE
    done
  done
done
