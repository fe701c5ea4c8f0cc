#!/usr/bin/env bash
# Sweep ulysses degree x ring variant, packed vs unpacked (role of the reference's scripts/run_qkvpack_compare.sh).
set -euo pipefail
GPUS=${GPUS:-8}; NHEADS=${NHEADS:-8}; HEAD=${HEAD:-128}; SEQ=${SEQ:-131072}; BS=${BS:-2}
cd "$(dirname "$0")/.."
for U in 8 4 2 1; do
  [ "$U" -gt "$GPUS" ] && continue
  for RING in basic zigzag strip; do
    torchrun --standalone --local-addr 127.0.0.1 --nproc_per_node "$GPUS" benchmark/benchmark_longctx_qkvpacked.py \
      --nheads "$NHEADS" --head_size "$HEAD" --seq_len "$SEQ" --batch_size "$BS" --ulysses_degree "$U" \
      --ring_impl_type "$RING" --fwd_only
    torchrun --standalone --local-addr 127.0.0.1 --nproc_per_node "$GPUS" benchmark/benchmark_longctx.py \
      --nheads "$NHEADS" --head_size "$HEAD" --seq_len $((SEQ / GPUS)) --batch_size "$BS" --ulysses_degree "$U" \
      --ring_impl_type "$RING" --fwd_only
  done
done
