#!/usr/bin/env bash
# The four GPU configs of BASELINE.json on one 8xB200 node, both arms (reference first), fwd and fwd+bwd.
set -euo pipefail
cd "$(dirname "$0")/.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555"
run() { for impl in reference ours; do $TR bench.py --gpus 8 --impl $impl "$@" | tail -1; done; }
run --mode fwd    --ulysses 8 --seq 32768  --heads 32 --head-dim 128 --ring-impl basic            # config 2: pure all-to-all
run --mode fwd    --ulysses 1 --seq 262144 --heads 8  --head-dim 128 --ring-impl zigzag           # config 3: pure ring (headline)
run --mode fwd    --ulysses 2 --seq 131072 --heads 32 --kv-heads 4 --head-dim 128 --ring-impl zigzag --window 4096   # config 4
run --mode fwdbwd --ulysses 4 --seq 65536  --heads 16 --head-dim 128 --ring-impl zigzag --qkvpacked   # config 5 (bf16; fp8 experimental)
