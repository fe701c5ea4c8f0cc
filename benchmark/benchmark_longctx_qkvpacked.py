#!/usr/bin/env python
"""Packed-QKV benchmark (reference: ``benchmark/benchmark_longctx_qkvpacked.py:1-182``): a GLOBAL (B, S, 3, H, D)
tensor is sharded with EXTRACT_FUNC_DICT and fed to LongContextAttentionQKVPacked; bf16, 100 iterations."""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lca_b200 import EXTRACT_FUNC_DICT, LongContextAttentionQKVPacked, set_seq_parallel_pg  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--ring_impl_type", type=str, default="basic", choices=["basic", "zigzag", "strip", "stripe"])
p.add_argument("--nheads", type=int, default=2)
p.add_argument("--head_size", type=int, default=128)
p.add_argument("--seq_len", type=int, default=4 * 1024, help="GLOBAL sequence length")
p.add_argument("--batch_size", type=int, default=2)
p.add_argument("--fwd_only", action="store_true")
p.add_argument("--ulysses_degree", type=int, default=1)
p.add_argument("--num_iter", type=int, default=100)
p.add_argument("--backend", type=str, default=None)
args = p.parse_args()


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    B, S, H, D = args.batch_size, args.seq_len, args.nheads, args.head_size
    U = min(args.ulysses_degree, world)
    R = world // U
    set_seq_parallel_pg(U, R, rank, world)
    torch.manual_seed(0)
    qkv = torch.randn(B, S, 3, H, D, device=dev, dtype=torch.bfloat16)
    if world > 1:
        dist.broadcast(qkv, src=0)
    local = EXTRACT_FUNC_DICT[args.ring_impl_type](qkv, rank, world, rd=R, ud=U).detach().clone().requires_grad_()
    dout = torch.randn(B, S // world, H, D, device=dev, dtype=torch.bfloat16)
    attn = LongContextAttentionQKVPacked(ring_impl_type=args.ring_impl_type, backend=args.backend)

    def step():
        if args.fwd_only:
            with torch.no_grad():
                return attn(local, causal=True)
        local.grad = None
        out = attn(local, causal=True)
        out.backward(dout)
        return out

    for _ in range(3):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.num_iter):
        step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec = float(t)
    flops = 2.0 * B * H * S * S * D * (1.0 if args.fwd_only else 3.5)
    if rank == 0:
        print(f"{args.num_iter / sec:.3f} iter/s, {sec:.3f} sec, {flops * args.num_iter / sec / 1e12:.1f} TFLOPS "
              f"(qkvpacked {args.ring_impl_type} ulysses {U} ring {R} seq {S})")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
