#!/usr/bin/env python
"""Reference-style benchmark CLI (same flags as the reference's ``benchmark/benchmark_longctx.py:13-75``).

    torchrun --nproc_per_node 8 benchmark/benchmark_longctx.py --nheads 8 --head_size 128 --seq_len 32768 \
        --ulysses_degree 1 --ring_impl_type zigzag --fwd_only

``--seq_len`` is the LOCAL shard length, as in the reference (global S = seq_len * world).  Differences: time is
measured on the device and reduced with MAX over ranks (the reference prints rank 0's wall-to-wall events), TFLOPS
are reported next to iter/s, ``--attn_type torch`` works, and ``--backend`` picks fused|collective.
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lca_b200 import (AsyncLongContextAttention, LongContextAttention, UlyssesAttention,  # noqa: E402
                      set_seq_parallel_pg)
from lca_b200.kernels import AttnType  # noqa: E402

p = argparse.ArgumentParser(description="args for benchmark.")
p.add_argument("--ring_impl_type", type=str, default="basic", choices=["basic", "zigzag", "strip", "stripe"])
p.add_argument("--nheads", type=int, default=2)
p.add_argument("--head_size", type=int, default=128)
p.add_argument("--seq_len", type=int, default=4 * 1024)
p.add_argument("--group_num", type=int, default=1)
p.add_argument("--batch_size", type=int, default=2)
p.add_argument("--fwd_only", action="store_true")
p.add_argument("--use_ulysses_lowdim", action="store_true", default=True)
p.add_argument("--use_qkvpack", action="store_true", default=False)
p.add_argument("--ulysses_degree", type=int, default=1)
p.add_argument("--use_profiler", action="store_true", default=False)
p.add_argument("--use_ulysses", action="store_true", default=False)
p.add_argument("--use_async", action="store_true", default=False)
p.add_argument("--attn_type", type=str, default="fa", choices=["fa", "fa3", "torch"])
p.add_argument("--no_causal", action="store_true", default=False)
p.add_argument("--backend", type=str, default=None, choices=[None, "auto", "fused", "collective"])
p.add_argument("--num_iter", type=int, default=10)
args = p.parse_args()


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cuda = torch.cuda.is_available()
    if world > 1:
        dist.init_process_group("nccl" if cuda else "gloo")
    dev = torch.device("cuda", local_rank) if cuda else torch.device("cpu")
    if cuda:
        torch.cuda.set_device(dev)
    dtype = torch.float16 if cuda else torch.float32
    B, S, H, D = args.batch_size, args.seq_len, args.nheads, args.head_size
    Hkv = H // args.group_num
    assert S % (2 * world) == 0 and D % 8 == 0 and H % args.group_num == 0
    causal = not args.no_causal
    q = torch.randn(B, S, H, D, device=dev, dtype=dtype, requires_grad=True)
    k = torch.randn(B, S, Hkv, D, device=dev, dtype=dtype, requires_grad=True)
    v = torch.randn(B, S, Hkv, D, device=dev, dtype=dtype, requires_grad=True)
    dout = torch.randn(B, S, H, D, device=dev, dtype=dtype)
    U = min(args.ulysses_degree, world)
    set_seq_parallel_pg(U, world // U, rank, world, args.use_ulysses_lowdim)
    at = AttnType.from_string(args.attn_type) if cuda else AttnType.TORCH
    if args.use_ulysses:
        attn = UlyssesAttention(attn_type=at, backend=args.backend)
    elif args.use_async:
        attn = AsyncLongContextAttention(ring_impl_type=args.ring_impl_type, attn_type=at, backend=args.backend)
    else:
        attn = LongContextAttention(ring_impl_type=args.ring_impl_type, attn_type=at, use_pack_qkv=args.use_qkvpack,
                                    backend=args.backend)

    def step():
        if args.fwd_only:
            with torch.no_grad():
                return attn(q, k, v, causal=causal)
        q.grad = k.grad = v.grad = None
        out = attn(q, k, v, causal=causal)
        out.backward(dout)
        return out

    for _ in range(3):
        step()
    if world > 1:
        dist.barrier()
    if cuda:
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    import time
    t0 = time.perf_counter()
    prof = None
    if args.use_profiler:
        prof = torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CPU, torch.profiler.ProfilerActivity.CUDA],
                                      on_trace_ready=torch.profiler.tensorboard_trace_handler("./profile/"))
        prof.__enter__()
    for _ in range(args.num_iter):
        step()
    if prof is not None:
        prof.__exit__(None, None, None)
    if cuda:
        e1.record()
        torch.cuda.synchronize()
        sec = e0.elapsed_time(e1) / 1e3
    else:
        sec = time.perf_counter() - t0
    t = torch.tensor([sec], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec = float(t)
    Sg = S * world
    flops = 4.0 * B * H * Sg * Sg * D * (0.5 if causal else 1.0) * (1.0 if args.fwd_only else 3.5)
    if rank == 0:
        print(f"\033[91m {args.num_iter / sec:.3f} iter/s, {sec:.3f} sec, {flops * args.num_iter / sec / 1e12:.1f} TFLOPS "
              f"(ring_impl_type {args.ring_impl_type} ulysses {U} ring {world // U} global_seq {Sg} fwd_only {args.fwd_only})\033[00m")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
