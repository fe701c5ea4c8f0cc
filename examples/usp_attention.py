"""Minimal end-to-end use of the library: one USP (Ulysses x Ring) attention call, checked against a single-device run.

    # CPU, two processes over gloo (PyTorch engine; what the CPU test-suite runs):
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 examples/usp_attention.py --device cpu --ulysses 1
    # one node of B200s: the fused NVLink kernels are picked automatically (backend="auto")
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 examples/usp_attention.py --ulysses 2 --seq 65536

Same calls as with the reference (``yunchang``): ``set_seq_parallel_pg`` -> ``EXTRACT_FUNC_DICT[...]`` ->
``LongContextAttention(...)(q, k, v, causal=True)``.
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lca_b200 import EXTRACT_FUNC_DICT, LongContextAttention, set_seq_parallel_pg  # noqa: E402
from lca_b200.kernels import AttnType  # noqa: E402
from lca_b200.kernels.attention import flash_attn_func, pytorch_attn_func  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cuda", choices=["cuda", "cpu"])
    ap.add_argument("--ulysses", type=int, default=1)
    ap.add_argument("--ring-impl", default="zigzag", choices=["basic", "zigzag", "strip"])
    ap.add_argument("--seq", type=int, default=2048)
    ap.add_argument("--heads", type=int, default=8)
    ap.add_argument("--kv-heads", type=int, default=4)
    ap.add_argument("--head-dim", type=int, default=64)
    ap.add_argument("--window", type=int, default=-1)
    a = ap.parse_args()

    cpu = a.device == "cpu"
    dist.init_process_group("gloo" if cpu else "nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cpu") if cpu else torch.device("cuda", rank % torch.cuda.device_count())
    if not cpu:
        torch.cuda.set_device(dev)
    dtype = torch.float32 if cpu else torch.bfloat16
    U, R = a.ulysses, world // a.ulysses
    set_seq_parallel_pg(U, R, rank, world)

    # the same global tensors on every rank (seeded), sharded with the layout that matches the ring variant
    g = torch.Generator().manual_seed(0)
    q, k, v, do = (torch.randn(1, a.seq, h, a.head_dim, generator=g).to(dev, dtype)
                   for h in (a.heads, a.kv_heads, a.kv_heads, a.heads))
    shard = lambda t: EXTRACT_FUNC_DICT[a.ring_impl](t, rank, world, rd=R, ud=U).detach().clone()   # noqa: E731
    lq, lk, lv = (shard(t).requires_grad_() for t in (q, k, v))

    attn = LongContextAttention(ring_impl_type=a.ring_impl, attn_type=AttnType.TORCH if cpu else AttnType.FA)
    kw = dict(causal=True, window_size=(a.window, 0) if a.window >= 0 else (-1, -1))
    out = attn(lq, lk, lv, **kw)
    out.backward(shard(do))

    # single-device reference on the global tensors
    q1, k1, v1 = (t.clone().requires_grad_() for t in (q, k, v))
    ref = (pytorch_attn_func if cpu else flash_attn_func)(q1, k1, v1, **kw)
    ref.backward(do)
    tol = 1e-4 if cpu else 3e-2
    err_o = (out.float() - shard(ref.detach()).float()).abs().max().item()
    err_q = (lq.grad.float() - shard(q1.grad).float()).abs().max().item() / (q1.grad.float().abs().max().item() + 1e-9)
    err_k = (lk.grad.float() - shard(k1.grad).float()).abs().max().item() / (k1.grad.float().abs().max().item() + 1e-9)
    ok = err_o < tol and err_q < tol and err_k < tol
    t = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(t)
    if rank == 0:
        print(f"mesh ulysses={U} x ring={R} ({a.ring_impl}), S={a.seq}: max |out - ref| = {err_o:.2e}, "
              f"rel dq err = {err_q:.2e}, rel dk err = {err_k:.2e} -> {'OK' if int(t.item()) == 0 else 'MISMATCH'}")
    dist.destroy_process_group()
    sys.exit(int(t.item() != 0))


if __name__ == "__main__":
    main()
