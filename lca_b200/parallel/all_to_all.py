"""Ulysses head<->sequence all-to-all (4-D and packed 5-D), collective path.

Parity: ``yunchang/comm/all_to_all.py:15-259`` (``all_to_all_4D/5D``, ``SeqAllToAll4D/5D``).
This is the *collective* (NCCL / gloo) implementation used for multi-node groups, the CPU
backend and as the measured baseline; the NVLink-fused path (:mod:`lca_b200.parallel.fused`)
removes these calls from the hot path altogether.

Differences from the reference implementation (same semantics):
* the send buffer is group-major ``(P, B, S/P, H/P, D)`` so that for ``B == 1`` the receive side
  is a pure view -- one staging copy per direction instead of two (reference: two
  ``.contiguous()`` round trips per call, ``all_to_all.py:45-49,62-65``);
* staging copies are the in-tree 16-byte-vectorised permute kernel on CUDA;
* ``use_sync`` synchronises the *current stream's device* only when asked, never by default.

Token order contract (tested in tests/test_all_to_all.py): after ``scatter_idx=2,gather_idx=1``
the sequence axis is the group-rank-major concatenation of the shards and head ``h`` lives on
group rank ``h // (H/P)``; ``scatter_idx=1,gather_idx=2`` is the exact inverse.
"""
from __future__ import annotations

from typing import Any, Tuple

import torch
import torch.distributed as dist
from torch import Tensor

from ..globals import group_size
from ..ops import native


def _to_group_major(x: Tensor, G: int) -> Tensor:
    """(B, S, G, ...) -> (G, B, S, ...) contiguous."""
    if x.is_cuda and native.available():
        return native.ext().permute_group(x.contiguous(), G, True)
    nd = x.dim()
    return x.permute(2, 0, 1, *range(3, nd)).contiguous()


def _from_group_major(x: Tensor, G: int) -> Tensor:
    """(G, B, S, ...) -> (B, S, G, ...) contiguous."""
    if x.is_cuda and native.available():
        return native.ext().permute_group(x.contiguous(), G, False)
    nd = x.dim()
    return x.permute(1, 2, 0, *range(3, nd)).contiguous()


def _a2a(send: Tensor, group) -> Tensor:
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    return recv


def _maybe_sync(x: Tensor, use_sync: bool):
    if use_sync and x.is_cuda:
        torch.cuda.synchronize(x.device)


def all_to_all_4D(input: Tensor, scatter_idx: int = 2, gather_idx: int = 1, group=None, use_sync: bool = False) -> Tensor:
    assert input.dim() == 4, f"input must be 4D tensor, got {input.dim()} and shape {input.shape}"
    P = group_size(group)
    if scatter_idx == 2 and gather_idx == 1:
        B, Sl, H, D = input.shape
        if P == 1:
            return input
        if H % P:
            raise ValueError(f"heads ({H}) must be divisible by the Ulysses degree ({P})")
        Hl = H // P
        send = _to_group_major(input.reshape(B, Sl, P, Hl * D), P)        # (P, B, Sl, Hl*D)
        recv = _a2a(send, group)                                          # [j] = tokens of rank j, my heads
        _maybe_sync(recv, use_sync)
        if B == 1:
            return recv.view(1, P * Sl, Hl, D)
        out = _from_group_major(recv.view(P, B, 1, Sl * Hl * D), P)        # (B, 1, P, Sl*Hl*D)
        return out.view(B, P * Sl, Hl, D)
    if scatter_idx == 1 and gather_idx == 2:
        B, S, Hl, D = input.shape
        if P == 1:
            return input
        if S % P:
            raise ValueError(f"sequence ({S}) must be divisible by the Ulysses degree ({P})")
        Sl = S // P
        send = _to_group_major(input.reshape(B, 1, P, Sl * Hl * D), P)     # (P, B, 1, Sl*Hl*D)
        recv = _a2a(send, group)                                          # [j] = my tokens, heads of rank j
        _maybe_sync(recv, use_sync)
        out = _from_group_major(recv.view(P, B, Sl, Hl * D), P)            # (B, Sl, P, Hl*D)
        return out.view(B, Sl, P * Hl, D)
    raise RuntimeError("scatter_idx must be 1 or 2 and gather_idx must be 1 or 2")


class SeqAllToAll4D(torch.autograd.Function):
    @staticmethod
    def forward(ctx: Any, group, input: Tensor, scatter_idx: int, gather_idx: int, use_sync: bool = False) -> Tensor:
        ctx.group, ctx.scatter_idx, ctx.gather_idx, ctx.use_sync = group, scatter_idx, gather_idx, use_sync
        return all_to_all_4D(input, scatter_idx, gather_idx, group=group, use_sync=use_sync)

    @staticmethod
    def backward(ctx: Any, *grad_output: Tensor) -> Tuple[None, Tensor, None, None, None]:
        return (None, SeqAllToAll4D.apply(ctx.group, grad_output[0], ctx.gather_idx, ctx.scatter_idx, ctx.use_sync),
                None, None, None)


def all_to_all_5D(input: Tensor, scatter_idx: int = 3, gather_idx: int = 1, group=None, use_sync: bool = False) -> Tensor:
    """Packed QKV: ``(B, S/P, 3, H, D) <-> (B, S, 3, H/P, D)`` in ONE collective."""
    assert input.dim() == 5, f"input must be 5D tensor, got {input.dim()} and shape {input.shape}"
    P = group_size(group)
    if scatter_idx == 3 and gather_idx == 1:
        B, Sl, T, H, D = input.shape
        if P == 1:
            return input
        if H % P:
            raise ValueError(f"heads ({H}) must be divisible by the Ulysses degree ({P})")
        Hl = H // P
        send = input.reshape(B, Sl, T, P, Hl * D).permute(3, 0, 1, 2, 4).contiguous()   # (P, B, Sl, T, Hl*D)
        recv = _a2a(send.view(P, B, Sl, T * Hl * D), group)
        _maybe_sync(recv, use_sync)
        if B == 1:
            return recv.view(1, P * Sl, T, Hl, D)
        out = _from_group_major(recv.view(P, B, 1, Sl * T * Hl * D), P)
        return out.view(B, P * Sl, T, Hl, D)
    if scatter_idx == 1 and gather_idx == 3:
        B, S, T, Hl, D = input.shape
        if P == 1:
            return input
        Sl = S // P
        send = _to_group_major(input.reshape(B, 1, P, Sl * T * Hl * D), P)
        recv = _a2a(send, group)                                          # (P, B, Sl, T, Hl, D): [j] = heads of rank j
        _maybe_sync(recv, use_sync)
        out = _from_group_major(recv.view(P, B, Sl, T * Hl * D), P)        # (B, Sl, P, T*Hl*D)
        return out.view(B, Sl, P, T, Hl, D).transpose(2, 3).reshape(B, Sl, T, P * Hl, D)
    raise RuntimeError("scatter_idx must be 1 or 3 and gather_idx must be 1 or 3")


class SeqAllToAll5D(torch.autograd.Function):
    @staticmethod
    def forward(ctx: Any, group, input: Tensor, scatter_idx: int = 3, gather_idx: int = 1, use_sync: bool = False) -> Tensor:
        ctx.group, ctx.scatter_idx, ctx.gather_idx, ctx.use_sync = group, scatter_idx, gather_idx, use_sync
        return all_to_all_5D(input, scatter_idx, gather_idx, group=group, use_sync=use_sync)

    @staticmethod
    def backward(ctx: Any, *grad_output: Tensor):
        return (None, SeqAllToAll5D.apply(ctx.group, grad_output[0], ctx.gather_idx, ctx.scatter_idx, ctx.use_sync),
                None, None, None)
