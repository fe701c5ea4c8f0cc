"""Ring attention on packed variable-length batches.

Parity: ``yunchang/ring/ring_flash_attn_varlen.py`` and ``zigzag_ring_flash_attn_varlen.py``
(``*_varlen_func(q, k, v, cu_seqlens, max_seqlen, ...)`` with q/k/v ``(total_local, H, D)``).
Every sequence must be split evenly over the ring (and over ``2R`` chunks for zigzag), exactly as
in the reference; ``cu_seqlens`` holds the cumulative *local* lengths.

The reference needs an LSE flatten/unflatten round trip per ring step (``:72-75,84``) and a
host-synchronising half-index computation (``zigzag...varlen.py:27-59``).  Here a packed shard is
just a list of (position, group) segments, so the very same ring loops and kernels as the dense
case run unchanged, the LSE is produced directly in ``(H, total)`` layout and nothing syncs.
"""
from __future__ import annotations

import torch

from ..ops.attention import AttnParams
from .layout import canonical_variant
from .ring_attention import _engine_for, ring_attn_backward, ring_attn_forward


class RingVarlenAttnFunc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, cu_seqlens, max_seqlen, variant, dropout_p, softmax_scale, causal, window_size,
                softcap, alibi_slopes, deterministic, return_softmax, group, attn_type):
        p = AttnParams.make(q.unsqueeze(0), softmax_scale, causal, window_size, softcap, alibi_slopes, dropout_p,
                            deterministic)
        if p.dropout_p > 0:
            raise NotImplementedError("dropout is not supported on the varlen ring path")
        engine = _engine_for(attn_type)
        cu = [int(x) for x in cu_seqlens.tolist()] if torch.is_tensor(cu_seqlens) else [int(x) for x in cu_seqlens]
        out, lse = ring_attn_forward(group, q.unsqueeze(0), k.unsqueeze(0), v.unsqueeze(0), variant, p, engine, 0, cu, cu)
        out, lse = out.squeeze(0), lse.squeeze(0)           # (total,H,D), (H,total)
        ctx.save_for_backward(q, k, v, out, lse)
        ctx.p, ctx.variant, ctx.group, ctx.engine, ctx.cu = p, variant, group, engine, cu
        if return_softmax:
            ctx.mark_non_differentiable(lse)
            return out, lse, None
        return out

    @staticmethod
    def backward(ctx, dout, *args):
        q, k, v, out, lse = ctx.saved_tensors
        dq, dk, dv = ring_attn_backward(ctx.group, dout.unsqueeze(0), q.unsqueeze(0), k.unsqueeze(0), v.unsqueeze(0),
                                        out.unsqueeze(0), lse.unsqueeze(0), ctx.variant, ctx.p, ctx.engine, 0,
                                        ctx.cu, ctx.cu)
        return (dq.squeeze(0), dk.squeeze(0), dv.squeeze(0)) + (None,) * 13


def _make(variant):
    variant = canonical_variant(variant)

    def func(q, k, v, cu_seqlens, max_seqlen, dropout_p=0.0, softmax_scale=None, causal=False, window_size=(-1, -1),
             softcap=0.0, alibi_slopes=None, deterministic=False, return_attn_probs=False, group=None, attn_type=None,
             backend=None):
        """q/k/v ``(total_local, H, D)``; ``backend`` as in the dense ring functions: on an NVLink group the packed
        shard runs through the fused engine (the kernels take one attention group per sequence), no NCCL P2P, no
        per-step launches, no LSE merge."""
        from .fused import resolve_backend, try_fused
        res = try_fused("ring", group, resolve_backend(backend), attn_type, q.unsqueeze(0), k.unsqueeze(0), v.unsqueeze(0),
                        variant, dropout_p, softmax_scale, causal, window_size, softcap, alibi_slopes, deterministic,
                        cu_seqlens, return_attn_probs)
        if res is not None:
            if return_attn_probs:
                return res[0].squeeze(0), res[1].squeeze(0), None
            return res.squeeze(0)
        return RingVarlenAttnFunc.apply(q, k, v, cu_seqlens, max_seqlen, variant, dropout_p, softmax_scale, causal,
                                        window_size, softcap, alibi_slopes, deterministic, return_attn_probs, group,
                                        attn_type)

    def kvpacked(q, kv, cu_seqlens, max_seqlen, *a, **kw):
        return func(q, kv[:, 0], kv[:, 1], cu_seqlens, max_seqlen, *a, **kw)

    def qkvpacked(qkv, cu_seqlens, max_seqlen, *a, **kw):
        return func(qkv[:, 0], qkv[:, 1], qkv[:, 2], cu_seqlens, max_seqlen, *a, **kw)

    return func, kvpacked, qkvpacked


(ring_flash_attn_varlen_func, ring_flash_attn_varlen_kvpacked_func,
 ring_flash_attn_varlen_qkvpacked_func) = _make("basic")
(zigzag_ring_flash_attn_varlen_func, zigzag_ring_flash_attn_varlen_kvpacked_func,
 zigzag_ring_flash_attn_varlen_qkvpacked_func) = _make("zigzag")
