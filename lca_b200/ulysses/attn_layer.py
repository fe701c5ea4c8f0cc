"""``UlyssesAttention`` -- pure DeepSpeed-Ulysses sequence parallelism (all-to-all only, no ring).

Parity: ``yunchang/ulysses/attn_layer.py:15-125``: ``3 x a2a -> local attention -> a2a`` with an
explicit ``sequence_process_group``.  The local attention is the position-aware block op, so
causal / window / softcap / ALiBi (sliced per head shard -- the reference passes the un-sharded
slopes, ``:110``) / GQA all work, forward and backward.  On NVLink groups the fused engine is used
(one kernel: push + tcgen05 attention + O scatter).
"""
from __future__ import annotations

from typing import Any, Optional

import torch
from torch import Tensor

from ..globals import group_rank
from ..kernels import AttnType, select_flash_attn_impl
from ..parallel.all_to_all import SeqAllToAll4D
from ..hybrid.attn_layer import _resolve_backend, _slice_alibi


class UlyssesAttention(torch.nn.Module):
    def __init__(self, sequence_process_group=None, scatter_idx: int = 2, gather_idx: int = 1,
                 use_sync: bool = False, attn_type: AttnType = AttnType.FA, backend: Optional[str] = None) -> None:
        super().__init__()
        self.spg = sequence_process_group
        self.scatter_idx, self.gather_idx = scatter_idx, gather_idx
        self.use_sync = use_sync
        self.attn_type = attn_type
        self.attn_fn = select_flash_attn_impl(attn_type, stage="fwd-bwd")
        self.backend = _resolve_backend(backend)

    def forward(self, query: Tensor, key: Tensor, value: Tensor, dropout_p=0.0, softmax_scale=None, causal=False,
                window_size=(-1, -1), softcap=0.0, alibi_slopes=None, deterministic=False, return_attn_probs=False,
                *args: Any) -> Tensor:
        from ..parallel.fused import try_fused
        out = try_fused("ulysses", self.spg, self.backend, self.attn_type, query, key, value, "basic", dropout_p,
                        softmax_scale, causal, window_size, softcap, alibi_slopes, deterministic)
        if out is not None:
            return out
        q = SeqAllToAll4D.apply(self.spg, query, self.scatter_idx, self.gather_idx, self.use_sync)
        k = SeqAllToAll4D.apply(self.spg, key, self.scatter_idx, self.gather_idx, self.use_sync)
        v = SeqAllToAll4D.apply(self.spg, value, self.scatter_idx, self.gather_idx, self.use_sync)
        if softmax_scale is None:
            softmax_scale = q.shape[-1] ** -0.5
        kw = {}
        if dropout_p and dropout_p > 0:
            # the dropout key uses GLOBAL head indices, so the head shard of this rank needs its offset
            kw = dict(head_offset=group_rank(self.spg) * q.shape[2])
        ctx = self.attn_fn(q, k, v, dropout_p=dropout_p, softmax_scale=softmax_scale, causal=causal,
                           window_size=window_size, softcap=softcap,
                           alibi_slopes=_slice_alibi(alibi_slopes, self.spg), deterministic=deterministic,
                           return_attn_probs=return_attn_probs, **kw)
        if isinstance(ctx, tuple):
            ctx = ctx[0]
        return SeqAllToAll4D.apply(self.spg, ctx, self.gather_idx, self.scatter_idx, self.use_sync)
