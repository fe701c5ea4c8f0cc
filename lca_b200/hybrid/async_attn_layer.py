"""``AsyncLongContextAttention`` -- head-group pipelined Ulysses + ring.

Parity: ``yunchang/hybrid/async_attn_layer.py:13-202``: the ``H/U`` local heads are processed as
``H/U`` groups; the all-to-all of group ``i+1`` overlaps the attention of group ``i`` and the
output all-to-all of group ``i`` overlaps the attention of group ``i+1``.

Differences: (1) the comm stream is correctly ordered after the producer of q/k/v (the reference
launches NCCL on a side stream with no dependency on the current stream -- a latent race,
SURVEY 2.8-11); (2) backward works (autograd through the per-group ops; the reference defines an
``nn.Module.backward`` autograd never calls); (3) GQA is supported when ``Hkv % U == 0``;
(4) on the fused NVLink backend this module is the same persistent kernel as
``LongContextAttention`` -- tile-granular overlap subsumes head-group pipelining -- so it simply
delegates when that backend is active.
"""
from __future__ import annotations

from typing import Any, Optional

import torch
from torch import Tensor

from ..globals import PROCESS_GROUP, group_size
from ..kernels import AttnType
from ..parallel.all_to_all import SeqAllToAll4D
from ..parallel.layout import canonical_variant
from .attn_layer import _dropout_kw, _resolve_backend, _slice_alibi
from .utils import RING_IMPL_DICT


class AsyncLongContextAttention(torch.nn.Module):
    def __init__(self, scatter_idx: int = 2, gather_idx: int = 1, ring_impl_type: str = "basic",
                 attn_type: AttnType = AttnType.FA, backend: Optional[str] = None) -> None:
        super().__init__()
        if not PROCESS_GROUP.initialized:
            raise AssertionError("use set_seq_parallel_pg() first")
        self.ring_pg = PROCESS_GROUP.RING_PG
        self.ulysses_pg = PROCESS_GROUP.ULYSSES_PG
        self.scatter_idx, self.gather_idx = scatter_idx, gather_idx
        self.ring_impl_type = ring_impl_type
        self.variant = canonical_variant(ring_impl_type)
        self.ring_attn_fn = RING_IMPL_DICT[ring_impl_type]
        self.attn_type = attn_type
        self.backend = _resolve_backend(backend)
        self._comm_stream = None

    def forward(self, query: Tensor, key: Tensor, value: Tensor, dropout_p=0.0, softmax_scale=None, causal=False,
                window_size=(-1, -1), softcap=0.0, alibi_slopes=None, deterministic=False, return_attn_probs=False,
                *args: Any) -> Tensor:
        from ..parallel.fused import try_fused
        out = try_fused("mesh", PROCESS_GROUP, self.backend, self.attn_type, query, key, value, self.variant, dropout_p,
                        softmax_scale, causal, window_size, softcap, alibi_slopes, deterministic)
        if out is not None:
            return out

        U = group_size(self.ulysses_pg)
        B, Sl, H, D = query.shape
        Hkv = key.shape[2]
        if H % U or Hkv % U:
            raise ValueError(f"heads ({H}, kv {Hkv}) must be divisible by the Ulysses degree {U}")
        n_groups = H // U                      # one local head per pipeline stage, as in the reference
        g = H // Hkv
        if softmax_scale is None:
            softmax_scale = D ** -0.5
        alibi = _slice_alibi(alibi_slopes, self.ulysses_pg)

        use_streams = query.is_cuda
        if use_streams and self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=query.device)
        cur = torch.cuda.current_stream(query.device) if use_streams else None

        # head h_global = u*(H/U) + i  -> stage i takes column i of every Ulysses block
        qv = query.view(B, Sl, U, n_groups, D)
        kv_groups = Hkv // U
        kvw = key.view(B, Sl, U, kv_groups, D)
        vvw = value.view(B, Sl, U, kv_groups, D)

        def shuffle_in(i):
            qi = qv[:, :, :, i].contiguous()                               # (B, Sl, U, D): one head per rank
            ki = kvw[:, :, :, i // g].contiguous()
            vi = vvw[:, :, :, i // g].contiguous()
            return tuple(SeqAllToAll4D.apply(self.ulysses_pg, t, self.scatter_idx, self.gather_idx, False)
                         for t in (qi, ki, vi))                           # each (B, S/R, 1, D)

        staged = [None] * n_groups
        events = [None] * n_groups
        if use_streams:
            self._comm_stream.wait_stream(cur)                             # inputs are ready on `cur`
            with torch.cuda.stream(self._comm_stream):
                for i in range(n_groups):
                    staged[i] = shuffle_in(i)
                    events[i] = torch.cuda.Event()
                    events[i].record(self._comm_stream)
        else:
            for i in range(n_groups):
                staged[i] = shuffle_in(i)

        outs = [None] * n_groups
        out_events = [None] * n_groups
        dkw = _dropout_kw(dropout_p, self.ulysses_pg, n_groups)     # one seed per call; stage i = global head u*H/U + i
        for i in range(n_groups):
            if use_streams:
                cur.wait_event(events[i])
            qi, ki, vi = staged[i]
            for t in (qi, ki, vi):
                if use_streams:
                    t.record_stream(cur)
            oi = self.ring_attn_fn(qi, ki, vi, dropout_p=dropout_p, softmax_scale=softmax_scale, causal=causal,
                                   window_size=window_size, softcap=softcap,
                                   alibi_slopes=None if alibi is None else alibi[..., i:i + 1].contiguous(),
                                   deterministic=deterministic, return_attn_probs=False, group=self.ring_pg,
                                   attn_type=self.attn_type, backend="collective",
                                   **(dict(dkw, head_offset=dkw["head_offset"] + i) if dkw else {}))
            if use_streams:
                done = torch.cuda.Event()
                done.record(cur)
                with torch.cuda.stream(self._comm_stream):
                    self._comm_stream.wait_event(done)
                    oi.record_stream(self._comm_stream)
                    outs[i] = SeqAllToAll4D.apply(self.ulysses_pg, oi, self.gather_idx, self.scatter_idx, False)
                    out_events[i] = torch.cuda.Event()
                    out_events[i].record(self._comm_stream)
            else:
                outs[i] = SeqAllToAll4D.apply(self.ulysses_pg, oi, self.gather_idx, self.scatter_idx, False)
        if use_streams:
            for e in out_events:
                cur.wait_event(e)
            for o in outs:
                o.record_stream(cur)
        # outs[i]: (B, Sl, U, D) = head i of every Ulysses block -> (B, Sl, U, n_groups, D) -> (B, Sl, H, D)
        return torch.stack(outs, dim=3).reshape(B, Sl, H, D)
