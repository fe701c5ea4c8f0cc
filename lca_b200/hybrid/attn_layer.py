"""``LongContextAttention`` / ``LongContextAttentionQKVPacked`` -- the USP (Ulysses x Ring) modules.

Parity: ``yunchang/hybrid/attn_layer.py:14-259`` (constructor kwargs, forward kwargs, shapes).
Input/output shards are ``(B, S/P, H, D)`` (packed: ``(B, S/P, 3, H, D)`` -> ``(B, S/P, H, D)``).

Two execution backends, selected per module (``backend=`` kwarg or ``LCA_B200_BACKEND`` env):

``"fused"``  (default when the SP group is NVLink-reachable on one node and the inputs are CUDA bf16/fp16)
    ONE persistent sm_100a kernel per rank does the Ulysses head shuffle, the ring K/V exchange and
    the attention math: comm CTAs push Q/K/V head-slices into peers' symmetric staging buffers over
    NVLink (P2P stores + release flags), compute CTAs run tcgen05 attention on segments as their
    flags arrive and scatter O tiles straight into the owners' output buffers.  No NCCL call, no
    layout copy, no LSE merge kernel.  (:mod:`lca_b200.parallel.fused`)
``"collective"``
    ``SeqAllToAll4D`` + position-aware ring loop over NCCL/gloo -- multi-node, CPU, and the
    like-for-like structure of the reference (3 a2a -> ring fn -> a2a, ``attn_layer.py:111-158``).

Fixes relative to the reference: ``use_pack_qkv=True`` works (reference: ``.continous()`` typo,
``:88``); ALiBi slopes are sliced per Ulysses head shard; sliding windows are exact across ring
blocks; GQA only needs ``Hkv % U == 0`` on the collective path and nothing on the fused path.
"""
from __future__ import annotations

import os
from typing import Any, Optional

import torch
from torch import Tensor

from ..globals import PROCESS_GROUP, group_rank, group_size
from ..kernels import AttnType
from ..parallel.all_to_all import SeqAllToAll4D, SeqAllToAll5D
from ..parallel.layout import canonical_variant
from .utils import RING_IMPL_DICT, RING_IMPL_QKVPACKED_DICT


def _slice_alibi(alibi_slopes, group):
    """Slopes are indexed by global head; after the Ulysses shuffle a rank holds heads
    ``[u*H/U, (u+1)*H/U)``."""
    if alibi_slopes is None:
        return None
    U, u = group_size(group), group_rank(group)
    if U == 1:
        return alibi_slopes
    H = alibi_slopes.shape[-1]
    hl = H // U
    return alibi_slopes[..., u * hl:(u + 1) * hl].contiguous()


def _resolve_backend(requested: Optional[str]) -> str:
    b = requested or os.environ.get("LCA_B200_BACKEND", "auto")
    if b not in ("auto", "fused", "collective"):
        raise ValueError(f"backend must be auto|fused|collective, got {b!r}")
    return b


def _dropout_kw(dropout_p, ulysses_pg, local_heads: int, stage: int = 0) -> dict:
    """Extra ring-function arguments when dropout is on: one seed per module call (drawn from torch's CPU generator,
    so ranks that share ``torch.manual_seed`` share it) and the global index of this rank's first local head --
    the dropout mask is a function of global coordinates (``ops/dropout.py``), hence identical to a single-device
    run with the same seed whatever the Ulysses x Ring layout."""
    if not dropout_p or dropout_p <= 0:
        return {}
    seed = int(torch.randint(1, 2**31 - 1, (1,)).item())
    return dict(dropout_seed=seed, head_offset=group_rank(ulysses_pg) * local_heads + stage)


class LongContextAttention(torch.nn.Module):
    """Arguments (same as the reference, plus ``backend``):
        scatter_idx, gather_idx : all-to-all axes (2, 1)
        ring_impl_type          : "basic" | "zigzag" | "strip"/"stripe" | "basic_pytorch" | "basic_flashinfer"
        use_pack_qkv            : move Q,K,V through ONE all-to-all (collective backend)
        use_sync                : device-synchronise after each all-to-all (debug)
        attn_type               : AttnType (FA -> native tcgen05 kernels; TORCH* -> PyTorch engine)
        attn_processor          : accepted for API compatibility (sparse-sage hook), unused
    """

    def __init__(self, scatter_idx: int = 2, gather_idx: int = 1, ring_impl_type: str = "basic",
                 use_pack_qkv: bool = False, use_sync: bool = False, attn_type: AttnType = AttnType.FA,
                 attn_processor: torch.nn.Module = None, backend: Optional[str] = None) -> None:
        super().__init__()
        if not PROCESS_GROUP.initialized:
            raise AssertionError("use set_seq_parallel_pg() first (ulysses/ring process groups are not set)")
        self.ring_pg = PROCESS_GROUP.RING_PG
        self.ulysses_pg = PROCESS_GROUP.ULYSSES_PG
        self.use_pack_qkv = use_pack_qkv
        self.use_sync = use_sync
        self.attn_type = attn_type
        self.scatter_idx = scatter_idx
        self.gather_idx = gather_idx
        self.attn_processor = attn_processor
        self.ring_impl_type = ring_impl_type
        self.variant = canonical_variant(ring_impl_type)
        self.ring_attn_fn = RING_IMPL_DICT[ring_impl_type]
        self.backend = _resolve_backend(backend)

    def forward(self, query: Tensor, key: Tensor, value: Tensor, dropout_p=0.0, softmax_scale=None, causal=False,
                window_size=(-1, -1), softcap=0.0, alibi_slopes=None, deterministic=False, return_attn_probs=False,
                *args: Any) -> Tensor:
        from ..parallel.fused import try_fused
        out = try_fused("mesh", PROCESS_GROUP, self.backend, self.attn_type, query, key, value, self.variant, dropout_p,
                        softmax_scale, causal, window_size, softcap, alibi_slopes, deterministic)
        if out is not None:
            return out

        alibi = _slice_alibi(alibi_slopes, self.ulysses_pg)
        if self.use_pack_qkv and key.shape == query.shape:
            qkv = torch.cat([query, key, value]).contiguous()           # (3B, S/P, H, D)
            qkv = SeqAllToAll4D.apply(self.ulysses_pg, qkv, self.scatter_idx, self.gather_idx, self.use_sync)
            query_layer, key_layer, value_layer = torch.chunk(qkv, 3, dim=0)
        else:
            query_layer = SeqAllToAll4D.apply(self.ulysses_pg, query, self.scatter_idx, self.gather_idx, self.use_sync)
            key_layer = SeqAllToAll4D.apply(self.ulysses_pg, key, self.scatter_idx, self.gather_idx, self.use_sync)
            value_layer = SeqAllToAll4D.apply(self.ulysses_pg, value, self.scatter_idx, self.gather_idx, self.use_sync)
        out = self.ring_attn_fn(query_layer, key_layer, value_layer, dropout_p=dropout_p, softmax_scale=softmax_scale,
                                causal=causal, window_size=window_size, softcap=softcap, alibi_slopes=alibi,
                                deterministic=deterministic, return_attn_probs=return_attn_probs, group=self.ring_pg,
                                attn_type=self.attn_type, attn_processor=self.attn_processor, backend="collective",
                                **_dropout_kw(dropout_p, self.ulysses_pg, query_layer.shape[2]))
        context_layer = out[0] if isinstance(out, tuple) else out
        # (B, S/R, H/U, D) -> (B, S/P, H, D)
        return SeqAllToAll4D.apply(self.ulysses_pg, context_layer, self.gather_idx, self.scatter_idx, self.use_sync)


class LongContextAttentionQKVPacked(torch.nn.Module):
    """Packed-QKV USP attention: ``qkv (B, S/P, 3, H, D) -> (B, S/P, H, D)``
    (``hybrid/attn_layer.py:164-259``; MHA only, like the reference)."""

    def __init__(self, scatter_idx: int = 3, gather_idx: int = 1, ring_impl_type: str = "basic",
                 use_sync: bool = False, attn_type: AttnType = AttnType.FA, backend: Optional[str] = None) -> None:
        super().__init__()
        if not PROCESS_GROUP.initialized:
            raise AssertionError("use set_seq_parallel_pg() first (ulysses/ring process groups are not set)")
        self.ring_pg = PROCESS_GROUP.RING_PG
        self.ulysses_pg = PROCESS_GROUP.ULYSSES_PG
        self.scatter_idx = scatter_idx
        self.gather_idx = gather_idx
        self.use_sync = use_sync
        self.attn_type = attn_type
        self.ring_impl_type = ring_impl_type
        self.variant = canonical_variant(ring_impl_type)
        self.ring_attn_fn = RING_IMPL_QKVPACKED_DICT[ring_impl_type]
        self.backend = _resolve_backend(backend)

    def forward(self, qkv: Tensor, dropout_p=0.0, softmax_scale=None, causal=False, window_size=(-1, -1),
                softcap=0.0, alibi_slopes=None, deterministic=False, return_attn_probs=False, *args: Any) -> Tensor:
        from ..parallel.fused import try_fused
        # strided views of the packed tensor feed the push kernel directly: no unpack copy
        out = try_fused("mesh", PROCESS_GROUP, self.backend, self.attn_type, qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2],
                        self.variant, dropout_p, softmax_scale, causal, window_size, softcap, alibi_slopes, deterministic)
        if out is not None:
            return out
        U = group_size(self.ulysses_pg)
        if U > 1:
            qkv = SeqAllToAll5D.apply(self.ulysses_pg, qkv, self.scatter_idx, self.gather_idx, self.use_sync)
        out = self.ring_attn_fn(qkv, dropout_p=dropout_p, softmax_scale=softmax_scale, causal=causal,
                                window_size=window_size, softcap=softcap,
                                alibi_slopes=_slice_alibi(alibi_slopes, self.ulysses_pg), deterministic=deterministic,
                                return_attn_probs=return_attn_probs, group=self.ring_pg, attn_type=self.attn_type,
                                backend="collective", **_dropout_kw(dropout_p, self.ulysses_pg, qkv.shape[3]))
        out = out[0] if isinstance(out, tuple) else out
        if U > 1:
            out = SeqAllToAll4D.apply(self.ulysses_pg, out, self.gather_idx, self.scatter_idx - 1, self.use_sync)
        return out
