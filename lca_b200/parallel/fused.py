"""Entry point of the fused NVLink backend shared by ``LongContextAttention``, ``LongContextAttentionQKVPacked`` and
``UlyssesAttention``: decides PER CALL whether the fused engine can take the inputs (the engine object itself is cached
per process group in :mod:`fused_engine`; nothing about one call's dtype / head_dim is remembered for the next),
reserves the symmetric slab (falling back consistently across ranks when it does not fit), agrees on the dropout seed,
and wraps the launch in an NVTX range."""
from __future__ import annotations

import os
from typing import Optional

import torch

from ..utils.logging import get_logger
from ..utils.profiling import nvtx_range

_LOG = get_logger()
_REPORTED = set()


def _note_once(key, msg: str, level: str = "info") -> None:
    """Fallback decisions are logged once per distinct reason (LCA_B200_LOGLEVEL=INFO shows them)."""
    if key not in _REPORTED:
        _REPORTED.add(key)
        getattr(_LOG, level)(msg)


def native_dropout_enabled() -> bool:
    """In-kernel dropout (coordinate-keyed masks, validated on hardware in round 2) is the default;
    ``LCA_B200_NATIVE_DROPOUT=0`` restores the PyTorch engine for ``dropout_p > 0``."""
    return os.environ.get("LCA_B200_NATIVE_DROPOUT", "1") == "1"


def get_engine_if_supported(process_group_state, q, strict: bool = False):
    try:
        from .fused_engine import engine_for_mesh
    except ImportError:
        if strict:
            raise
        return None
    return engine_for_mesh(process_group_state, q, strict)


def get_ulysses_engine_if_supported(group, q, strict: bool = False):
    try:
        from .fused_engine import engine_for_ulysses_group
    except ImportError:
        if strict:
            raise
        return None
    return engine_for_ulysses_group(group, q, strict)


def get_ring_engine_if_supported(group, q, strict: bool = False):
    try:
        from .fused_engine import engine_for_ring_group
    except ImportError:
        if strict:
            raise
        return None
    return engine_for_ring_group(group, q, strict)


def resolve_backend(requested: Optional[str]) -> str:
    b = requested or os.environ.get("LCA_B200_BACKEND", "auto")
    if b not in ("auto", "fused", "collective"):
        raise ValueError(f"backend must be auto|fused|collective, got {b!r}")
    return b


def try_fused(kind: str, pg, backend: str, attn_type, q, k, v, variant: str, dropout_p, softmax_scale, causal,
              window_size, softcap, alibi_slopes, deterministic, cu_seqlens=None, return_lse: bool = False):
    """Run one attention call on the fused engine.  ``kind``: "mesh" (``pg`` = PROCESS_GROUP state), "ulysses"
    (``pg`` = the sequence process group) or "ring" (``pg`` = a ring group; ``cu_seqlens`` = cumulative LOCAL lengths
    of a packed variable-length shard).  Returns ``None`` when the caller has to take the collective path, else the
    output (``return_lse``: ``(out, lse (B, H, rows))``)."""
    tag = getattr(attn_type, "value", "")
    if backend == "collective" or tag.startswith("torch"):
        return None
    if tag in ("sage_fp8", "sage_fp8_sm90", "sage_auto", "sparse_sage"):
        # quantised / user-supplied forward kernels run per ring block on the collective path (as in the reference,
        # where SAGE_* only exist as "fwd-only" block kernels inside the ring loop): the push CTAs are part of the
        # 16-bit kernels
        _note_once(("attn_type", tag), f"AttnType {tag}: collective path (per-block forward kernel)")
        return None
    if not q.is_cuda:
        return None
    strict = backend == "fused"
    dropout_p = float(dropout_p or 0.0)
    if dropout_p > 0.0:
        from ..ops import dropout as _d
        if not native_dropout_enabled() or float(softcap or 0.0) != 0.0:
            _note_once(("dropout", kind), "fused backend skipped: dropout with softcap / LCA_B200_NATIVE_DROPOUT=0")
            return None
        if _d.p8_of(dropout_p) == 0:
            dropout_p = 0.0
    if kind == "mesh":
        eng = get_engine_if_supported(pg, q, strict)
    elif kind == "ulysses":
        eng = get_ulysses_engine_if_supported(pg, q, strict)
    else:
        eng = get_ring_engine_if_supported(pg, q, strict)
    if eng is None:
        return None
    if cu_seqlens is not None:
        cu_seqlens = [int(x) for x in (cu_seqlens.tolist() if torch.is_tensor(cu_seqlens) else cu_seqlens)]
        if eng.U != 1 or eng.too_many_segments(variant, q.shape[1], cu_seqlens):
            if strict:
                raise RuntimeError("fused backend: packed batch has more sequence segments than one launch takes")
            _note_once(("segs", len(cu_seqlens)), "fused backend skipped: too many packed sequences for one launch")
            return None
    if not eng.supports_shapes(q, k):
        if strict:
            raise RuntimeError(f"fused backend cannot take q {tuple(q.shape)} / k {tuple(k.shape)} on a "
                               f"{eng.U}x{eng.R} mesh (rows % 8, head divisibility)")
        _note_once(("shape", tuple(q.shape), tuple(k.shape)), f"fused backend skipped for q {tuple(q.shape)} k {tuple(k.shape)}")
        return None
    need_bwd = torch.is_grad_enabled() and (q.requires_grad or k.requires_grad or v.requires_grad)
    kv_chunk = eng.reserve(q, k, need_bwd)
    if not kv_chunk:
        if strict:
            raise RuntimeError("fused backend: the symmetric slab does not fit (LCA_B200_SLAB_MAX_GB / free memory)")
        _note_once(("slab", tuple(q.shape)), f"fused backend skipped: staging slab for q {tuple(q.shape)} does not fit; "
                   "using the collective path", "warning")
        return None
    seed = eng.dropout_seed() if dropout_p > 0.0 else 0
    with nvtx_range(f"lca.fused.{kind}.{variant}"):
        return eng.attention(q, k, v, variant, softmax_scale, causal, window_size, softcap, alibi_slopes, deterministic,
                             dropout_p, seed, cu_seqlens, return_lse, kv_chunk)
