"""Sequence layouts: which global tokens a rank owns under each ring variant.

Parity target: ``yunchang/comm/extract_local.py:7-60``.  On top of the reference's three
``*_extract_local`` functions this module provides what the fused kernels and the exact
(global-position) masks need and the reference lacks:

* :func:`local_token_index` -- the explicit global token indices of a rank's shard
  (so RoPE position ids / labels can be permuted the same way, and so layouts are testable);
* :func:`gather_global` -- the inverse (shards -> global tensor);
* :class:`Seg` / :func:`ring_positions` -- compact ``(start, count, stride)`` descriptions of
  the global positions of a ring rank's tokens *after* the Ulysses gather; the CUDA kernels
  take these instead of materialised position tensors.

Layout algebra (S global tokens, R ring ranks, U Ulysses ranks, P = U*R):
  basic : ring rank r owns ``[r*S/R, (r+1)*S/R)``; Ulysses rank u the u-th of U equal pieces.
  zigzag: 2R chunks; ring rank r owns ``cat(c_r, c_{2R-1-r})``; cut into U pieces.
  stripe: token t lives on ring rank ``t % R`` at local index ``t // R``; cut into U pieces.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple

import torch

from .mesh import rank_to_coords


@dataclass(frozen=True)
class Seg:
    """``count`` tokens at global positions ``start + i*stride``."""

    start: int
    count: int
    stride: int = 1
    group: int = 0      # tokens only attend within their group (varlen: one group per sequence)

    @property
    def last(self) -> int:
        return self.start + (self.count - 1) * self.stride


PosSpec = Tuple[Seg, ...]

_VARIANT_ALIASES = {
    "basic": "basic", "zigzag": "zigzag", "strip": "stripe", "stripe": "stripe",
    "basic_pytorch": "basic", "basic_flashinfer": "basic", "basic_npu": "basic",
}


def canonical_variant(name: str) -> str:
    try:
        return _VARIANT_ALIASES[name]
    except KeyError:
        raise ValueError(f"unknown ring_impl_type {name!r}; choose from {sorted(_VARIANT_ALIASES)}")


def ring_positions(variant: str, ring_rank: int, ring_degree: int, local_len: int) -> PosSpec:
    """Global positions of the ``local_len`` tokens ring rank ``ring_rank`` holds (post-Ulysses)."""
    variant = canonical_variant(variant)
    R, r, L = ring_degree, ring_rank, local_len
    if variant == "basic" or R == 1 and variant != "stripe":
        # zigzag with R == 1 is cat(c_0, c_1) == identity
        return (Seg(r * L, L, 1),)
    if variant == "zigzag":
        if L % 2:
            raise ValueError(f"zigzag needs an even local length, got {L}")
        h = L // 2
        return (Seg(r * h, h, 1), Seg((2 * R - 1 - r) * h, h, 1))
    if variant == "stripe":
        return (Seg(r, L, R),)
    raise AssertionError(variant)


def pos_tensor(spec: PosSpec, device=None) -> torch.Tensor:
    parts = [
        s.start + s.stride * torch.arange(s.count, dtype=torch.int64, device=device) for s in spec
    ]
    return torch.cat(parts) if len(parts) > 1 else parts[0]


def group_tensor(spec: PosSpec, device=None) -> torch.Tensor:
    return torch.cat([torch.full((s.count,), s.group, dtype=torch.int64, device=device) for s in spec])


def has_groups(spec: PosSpec) -> bool:
    return any(s.group != 0 for s in spec)


def varlen_positions(variant: str, ring_rank: int, ring_degree: int, cu_seqlens) -> PosSpec:
    """Positions of a packed (varlen) local shard: sequence ``i`` contributes
    ``cu_seqlens[i+1]-cu_seqlens[i]`` local tokens laid out like a dense shard of that length and
    forms attention group ``i``.  (Reference: per-sequence halves in
    ``zigzag_ring_flash_attn_varlen.py:27-42``.)"""
    cu = [int(x) for x in cu_seqlens]
    out = []
    for i in range(len(cu) - 1):
        L = cu[i + 1] - cu[i]
        if L == 0:
            continue
        for s in ring_positions(variant, ring_rank, ring_degree, L):
            out.append(Seg(s.start, s.count, s.stride, i))
    return tuple(out)


def pos_min_max(spec: PosSpec) -> Tuple[int, int]:
    lo = min(min(s.start, s.last) for s in spec if s.count > 0)
    hi = max(max(s.start, s.last) for s in spec if s.count > 0)
    return lo, hi


def pos_len(spec: PosSpec) -> int:
    return sum(s.count for s in spec)


def slice_pos(spec: PosSpec, begin: int, end: int) -> PosSpec:
    """Positions of local rows ``[begin, end)``."""
    out: List[Seg] = []
    off = 0
    for s in spec:
        lo, hi = max(begin, off), min(end, off + s.count)
        if hi > lo:
            out.append(Seg(s.start + (lo - off) * s.stride, hi - lo, s.stride, s.group))
        off += s.count
    return tuple(out)


# ------------------------------------------------------------------------------------------
# explicit token index maps
# ------------------------------------------------------------------------------------------
def local_token_index(
    variant: str, seqlen: int, ulysses_rank: int, ring_rank: int, ulysses_degree: int, ring_degree: int
) -> torch.Tensor:
    """int64 tensor of the ``seqlen/(U*R)`` global token ids owned by mesh coordinate (u, r)."""
    U, R = ulysses_degree, ring_degree
    P = U * R
    variant = canonical_variant(variant)
    need = 2 * P if variant == "zigzag" and R > 1 else P
    if seqlen % P or (variant == "zigzag" and R > 1 and seqlen % (2 * R)):
        raise ValueError(f"seqlen {seqlen} not divisible as required by {variant} (need % {need} == 0)")
    ring_tokens = pos_tensor(ring_positions(variant, ring_rank, R, seqlen // R))
    return ring_tokens.chunk(U)[ulysses_rank].clone()


def _coords(rank: int, rd: int, ud: int, use_ulysses_low: bool = True):
    u, r, _ = rank_to_coords(rank, ud, rd, use_ulysses_low)
    return u, r


def _extract(variant: str, value: torch.Tensor, rank: int, world_size: int, rd: int, ud: int, dim: int = 1,
             use_ulysses_low: bool = True) -> torch.Tensor:
    if rd * ud > world_size or world_size % (rd * ud):
        raise ValueError(f"ring degree {rd} x ulysses degree {ud} incompatible with world_size {world_size}")
    u, r = _coords(rank, rd, ud, use_ulysses_low)
    idx = local_token_index(variant, value.shape[dim], u, r, ud, rd).to(value.device)
    return value.index_select(dim, idx).contiguous()


def basic_extract_local(value, rank, world_size, rd=None, ud=None, *args, **kwargs):
    """Contiguous chunk ``rank`` of ``world_size`` (``extract_local.py:25-26``)."""
    if rd is None or ud is None:
        return value.chunk(world_size, dim=1)[rank % world_size].detach().clone()
    return _extract("basic", value, rank, world_size, rd, ud, kwargs.get("dim", 1),
                    kwargs.get("use_ulysses_low", True)).detach()


def zigzag_extract_local(value, rank, world_size, rd, ud, dim=1, *args, **kwargs):
    """``cat(c_r, c_{2R-1-r})`` split over the Ulysses group (``extract_local.py:29-49``)."""
    return _extract("zigzag", value, rank, world_size, rd, ud, dim, kwargs.get("use_ulysses_low", True))


def stripe_extract_local(value, rank, world_size, rd, ud, *args, **kwargs):
    """Round-robin tokens over ring ranks (``extract_local.py:7-22``).

    Unlike the reference (which indexes with the *global* rank, so it silently assumes
    ``dp_degree == 1``), mesh coordinates are used, so DP replicas shard identically.
    """
    return _extract("stripe", value, rank, world_size, rd, ud, kwargs.get("dim", 1),
                    kwargs.get("use_ulysses_low", True))


EXTRACT_FUNC_DICT = {
    "basic": basic_extract_local,
    "strip": stripe_extract_local,
    "stripe": stripe_extract_local,   # the reference only has the (misspelt) "strip" key
    "zigzag": zigzag_extract_local,
    "basic_pytorch": basic_extract_local,
    "basic_flashinfer": basic_extract_local,
    "basic_npu": basic_extract_local,
}


def gather_global(variant: str, shards: Sequence[torch.Tensor], rd: int, ud: int, dim: int = 1,
                  use_ulysses_low: bool = True) -> torch.Tensor:
    """Inverse of the extract functions: ``shards[rank]`` for the ``rd*ud`` ranks -> global tensor."""
    P = rd * ud
    if len(shards) != P:
        raise ValueError(f"need {P} shards, got {len(shards)}")
    seqlen = shards[0].shape[dim] * P
    shape = list(shards[0].shape)
    shape[dim] = seqlen
    out = shards[0].new_empty(shape)
    for rank, sh in enumerate(shards):
        u, r = _coords(rank, rd, ud, use_ulysses_low)
        idx = local_token_index(variant, seqlen, u, r, ud, rd).to(sh.device)
        out.index_copy_(dim, idx, sh)
    return out
