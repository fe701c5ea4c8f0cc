"""Pure-python arithmetic of the Ulysses x Ring ("USP") device mesh.

Parity target: the group construction loops of ``yunchang/globals.py:39-78``.  Kept free of
``torch.distributed`` so it can be property-tested on CPU (tests/test_mesh.py).

Rank <-> coordinate maps
------------------------
``use_ulysses_low=True``  : ``u = rank % U``, ``r = (rank // U) % R``  (Ulysses groups are
contiguous ranks, ring groups are strided by ``U``) -- the a2a-heavy dimension sits on the
lowest ranks, which on a multi-node job are the NVLink-connected ones.
``use_ulysses_low=False`` : ``r = rank % R``, ``u = (rank // R) % U``.
In both cases ``dp = rank // (U*R)``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

Group = Tuple[int, ...]


@dataclass(frozen=True)
class MeshSpec:
    ulysses_degree: int
    ring_degree: int
    world_size: int
    rank: int
    use_ulysses_low: bool
    dp_degree: int
    dp_rank: int
    ulysses_rank: int
    ring_rank: int
    ulysses_group: Group          # global ranks of my Ulysses group, ordered by ulysses_rank
    ring_group: Group             # global ranks of my ring group, ordered by ring_rank
    sp_group: Group               # all U*R ranks of my replica, ordered by global rank
    dp_group: Group
    all_ulysses_groups: Tuple[Group, ...]
    all_ring_groups: Tuple[Group, ...]
    all_sp_groups: Tuple[Group, ...]
    all_dp_groups: Tuple[Group, ...]

    @property
    def sp_degree(self) -> int:
        return self.ulysses_degree * self.ring_degree

    def global_rank(self, ulysses_rank: int, ring_rank: int) -> int:
        """Global rank of mesh coordinate ``(u, r)`` inside my replica."""
        return coords_to_rank(
            ulysses_rank, ring_rank, self.dp_rank, self.ulysses_degree, self.ring_degree,
            self.use_ulysses_low,
        )

    def sp_local_rank(self, ulysses_rank: int, ring_rank: int) -> int:
        """Index of ``(u, r)`` inside ``sp_group`` (== global rank - replica offset)."""
        return self.global_rank(ulysses_rank, ring_rank) - self.dp_rank * self.sp_degree


def rank_to_coords(rank: int, U: int, R: int, use_ulysses_low: bool = True):
    """-> (ulysses_rank, ring_rank, dp_rank)."""
    sp = U * R
    dp, local = divmod(rank, sp)
    if use_ulysses_low:
        r, u = divmod(local, U)
    else:
        u, r = divmod(local, R)
    return u, r, dp


def coords_to_rank(u: int, r: int, dp: int, U: int, R: int, use_ulysses_low: bool = True) -> int:
    local = r * U + u if use_ulysses_low else u * R + r
    return dp * U * R + local


def build_mesh_spec(U: int, R: int, rank: int, world_size: int, use_ulysses_low: bool = True) -> MeshSpec:
    if U < 1 or R < 1:
        raise ValueError(f"degrees must be >= 1, got ulysses={U} ring={R}")
    sp = U * R
    if world_size % sp != 0:
        raise ValueError(
            f"world_size ({world_size}) must be divisible by ulysses_degree*ring_degree "
            f"({U}*{R}={sp})"
        )
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} out of range for world_size {world_size}")
    dp_degree = world_size // sp

    ulysses_groups, ring_groups, sp_groups = [], [], []
    for dp in range(dp_degree):
        sp_groups.append(tuple(range(dp * sp, (dp + 1) * sp)))
        for r in range(R):
            ulysses_groups.append(
                tuple(coords_to_rank(u, r, dp, U, R, use_ulysses_low) for u in range(U))
            )
        for u in range(U):
            ring_groups.append(
                tuple(coords_to_rank(u, r, dp, U, R, use_ulysses_low) for r in range(R))
            )
    dp_groups = [tuple(l + dp * sp for dp in range(dp_degree)) for l in range(sp)]

    u, r, dp = rank_to_coords(rank, U, R, use_ulysses_low)
    mine = lambda groups: next(g for g in groups if rank in g)
    return MeshSpec(
        ulysses_degree=U, ring_degree=R, world_size=world_size, rank=rank,
        use_ulysses_low=use_ulysses_low, dp_degree=dp_degree, dp_rank=dp,
        ulysses_rank=u, ring_rank=r,
        ulysses_group=mine(ulysses_groups), ring_group=mine(ring_groups),
        sp_group=mine(sp_groups), dp_group=mine(dp_groups),
        all_ulysses_groups=tuple(ulysses_groups), all_ring_groups=tuple(ring_groups),
        all_sp_groups=tuple(sp_groups), all_dp_groups=tuple(dp_groups),
    )
