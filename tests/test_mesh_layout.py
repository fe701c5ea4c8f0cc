import pytest
import torch

from lca_b200.parallel.layout import (EXTRACT_FUNC_DICT, Seg, gather_global, local_token_index, pos_tensor,
                                      ring_positions, slice_pos, varlen_positions)
from lca_b200.parallel.mesh import build_mesh_spec, coords_to_rank, rank_to_coords


@pytest.mark.parametrize("U,R,world", [(2, 4, 8), (4, 2, 8), (8, 1, 8), (1, 8, 8), (2, 2, 8), (1, 1, 1), (2, 1, 4)])
@pytest.mark.parametrize("low", [True, False])
def test_mesh_groups_partition_world(U, R, world, low):
    seen_u, seen_r = set(), set()
    for rank in range(world):
        m = build_mesh_spec(U, R, rank, world, low)
        assert rank in m.ulysses_group and rank in m.ring_group and rank in m.sp_group
        assert len(m.ulysses_group) == U and len(m.ring_group) == R
        assert m.ulysses_group[m.ulysses_rank] == rank and m.ring_group[m.ring_rank] == rank
        assert coords_to_rank(*rank_to_coords(rank, U, R, low), U, R, low) == rank
        assert set(m.ulysses_group) & set(m.ring_group) == {rank}
        seen_u.add(m.ulysses_group)
        seen_r.add(m.ring_group)
    assert sorted(x for g in seen_u for x in g) == list(range(world))
    assert sorted(x for g in seen_r for x in g) == list(range(world))


def test_mesh_matches_reference_example():
    # world=8,U=2,R=4 -> Ulysses {0,1},{2,3},.. ; ring {0,2,4,6},{1,3,5,7}
    m = build_mesh_spec(2, 4, 5, 8, True)
    assert m.ulysses_group == (4, 5) and m.ring_group == (1, 3, 5, 7)
    assert set(build_mesh_spec(2, 4, 0, 8, True).all_ulysses_groups) == {(0, 1), (2, 3), (4, 5), (6, 7)}


def test_mesh_rejects_bad_degrees():
    with pytest.raises(ValueError):
        build_mesh_spec(3, 2, 0, 8)


@pytest.mark.parametrize("variant", ["basic", "zigzag", "stripe"])
@pytest.mark.parametrize("U,R", [(1, 4), (2, 2), (4, 1), (2, 4)])
def test_layout_is_a_partition_and_invertible(variant, U, R):
    S = 16 * U * R
    allidx = torch.cat([local_token_index(variant, S, u, r, U, R) for r in range(R) for u in range(U)])
    assert sorted(allidx.tolist()) == list(range(S))
    x = torch.arange(2 * S * 3, dtype=torch.float32).view(2, S, 3)
    key = {"basic": "basic", "zigzag": "zigzag", "stripe": "strip"}[variant]
    shards = [EXTRACT_FUNC_DICT[key](x, rank, U * R, rd=R, ud=U) for rank in range(U * R)]
    assert all(s.shape == (2, S // (U * R), 3) for s in shards)
    assert torch.equal(gather_global(variant, shards, R, U), x)


def test_zigzag_matches_reference_definition():
    # cat(c_r, c_{2R-1-r}).chunk(U)[u]  (extract_local.py:37-46)
    S, R, U = 64, 4, 2
    x = torch.arange(S).view(1, S)
    chunks = x.chunk(2 * R, dim=1)
    for r in range(R):
        for u in range(U):
            ref = torch.cat([chunks[r], chunks[2 * R - 1 - r]], dim=1).chunk(U, dim=1)[u]
            got = EXTRACT_FUNC_DICT["zigzag"](x, r * U + u, U * R, rd=R, ud=U)
            assert torch.equal(got, ref)


def test_stripe_matches_reference_definition():
    S, R, U = 48, 4, 1
    x = torch.arange(S).view(1, S, 1)
    for r in range(R):
        got = EXTRACT_FUNC_DICT["strip"](x, r, R, rd=R, ud=U)
        assert got.flatten().tolist() == list(range(r, S, R))


def test_ring_positions_and_slices():
    assert pos_tensor(ring_positions("zigzag", 1, 4, 8)).tolist() == [4, 5, 6, 7, 24, 25, 26, 27]
    assert pos_tensor(ring_positions("stripe", 2, 4, 3)).tolist() == [2, 6, 10]
    assert pos_tensor(ring_positions("basic", 3, 4, 2)).tolist() == [6, 7]
    sp = slice_pos(ring_positions("zigzag", 1, 4, 8), 2, 6)
    assert pos_tensor(sp).tolist() == [6, 7, 24, 25]


def test_varlen_positions_groups():
    spec = varlen_positions("zigzag", 0, 2, [0, 4, 10])
    assert [s.group for s in spec] == [0, 0, 1, 1]
    assert pos_tensor(spec).tolist() == [0, 1, 6, 7, 0, 1, 2, 9, 10, 11]
