"""Property-based tests (hypothesis) of the pure-Python algebra everything else rests on: layouts, position segments,
the block-visibility test that lets ring steps be skipped, and the dropout key."""
import pytest
import torch
from hypothesis import given, settings, strategies as st

from lca_b200.ops.attention import AttnParams, block_is_visible
from lca_b200.ops.ref_attention import _bias_and_mask
from lca_b200.parallel.layout import (EXTRACT_FUNC_DICT, Seg, gather_global, local_token_index, pos_tensor,
                                      ring_positions, slice_pos, varlen_positions)

VARIANTS = ["basic", "zigzag", "stripe"]
mesh = st.sampled_from([(1, 1), (1, 2), (2, 1), (2, 2), (1, 4), (4, 1), (2, 4), (4, 2), (1, 8), (8, 1)])


@settings(max_examples=60, deadline=None)
@given(mesh, st.sampled_from(VARIANTS), st.integers(1, 6), st.integers(1, 3))
def test_extract_then_gather_is_identity(UR, variant, chunk, H):
    U, R = UR
    P = U * R
    S = 2 * P * chunk                                   # zigzag needs 2P | S
    x = torch.arange(S * H, dtype=torch.float32).view(1, S, H, 1)
    key = "strip" if variant == "stripe" else variant
    shards = [EXTRACT_FUNC_DICT[key](x, rank, P, rd=R, ud=U) for rank in range(P)]
    assert torch.equal(gather_global(variant, shards, R, U), x)
    idx = torch.cat([local_token_index(variant, S, rank % U, rank // U, U, R) for rank in range(P)])
    assert sorted(idx.tolist()) == list(range(S))       # a partition of the tokens
    for rank in range(P):                               # the shard really is x[index map]
        assert torch.equal(shards[rank], x[:, local_token_index(variant, S, rank % U, rank // U, U, R)])


@settings(max_examples=60, deadline=None)
@given(st.sampled_from(VARIANTS), st.integers(1, 8), st.integers(1, 5), st.data())
def test_ring_positions_and_slicing(variant, R, chunk, data):
    L = 2 * chunk                                       # local tokens per ring rank
    allpos = []
    for r in range(R):
        spec = ring_positions(variant, r, R, L)
        pos = pos_tensor(spec, "cpu")
        assert pos.numel() == L
        allpos += pos.tolist()
        a = data.draw(st.integers(0, L))
        b = data.draw(st.integers(a, L))
        assert torch.equal(pos_tensor(slice_pos(spec, a, b), "cpu"), pos[a:b]) if b > a else True
    assert sorted(allpos) == list(range(R * L))


@settings(max_examples=40, deadline=None)
@given(st.sampled_from(["basic", "zigzag"]), st.integers(1, 4), st.lists(st.integers(1, 4), min_size=1, max_size=4))
def test_varlen_positions_cover_every_sequence_once(variant, R, chunks):
    lens = [2 * R * c for c in chunks]                  # global lengths, evenly divisible as the reference requires
    local = [l // R for l in lens]
    cu = [0]
    for l in local:
        cu.append(cu[-1] + l)
    seen = {i: [] for i in range(len(lens))}
    for r in range(R):
        for s in varlen_positions(variant, r, R, cu):
            seen[s.group] += [s.start + i * s.stride for i in range(s.count)]
    for i, l in enumerate(lens):
        assert sorted(seen[i]) == list(range(l))


segs = st.lists(st.tuples(st.integers(0, 60), st.integers(1, 12), st.sampled_from([1, 2, 4]), st.integers(0, 1)),
                min_size=1, max_size=3).map(lambda xs: tuple(Seg(a, n, s, g) for a, n, s, g in xs))


@settings(max_examples=200, deadline=None)
@given(segs, segs, st.booleans(), st.integers(-1, 20), st.integers(-1, 20))
def test_block_visibility_never_skips_a_visible_pair(q_pos, k_pos, causal, wl, wr):
    """Soundness of the ring-step skip: ``block_is_visible == False`` must imply that the exact mask hides every pair
    (the converse may be conservative)."""
    p = AttnParams(1.0, causal, (wl, wr))
    from lca_b200.parallel.layout import group_tensor
    mask, _ = _bias_and_mask(pos_tensor(q_pos, "cpu"), pos_tensor(k_pos, "cpu"), causal, (wl, wr), None, 1, "cpu",
                             group_tensor(q_pos, "cpu"), group_tensor(k_pos, "cpu"))
    any_visible = True if mask is None else bool((~mask).any())
    if not block_is_visible(q_pos, k_pos, p):
        assert not any_visible


@settings(max_examples=50, deadline=None)
@given(st.integers(0, 2**31 - 1), st.integers(0, 3), st.integers(0, 7), st.integers(0, 2**20), st.integers(0, 2**20))
def test_dropout_key_is_a_function_of_global_coordinates(seed, b, h, q0, k0):
    from lca_b200.ops import dropout as d
    qp, kp = torch.arange(q0, q0 + 5), torch.arange(k0, k0 + 9)
    full = d.keep_mask(seed, b + 1, h + 1, qp, kp, 0.37)
    one = d.keep_mask(seed, 1, 1, qp[2:4], kp[3:8], 0.37, head_offset=h, batch_offset=b)
    assert torch.equal(one[0, 0], full[b, h, 2:4, 3:8])


@settings(max_examples=100, deadline=None)
@given(st.lists(st.integers(1, 3), min_size=1, max_size=160))
def test_chunk_by_group_keeps_groups_whole_and_within_the_segment_budget(segs_per_group):
    """Many-sequence varlen batches are launched in chunks of whole attention groups with <= MAX_SEG segments per side
    (``native._chunk_by_group``): every segment appears exactly once, a group is never split across launches."""
    from lca_b200.ops.native import MAX_SEG, _chunk_by_group
    qrows, krows, row = [], [], 0
    for g, n in enumerate(segs_per_group):
        for _ in range(n):
            qrows.append((row, 4, 0, g))
            krows.append((row, 4, 0, g))
            row += 4
    seen_q, seen_k, group_launch = [], [], {}
    for li, (qc, kc) in enumerate(_chunk_by_group(qrows, krows)):
        assert 0 < len(qc) <= MAX_SEG and 0 < len(kc) <= MAX_SEG
        assert {r[3] for r in qc} == {r[3] for r in kc}
        for r in qc:
            assert group_launch.setdefault(r[3], li) == li
        seen_q += qc
        seen_k += kc
    assert sorted(seen_q) == sorted(qrows) and sorted(seen_k) == sorted(krows)


def _tile_iter_per_tile(q_pos0, q_rows, q_stride, q_group, ksegs, k_stride, wl, wr, BN):
    """The window tests tile by tile (the kernels' round-1 iterator): the definition the range form must reproduce."""
    qmin, qmax = q_pos0, q_pos0 + (q_rows - 1) * q_stride
    visited = []
    for si, (row0, nrows, pos0, group) in enumerate(ksegs):
        nt = (nrows + BN - 1) // BN if group == q_group else 0
        for kt in range(nt):
            r0 = kt * BN
            nv = min(BN, nrows - r0)
            ka = pos0 + r0 * k_stride
            kb = ka + (nv - 1) * k_stride
            if wr >= 0 and ka - qmax > wr:
                break                                   # later tiles of this segment are further right
            if wl >= 0 and qmin - kb > wl:
                continue                                # entirely left of the window
            visited.append((si, kt))
    return visited


def _tile_iter_twin(q_pos0, q_rows, q_stride, q_group, ksegs, k_stride, wl, wr, BN):
    """Python transcription of the kernels' ``TileIter::next`` (``csrc/fmha_fwd_sm100.cu``; the backward's iterator is
    the same with X/Y renamed): K/V tiles a block of query rows visits.  Positions grow with the tile index inside a
    segment, so the visible tiles are one contiguous range ``[lo, hi)`` per segment, computed on segment entry (C++
    integer division: every dividend below is non-negative)."""
    qmin, qmax = q_pos0, q_pos0 + (q_rows - 1) * q_stride
    visited = []
    for si, (row0, nrows, pos0, group) in enumerate(ksegs):
        nt = (nrows + BN - 1) // BN if group == q_group else 0
        lo, hi = 0, nt
        if wr >= 0 and nt > 0:
            lim = qmax + wr - pos0
            hi = 0 if lim < 0 else min(nt, lim // (BN * k_stride) + 1)
        if wl >= 0 and hi > 0:
            need = qmin - wl - pos0
            if need > 0:
                e_min = (need + k_stride - 1) // k_stride + 1
                lo = hi if e_min > nrows else (e_min + BN - 1) // BN - 1
        visited += [(si, kt) for kt in range(lo, hi)]
    return visited


@settings(max_examples=500, deadline=None)
@given(st.integers(0, 300), st.integers(1, 24), st.sampled_from([1, 2, 4]), st.integers(0, 1),
       st.lists(st.tuples(st.integers(1, 40), st.integers(0, 300), st.integers(0, 1)), min_size=1, max_size=4),
       st.sampled_from([1, 2, 4]), st.integers(-1, 64), st.integers(-1, 64), st.sampled_from([4, 8]))
def test_tile_range_form_equals_the_per_tile_window_tests(q_pos0, q_rows, q_stride, q_group, kdefs, k_stride, wl, wr, BN):
    ksegs, row = [], 0
    for nrows, pos0, group in kdefs:
        ksegs.append((row, nrows, pos0, group))
        row += nrows
    a = (q_pos0, q_rows, q_stride, q_group, ksegs, k_stride, wl, wr, BN)
    assert _tile_iter_twin(*a) == _tile_iter_per_tile(*a)


@settings(max_examples=300, deadline=None)
@given(st.integers(0, 300), st.integers(1, 24), st.sampled_from([1, 2, 4]), st.integers(0, 1),
       st.lists(st.tuples(st.integers(1, 40), st.integers(0, 300), st.integers(0, 1)), min_size=1, max_size=4),
       st.sampled_from([1, 2, 4]), st.integers(-1, 64), st.integers(-1, 64), st.booleans())
def test_tile_skipping_never_drops_a_visible_tile(q_pos0, q_rows, q_stride, q_group, kdefs, k_stride, wl, wr, causal):
    BN = 8
    if causal:
        wr = 0                                          # native.window_bounds: causal => right bound 0
    ksegs, row = [], 0
    for nrows, pos0, group in kdefs:
        ksegs.append((row, nrows, pos0, group))
        row += nrows
    visited = set(_tile_iter_twin(q_pos0, q_rows, q_stride, q_group, ksegs, k_stride, wl, wr, BN))
    qpos = q_pos0 + q_stride * torch.arange(q_rows)
    for si, (row0, nrows, pos0, group) in enumerate(ksegs):
        for kt in range((nrows + BN - 1) // BN):
            nv = min(BN, nrows - kt * BN)
            kpos = pos0 + k_stride * (kt * BN + torch.arange(nv))
            rel = kpos.view(1, -1) - qpos.view(-1, 1)
            vis = torch.ones_like(rel, dtype=torch.bool)
            if wr >= 0:
                vis &= rel <= wr
            if wl >= 0:
                vis &= rel >= -wl
            if group != q_group:
                vis &= False
            if bool(vis.any()):
                assert (si, kt) in visited, (si, kt)


