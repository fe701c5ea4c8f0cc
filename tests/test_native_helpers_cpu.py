"""Host-side helpers of the native path that do not need a GPU."""
import torch

from lca_b200.ops import native
from lca_b200.ops.attention import AttnParams, block_is_visible
from lca_b200.parallel.layout import Seg, ring_positions, varlen_positions


def test_window_bounds_fold_causal_into_right_bound():
    q = torch.zeros(1, 4, 1, 8)
    assert native.window_bounds(AttnParams.make(q, None, True)) == (-1, 0)
    assert native.window_bounds(AttnParams.make(q, None, True, (7, -1))) == (7, 0)
    assert native.window_bounds(AttnParams.make(q, None, False, (3, 5))) == (3, 5)
    assert native.window_bounds(AttnParams.make(q, None, False)) == (-1, -1)


def test_rows_and_strides():
    spec = ring_positions("zigzag", 1, 4, 8)
    assert native._rows(spec) == [(0, 4, 4, 0), (4, 4, 24, 0)]
    assert native._common_stride(ring_positions("stripe", 2, 4, 6)) == 4
    assert native._common_stride(spec) == 1


def test_chunk_by_group_splits_many_sequences():
    cu = list(range(0, 41 * 16, 16))                    # 40 sequences of 16 tokens -> 40 groups
    spec = varlen_positions("basic", 0, 1, cu)
    rows = native._rows(spec)
    chunks = list(native._chunk_by_group(rows, rows))
    assert len(chunks) == 2 and all(len(q) <= native.MAX_SEG and len(k) <= native.MAX_SEG for q, k in chunks)
    assert sorted(r for q, _ in chunks for r in q) == sorted(rows)
    for q, k in chunks:                                  # a chunk holds whole groups on both sides
        assert {r[3] for r in q} == {r[3] for r in k}


def test_block_visibility_matches_reference_step_rule():
    """basic causal ring: step visible iff source rank <= my rank (ring_flash_attn.py:35); zigzag: always."""
    q = torch.zeros(1, 8, 1, 8)
    p = AttnParams.make(q, None, True)
    R, L = 4, 8
    for r in range(R):
        for src in range(R):
            vis = block_is_visible(ring_positions("basic", r, R, L), ring_positions("basic", src, R, L), p)
            assert vis == (src <= r)
            assert block_is_visible(ring_positions("zigzag", r, R, L), ring_positions("zigzag", src, R, L), p)
    pw = AttnParams.make(q, None, True, (3, 0))          # window 3: only the adjacent earlier block matters
    assert not block_is_visible((Seg(24, 8),), (Seg(0, 8),), pw)
    assert block_is_visible((Seg(24, 8),), (Seg(16, 8),), pw)


def test_padded_dim_and_support_messages():
    assert native._padded_dim(32) == 64 and native._padded_dim(96) == 128 and native._padded_dim(128) == 128
    assert "not on CUDA" in native.why_not(torch.zeros(1, 1, 1, 64))
