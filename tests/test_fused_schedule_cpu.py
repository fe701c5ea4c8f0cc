"""Host-side logic of the fused NVLink engine, checked on CPU for every mesh shape: the segment lists handed to the
kernels (rows of the staging tensors + global positions + owners) must describe exactly the attention problem of
each rank.  The staging contents are emulated with torch (the push CTAs' row offsets: sp-rank-major gathered order)."""
import pytest
import torch

from lca_b200.ops.attention import AttnParams
from lca_b200.ops.ref_attention import attention_ref, attn_block_fwd_ref
from lca_b200.parallel.fused_engine import SIG_KV, SIG_Q, FusedUSPEngine
from lca_b200.parallel.layout import local_token_index


class _FakeSlab:
    def __init__(self, P):
        self.peer_ptrs = [1 << 20 | (i << 8) for i in range(P)]
        self.ptr = self.peer_ptrs[0]


def _engine(U, R, u, r):
    e = object.__new__(FusedUSPEngine)
    e.U, e.R, e.u, e.r, e.P, e.me = U, R, u, r, U * R, r * U + u
    e.slab, e.sig = _FakeSlab(U * R), _FakeSlab(U * R)
    e._segs = {}
    return e


@pytest.mark.parametrize("U,R", [(1, 2), (2, 1), (2, 2), (4, 2), (2, 4), (1, 8), (8, 1)])
@pytest.mark.parametrize("variant", ["basic", "zigzag", "stripe"])
def test_segments_describe_each_ranks_problem(U, R, variant):
    P, rows, H, D = U * R, 16, 2, 8
    S = P * rows
    g = torch.Generator().manual_seed(0)
    q, k, v = (torch.randn(1, S, H, D, generator=g) for _ in range(3))
    kw = dict(causal=True, window_size=(S // 3, 0))
    ref, ref_lse = attention_ref(q, k, v, **kw)
    p = AttnParams.make(q, None, True, (S // 3, 0))
    own = {(su, sr): local_token_index(variant, S, su, sr, U, R) for su in range(U) for sr in range(R)}
    # what the push CTAs build on every rank: rows ordered by sp-rank (sr*U + su), `rows` tokens each
    k_stage = torch.cat([k[:, own[(src % U, src // U)]] for src in range(P)], dim=1)
    v_stage = torch.cat([v[:, own[(src % U, src // U)]] for src in range(P)], dim=1)
    stride = R if variant == "stripe" else 1
    for r in range(R):
        q_stage = torch.cat([q[:, own[(su, r)]] for su in range(U)], dim=1)       # Ulysses gather of ring block r
        for u in range(U):
            e = _engine(U, R, u, r)
            qsegs, n_my_tiles = e._q_segments(variant, rows, U > 1, 0)
            ksegs = e._k_segments(variant, rows)
            assert len(qsegs) <= 32 and len(ksegs) <= 32
            # every staging row appears in exactly one segment
            assert sorted(x for s in ksegs for x in range(s[0], s[0] + s[1])) == list(range(P * rows))
            assert sorted(x for s in qsegs for x in range(s[0], s[0] + s[1])) == list(range(U * rows))
            kpos = torch.empty(P * rows, dtype=torch.int64)
            for row0, n, pos0, flag, grp in ksegs:
                kpos[row0:row0 + n] = pos0 + stride * torch.arange(n)
                assert flag == SIG_KV + row0 // rows                               # flag = source sp-rank of those rows
            qpos = torch.empty(U * rows, dtype=torch.int64)
            for seg in qsegs:
                row0, n, pos0, flag, o_row0 = seg[:5]
                qpos[row0:row0 + n] = pos0 + stride * torch.arange(n)
                su = row0 // rows
                assert o_row0 == row0 - su * rows
                if U > 1:
                    assert flag == SIG_Q + su and seg[5] == e.slab.peer_ptrs[r * U + su]
            out, lse = attn_block_fwd_ref(q_stage, k_stage, v_stage, qpos, kpos, p.softmax_scale, True, p.window_size)
            # rows of ring block r in gathered order are the tokens own[(su, r)] for su = 0..U-1
            tok = torch.cat([own[(su, r)] for su in range(U)])
            torch.testing.assert_close(out, ref[:, tok], atol=1e-5, rtol=1e-5)
            torch.testing.assert_close(lse, ref_lse[:, :, tok], atol=1e-5, rtol=1e-5)
            assert n_my_tiles == sum((s[1] + 127) // 128 for s in qsegs if s[0] // rows == u)


@pytest.mark.parametrize("U,R", [(1, 2), (2, 1), (2, 2), (4, 2), (2, 4), (1, 8), (8, 1), (1, 4)])
@pytest.mark.parametrize("variant", ["basic", "zigzag", "stripe"])
def test_owner_computes_backward_segments_cover_every_gradient_once(U, R, variant):
    """Host logic of the owner-computes backward (``FusedUSPEngine.backward``): the stationary/streamed segment lists
    of both passes and the owner scatter (``row0 - src*rows``) must produce every (token, head) of dQ, dK, dV exactly
    once and equal to the global gradient.  The all-rank staging (q/dO/k/v of every sp-rank in sp-rank-major order,
    head slice of this Ulysses rank) is emulated with torch."""
    from lca_b200.ops.ref_attention import attn_block_bwd_ref
    P, rows, H, D = U * R, 8, 2 * U, 8
    S, Hl = P * rows, H // U
    g = torch.Generator().manual_seed(1)
    q, k, v, do = (torch.randn(1, S, H, D, generator=g) for _ in range(4))
    win = (S // 2, 0)
    pos_all = torch.arange(S)
    p = AttnParams.make(q, None, True, win)
    out, lse = attn_block_fwd_ref(q, k, v, pos_all, pos_all, p.softmax_scale, True, win)
    rdq, rdk, rdv = attn_block_bwd_ref(do, q, k, v, out, lse, pos_all, pos_all, p.softmax_scale, True, win)
    own = {(su, sr): local_token_index(variant, S, su, sr, U, R) for su in range(U) for sr in range(R)}
    tok_of_stage = torch.cat([own[(src % U, src // U)] for src in range(P)])       # staging row -> global token id
    stride = R if variant == "stripe" else 1
    dq = torch.full((S, H, D), float("nan"))
    dk, dv = dq.clone(), dq.clone()
    written = {name: torch.zeros(S, H, dtype=torch.int32) for name in ("dq", "dk", "dv")}
    done_tiles = torch.zeros(P, dtype=torch.int64)
    for r in range(R):
        for u in range(U):
            e = _engine(U, R, u, r)
            all_segs, mine, tiles_of_me = e._bwd_segments(variant, rows)
            assert len(all_segs) <= 32
            assert sorted(x for (_, row0, n, _, _) in all_segs for x in range(row0, row0 + n)) == list(range(S))
            hs = slice(u * Hl, (u + 1) * Hl)
            stage = lambda t: t[:, tok_of_stage][:, :, hs]                            # noqa: E731
            qs, ks, vs, dos, outs = (stage(t) for t in (q, k, v, do, out))
            lses = lse[:, hs][:, :, tok_of_stage]
            spos = torch.empty(S, dtype=torch.int64)
            for (src, row0, n, pos0, _grp) in all_segs:
                spos[row0:row0 + n] = pos0 + stride * torch.arange(n)
                assert row0 // rows == src                                              # staging is sp-rank major
            assert torch.equal(spos, tok_of_stage)                                      # positions == token ids
            mrows = torch.cat([torch.arange(row0, row0 + n) for (_, row0, n, _, _) in mine])
            assert len(mrows) == U * rows and all(src // U == r for (src, _, _, _, _) in mine)
            # pass 1: dQ of my ring block (stationary) against everything (streamed)
            bdq, _, _ = attn_block_bwd_ref(dos[:, mrows], qs[:, mrows], ks, vs, outs[:, mrows], lses[:, :, mrows],
                                           spos[mrows], spos, p.softmax_scale, True, win)
            # pass 2: dK/dV of my ring block (stationary) against every query (streamed)
            _, bdk, bdv = attn_block_bwd_ref(dos, qs, ks[:, mrows], vs[:, mrows], outs, lses, spos, spos[mrows],
                                             p.softmax_scale, True, win)
            off = 0
            for (src, row0, n, pos0, _grp) in mine:
                o_row0 = row0 - src * rows                                              # row inside the owner's shard
                owner_tok = own[(src % U, src // U)][o_row0:o_row0 + n]
                assert torch.equal(owner_tok, tok_of_stage[row0:row0 + n])
                dq[owner_tok, hs] = bdq[0, off:off + n]
                dk[owner_tok, hs] = bdk[0, off:off + n]
                dv[owner_tok, hs] = bdv[0, off:off + n]
                for name in written:
                    written[name][owner_tok, hs] += 1
                done_tiles[src] += (n + 127) // 128
                off += n
            # what this rank waits for: tiles of its own tokens, produced by the U ranks of its ring block
            assert tiles_of_me == sum((n + 127) // 128 for (src, _, n, _, _) in all_segs if src == e.me)
    for name in written:
        assert bool((written[name] == 1).all()), name
    torch.testing.assert_close(dq, rdq[0], atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(dk, rdk[0], atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(dv, rdv[0], atol=1e-5, rtol=1e-5)
    # every owner receives U x (its tile count) completions per pass (one per Ulysses rank of the producing ring block)
    for src in range(P):
        e = _engine(U, R, src % U, src // U)
        assert done_tiles[src] == U * e._bwd_segments(variant, rows)[2]


@pytest.mark.parametrize("U,R", [(1, 2), (2, 2), (1, 8), (8, 1), (4, 2)])
@pytest.mark.parametrize("with_bwd", [False, True])
def test_slab_regions_are_disjoint_and_inside_the_allocation(U, R, with_bwd):
    """``FusedUSPEngine._layout``: every staging / output tensor the forward and both backward flavours place in the
    symmetric slab must lie inside the allocation and must not overlap a tensor that is live at the same time; a
    forward-only reservation (``torch.no_grad``) must be smaller than the training one."""
    for (B, rows, Hq_per_u, g, D, esz) in [(1, 8, 1, 1, 64, 2), (2, 1000, 2, 2, 128, 2), (1, 32768, 8 // min(U, 8) or 1, 1, 128, 2)]:
        H = Hq_per_u * U * g
        for Hkv in {H // g, max(1, U // 2)}:
            if H % Hkv or not (Hkv % U == 0 or U % Hkv == 0):
                continue
            e = object.__new__(FusedUSPEngine)
            e.U, e.R, e.P, e.u, e.r, e.me = U, R, U * R, 0, 0, 0
            offs, total = e._layout(B, rows, H, Hkv, D, esz, with_bwd)
            for name, off in offs.items():
                setattr(e, "off_" + name, off)
            assert e._layout(B, rows, H, Hkv, D, esz, False)[1] <= e._layout(B, rows, H, Hkv, D, esz, True)[1]
            P, S, Sr = U * R, U * R * rows, U * rows
            Hl, Hkvl = H // U, (Hkv // U if Hkv >= U else 1)
            fwd = {"q": (e.off_q, B * Sr * Hl * D * esz), "k": (e.off_k, B * S * Hkvl * D * esz),
                   "v": (e.off_v, B * S * Hkvl * D * esz), "o": (e.off_o, B * rows * H * D * esz),
                   "lse_own": (e.off_lse_own, B * H * rows * 4)}
            live_sets = [fwd]
            if with_bwd:
                live_sets.append({"q_all": (e.off_q, B * S * Hl * D * esz), "do_all": (e.off_do, B * S * Hl * D * esz),
                                  "k": fwd["k"], "v": fwd["v"], "dq": fwd["o"],
                                  "delta": (e.off_delta, B * Hl * S * 4), "lse2": (e.off_lse2, B * Hl * S * 4),
                                  "dk": (e.off_dk, B * rows * Hkv * D * 4), "dv": (e.off_dv, B * rows * Hkv * D * 4)})
            for regions in live_sets:
                spans = sorted((off, off + n, name) for name, (off, n) in regions.items())
                for (a0, a1, an), (b0, b1, bn) in zip(spans, spans[1:]):
                    assert a1 <= b0, f"{an} [{a0},{a1}) overlaps {bn} [{b0},{b1})"
                assert spans[-1][1] <= total
                assert all(off % 16 == 0 for off, _, _ in spans)


@pytest.mark.parametrize("R", [2, 4, 8])
@pytest.mark.parametrize("variant", ["basic", "zigzag"])
def test_varlen_segments_describe_packed_sequences(R, variant):
    """Packed variable-length shards on a ring-only mesh: the fused engine must hand the kernels one attention group
    per sequence, with the global positions of the per-sequence ring layout, and the union over all ranks must
    reproduce per-sequence causal attention exactly (emulated with the fp32 oracle on the segment description)."""
    H, D = 2, 8
    glens = [8 * R, 4 * R, 12 * R]                          # global lengths, each divisible by 2R
    g = torch.Generator().manual_seed(1)
    seqs = [tuple(torch.randn(1, L, H, D, generator=g) for _ in range(3)) for L in glens]
    refs = [attention_ref(q, k, v, causal=True)[0] for (q, k, v) in seqs]
    loc_idx = {rr: [local_token_index(variant, L, 0, rr, 1, R) for L in glens] for rr in range(R)}
    lens = [len(i) for i in loc_idx[0]]
    cu = [0]
    for n in lens:
        cu.append(cu[-1] + n)
    rows = cu[-1]
    # staging of rank r: every source rank's packed shard, sp-rank major
    pack = lambda rr, j: torch.cat([seqs[i][j][:, loc_idx[rr][i]] for i in range(len(glens))], dim=1)   # noqa: E731
    k_stage = torch.cat([pack(rr, 1) for rr in range(R)], dim=1)
    v_stage = torch.cat([pack(rr, 2) for rr in range(R)], dim=1)
    for r in range(R):
        e = _engine(1, R, 0, r)
        qsegs, _ = e._q_segments(variant, rows, False, 0, None, cu)
        ksegs = e._k_segments(variant, rows, cu)
        all_segs, mine, tiles = e._bwd_segments(variant, rows, cu)
        assert {s[4] for s in ksegs} == set(range(len(glens))) and len(all_segs) == len(ksegs)
        assert sorted(x for s in ksegs for x in range(s[0], s[0] + s[1])) == list(range(R * rows))
        assert sorted(x for s in qsegs for x in range(s[0], s[0] + s[1])) == list(range(rows))
        q_loc = pack(r, 0)
        kpos = torch.empty(R * rows, dtype=torch.int64)
        kgrp = torch.empty(R * rows, dtype=torch.int64)
        for row0, n, pos0, flag, grp in ksegs:
            kpos[row0:row0 + n] = pos0 + torch.arange(n)
            kgrp[row0:row0 + n] = grp
            assert flag == SIG_KV + row0 // rows
        qpos = torch.empty(rows, dtype=torch.int64)
        qgrp = torch.empty(rows, dtype=torch.int64)
        for seg in qsegs:
            row0, n, pos0 = seg[:3]
            qpos[row0:row0 + n] = pos0 + torch.arange(n)
            qgrp[row0:row0 + n] = seg[7]
        out = torch.zeros(1, rows, H, D)
        for gi in range(len(glens)):                      # one attention group per sequence
            qm, km = qgrp == gi, kgrp == gi
            o, _ = attn_block_fwd_ref(q_loc[:, qm], k_stage[:, km], v_stage[:, km], qpos[qm], kpos[km], D ** -0.5, True,
                                      (-1, -1))
            out[:, qm] = o
        want = torch.cat([refs[i][:, loc_idx[r][i]] for i in range(len(glens))], dim=1)
        torch.testing.assert_close(out, want, atol=1e-5, rtol=1e-5)


def test_head_group_candidates_keep_ulysses_and_gqa_divisibility():
    from lca_b200.parallel.fused_engine import head_chunk_candidates
    assert head_chunk_candidates(8, 1) == [8, 4, 2, 1]
    assert head_chunk_candidates(8, 2) == [8, 4, 2]
    assert head_chunk_candidates(8, 8) == [8]
    assert head_chunk_candidates(2, 8) == [2]            # replicated kv heads are never split
    assert head_chunk_candidates(12, 4) == [12, 4]
    for Hkv in range(1, 17):
        for U in (1, 2, 4, 8):
            for c in head_chunk_candidates(Hkv, U):
                assert Hkv % c == 0 and (c % U == 0 or Hkv % U != 0)


@pytest.mark.parametrize("U,R", [(2, 2), (4, 2), (2, 4), (1, 8), (8, 1)])
def test_staging_shrinks_with_the_head_group(U, R):
    """Bounded staging: the slab of a launch over c kv heads is ~c/Hkv of the whole call's slab."""
    e = _engine(U, R, 0, 0)
    B, rows, D, esz = 1, 4096, 128, 2
    Hkv, g = 8, 4
    from lca_b200.parallel.fused_engine import head_chunk_candidates
    full = e.staging_bytes(B, rows, g * Hkv, Hkv, D, esz, True)
    for c in head_chunk_candidates(Hkv, U):
        part = e.staging_bytes(B, rows, g * c, c, D, esz, True)
        assert part <= full * c / Hkv * 1.02 + 64 * 1024


@pytest.mark.parametrize("U,R", [(2, 2), (4, 2), (2, 4)])
def test_ulysses_high_mesh_permutes_only_the_pointer_tables(U, R):
    """use_ulysses_low=False: logical index d = r*U + u maps to group rank u*R + r (a permutation; identity otherwise)."""
    from lca_b200.parallel.mesh import coords_to_rank
    P = U * R
    for low in (True, False):
        table = [d if low else (d % U) * R + d // U for d in range(P)]
        assert sorted(table) == list(range(P))
        for r in range(R):
            for u in range(U):
                assert table[r * U + u] == coords_to_rank(u, r, 0, U, R, low)


@pytest.mark.parametrize("U,R", [(1, 4), (2, 2), (2, 4), (1, 8)])
@pytest.mark.parametrize("variant", ["basic", "zigzag", "stripe"])
@pytest.mark.parametrize("window", [(-1, -1), (24, 0)])
def test_push_masks_never_leave_out_a_needed_destination(U, R, variant, window):
    """Destination masks of the push CTAs: a bit may only be clear if NO (query of the destination's ring block, key of
    my shard) pair -- resp. (my query, key of its ring block) for the backward -- is visible under the exact mask; and a
    causal basic ring must actually skip the later-to-earlier direction."""
    P, rows = U * R, 16
    S = P * rows
    q = torch.zeros(1, S, 1, 8)
    p = AttnParams.make(q, None, True, window)
    own = {(su, sr): local_token_index(variant, S, su, sr, U, R) for su in range(U) for sr in range(R)}
    blk = {sr: torch.cat([own[(su, sr)] for su in range(U)]) for sr in range(R)}
    wl = window[0]

    def visible(qpos, kpos):
        rel = kpos.view(1, -1) - qpos.view(-1, 1)
        ok = rel <= 0
        if wl >= 0:
            ok &= rel >= -wl
        return bool(ok.any())

    skipped = 0
    for r in range(R):
        for u in range(U):
            e = _engine(U, R, u, r)
            for backward in (False, True):
                kvm, qm = e._push_masks(variant, rows, p, None, backward)
                for d in range(P):
                    dr = d // U
                    if not (kvm >> d) & 1:
                        skipped += 1
                        assert dr != r and not visible(blk[dr], own[(u, r)])
                    if not (qm >> d) & 1:
                        skipped += 1
                        assert backward and dr != r and not visible(own[(u, r)], blk[dr])
    if variant == "basic" and R > 1:
        assert skipped > 0


def test_head_group_launches_equal_one_launch(monkeypatch):
    """``FusedUSPEngine.attention`` with ``kv_heads_per_launch < Hkv``: the per-group slices of q / k / v / ALiBi slopes
    and the concatenation of outputs and LSE must reproduce the single-launch result (GQA groups stay aligned, the
    dropout head offset of a group is its first global query head).  The kernel launch is replaced by the PyTorch engine."""
    from lca_b200.parallel import fused_engine as fe
    from lca_b200.ops.attention import attn_block_fwd
    from lca_b200.parallel.layout import Seg

    seen = []

    def fake_apply(q, k, v, eng, variant, p, cu):
        seen.append((q.shape[2], k.shape[2], int(p.head_offset), None if p.alibi_slopes is None else p.alibi_slopes.clone()))
        pos = (Seg(0, q.shape[1], 1),)
        return attn_block_fwd(q, k, v, pos, pos, p, "torch")

    monkeypatch.setattr(fe._FusedAttnFunc, "apply", staticmethod(fake_apply))
    e = _engine(1, 2, 0, 0)
    g = torch.Generator().manual_seed(3)
    B, S, H, Hkv, D = 2, 48, 8, 4, 16
    q, k, v = (torch.randn(B, S, h, D, generator=g) for h in (H, Hkv, Hkv))
    slopes = torch.tensor([2.0 ** -(i + 1) for i in range(H)])
    kw = dict(softmax_scale=None, causal=True, window_size=(-1, -1), softcap=0.0, alibi_slopes=slopes, deterministic=False)
    whole, lse_w = fe.FusedUSPEngine.attention(e, q, k, v, "zigzag", return_lse=True, **kw)
    assert [s[:3] for s in seen] == [(8, 4, 0)]
    seen.clear()
    parts, lse_p = fe.FusedUSPEngine.attention(e, q, k, v, "zigzag", return_lse=True, kv_heads_per_launch=1, **kw)
    assert [s[:2] for s in seen] == [(2, 1)] * 4
    for i, s in enumerate(seen):
        torch.testing.assert_close(s[3], slopes[2 * i:2 * i + 2])
    torch.testing.assert_close(parts, whole)
    torch.testing.assert_close(lse_p, lse_w)
    seen.clear()
    fe.FusedUSPEngine.attention(e, q, k, v, "zigzag", kv_heads_per_launch=2, dropout_p=0.25, dropout_seed=7,
                                **{**kw, "alibi_slopes": None})
    assert [s[:3] for s in seen] == [(4, 2, 0), (4, 2, 4)]            # head offsets = first global query head of the group
