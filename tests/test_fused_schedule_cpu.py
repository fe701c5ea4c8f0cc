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
