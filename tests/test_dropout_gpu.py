"""GPU test (B200) of the EXPERIMENTAL native dropout instantiations (``LCA_B200_NATIVE_DROPOUT=1``): the kernels must
regenerate exactly the keep mask of ``lca_b200/ops/dropout.py`` -- forward and both backward passes are compared with
the PyTorch engine evaluated on the same coordinates (so every kept/dropped score agrees, not just the statistics)."""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("LCA_B200_NATIVE_DROPOUT", "1") != "1",
                                 reason="native dropout kernels are opt-in until validated on hardware")]

CASES = [
    # B, S, H, Hkv, D, causal, q_pos spec, k_pos spec builder
    (1, 256, 2, 2, 128, True, "dense"),
    (2, 384, 4, 2, 64, False, "dense"),
    (1, 512, 2, 1, 128, True, "zigzag"),       # two segments per side, aligned
    (1, 256, 2, 2, 128, True, "stripe"),       # position stride 4: one hash per score
    (1, 300, 3, 3, 128, True, "offset"),       # segment starts at a position that is not a multiple of 4
    (1, 384, 2, 2, 128, True, "varlen"),       # two packed sequences (groups)
]


def _specs(kind, S):
    from lca_b200.parallel.layout import Seg
    if kind == "dense":
        return (Seg(0, S, 1),), (Seg(0, S, 1),)
    if kind == "zigzag":
        h = S // 2
        return (Seg(h, h, 1), Seg(7 * h, h, 1)), (Seg(h, h, 1), Seg(7 * h, h, 1))
    if kind == "stripe":
        return (Seg(1, S, 4),), (Seg(1, S, 4),)
    if kind == "offset":
        return (Seg(1001, S, 1),), (Seg(1001, S, 1),)
    if kind == "varlen":
        a = S // 3
        return (Seg(0, a, 1, 0), Seg(0, S - a, 1, 1)), (Seg(0, a, 1, 0), Seg(0, S - a, 1, 1))
    raise ValueError(kind)


@pytest.mark.parametrize("B,S,H,Hkv,D,causal,kind", CASES)
@pytest.mark.parametrize("pdrop", [0.1, 0.5])
def test_native_dropout_matches_spec(B, S, H, Hkv, D, causal, kind, pdrop):
    from dataclasses import replace
    from lca_b200.ops import native
    from lca_b200.ops.attention import AttnParams, attn_block_bwd, attn_block_fwd
    assert native.available()
    torch.manual_seed(0)
    q = torch.randn(B, S, H, D, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(B, S, Hkv, D, device="cuda", dtype=torch.bfloat16)
    v = torch.randn(B, S, Hkv, D, device="cuda", dtype=torch.bfloat16)
    do = torch.randn_like(q)
    qp, kp = _specs(kind, S)
    p = replace(AttnParams.make(q, None, causal, (-1, -1), 0.0, None, pdrop), dropout_seed=20240921, head_offset=5)
    assert native.dropout_supported(p)
    out, lse = native.fmha_fwd(q, k, v, qp, kp, p)
    ro, rl = attn_block_fwd(q, k, v, qp, kp, p, engine="torch")
    torch.testing.assert_close(out.float(), ro.float(), atol=3e-2, rtol=0)
    fin = torch.isfinite(rl)
    torch.testing.assert_close(lse[fin], rl[fin], atol=2e-3, rtol=1e-4)
    dq, dk, dv = native.fmha_bwd(do, q, k, v, out, lse, qp, kp, p)
    rq, rk, rv = attn_block_bwd(do, q, k, v, ro, rl, qp, kp, p, engine="torch")
    for got, ref, name in [(dq, rq, "dq"), (dk, rk, "dk"), (dv, rv, "dv")]:
        err = (got.float() - ref.float()).abs().max().item()
        assert err <= 0.05 * max(1.0, ref.float().abs().max().item()), f"{name}: max err {err}"
    # a different seed must change the result (the mask is really applied)
    out2, _ = native.fmha_fwd(q, k, v, qp, kp, replace(p, dropout_seed=1))
    assert (out2.float() - out.float()).abs().max().item() > 1e-2
