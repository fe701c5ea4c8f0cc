"""CPU tests of the oracle itself: vs torch SDPA, autograd-consistency of the explicit backward,
merge exactness, block decomposition == whole."""
import math

import pytest
import torch
import torch.nn.functional as F

from lca_b200.ops.attention import AttnParams, attn_block_bwd, attn_block_fwd, merge_out_lse_
from lca_b200.ops.ref_attention import attention_ref, attn_block_bwd_ref, attn_block_fwd_ref
from lca_b200.parallel.layout import Seg, pos_tensor, ring_positions


def _qkv(B, Sq, Sk, H, Hkv, D, dtype=torch.float32, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(B, Sq, H, D, generator=g, dtype=dtype), torch.randn(B, Sk, Hkv, D, generator=g, dtype=dtype),
            torch.randn(B, Sk, Hkv, D, generator=g, dtype=dtype))


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("H,Hkv", [(4, 4), (4, 2), (4, 1)])
def test_matches_sdpa(causal, H, Hkv):
    q, k, v = _qkv(2, 33, 33, H, Hkv, 16)
    out, lse = attention_ref(q, k, v, causal=causal)
    kk, vv = (t.repeat_interleave(H // Hkv, dim=2) for t in (k, v))
    ref = F.scaled_dot_product_attention(q.transpose(1, 2), kk.transpose(1, 2), vv.transpose(1, 2), is_causal=causal)
    torch.testing.assert_close(out, ref.transpose(1, 2), atol=1e-5, rtol=1e-5)
    s = torch.einsum("bqhd,bkhd->bhqk", q, kk) / math.sqrt(16)
    if causal:
        s = s.masked_fill(torch.ones(33, 33, dtype=torch.bool).triu(1), float("-inf"))
    torch.testing.assert_close(lse, torch.logsumexp(s, -1), atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize("kw", [dict(causal=True), dict(window_size=(5, 3)), dict(causal=True, window_size=(7, -1)),
                                dict(softcap=3.0), dict(alibi=True, causal=True), dict()])
def test_explicit_backward_matches_autograd(kw):
    kw = dict(kw)
    q, k, v = _qkv(2, 24, 24, 4, 2, 8, torch.float64)
    alibi = torch.rand(4, dtype=torch.float64) if kw.pop("alibi", False) else None
    qp = kp = torch.arange(24)
    scale = 0.3
    q1, k1, v1 = (t.clone().requires_grad_() for t in (q, k, v))
    causal = kw.get("causal", False)
    ws = kw.get("window_size", (-1, -1))
    sc = kw.get("softcap", 0.0)
    # differentiable re-implementation through plain autograd
    kk, vv = (t.repeat_interleave(2, dim=2) for t in (k1, v1))
    s = torch.einsum("bqhd,bkhd->bhqk", q1, kk) * scale
    if sc > 0:
        s = sc * torch.tanh(s / sc)
    rel = kp.view(1, -1) - qp.view(-1, 1)
    if alibi is not None:
        s = s - alibi.view(1, -1, 1, 1) * rel.abs()
    m = torch.zeros(24, 24, dtype=torch.bool)
    if causal:
        m |= rel > 0
    if ws[0] >= 0:
        m |= rel < -ws[0]
    if ws[1] >= 0 and not causal:
        m |= rel > ws[1]
    s = s.masked_fill(m, float("-inf"))
    o_auto = torch.einsum("bhqk,bkhd->bqhd", torch.softmax(s, -1), vv)
    do = torch.randn_like(o_auto)
    o_auto.backward(do)
    out, lse = attn_block_fwd_ref(q.float(), k.float(), v.float(), qp, kp, scale, causal, ws, sc,
                                  None if alibi is None else alibi.float())
    torch.testing.assert_close(out.double(), o_auto.detach(), atol=1e-4, rtol=1e-4)
    dq, dk, dv = attn_block_bwd_ref(do.float(), q.float(), k.float(), v.float(), out, lse, qp, kp, scale, causal, ws, sc,
                                    None if alibi is None else alibi.float())
    torch.testing.assert_close(dq.double(), q1.grad, atol=1e-4, rtol=1e-3)
    torch.testing.assert_close(dk.double(), k1.grad, atol=1e-4, rtol=1e-3)
    torch.testing.assert_close(dv.double(), v1.grad, atol=1e-4, rtol=1e-3)


@pytest.mark.parametrize("variant", ["basic", "zigzag", "stripe"])
@pytest.mark.parametrize("kw", [dict(causal=True), dict(causal=True, window_size=(9, 0)), dict(causal=False)])
def test_block_decomposition_equals_whole(variant, kw):
    """Shard -> per-block attention with global positions -> merge == unsharded attention.  This is the
    single-process model of the ring and covers empty rows (-inf LSE) in the merge."""
    R, S, H, D = 4, 64, 2, 8
    q, k, v = _qkv(1, S, S, H, H, D)
    p = AttnParams.make(q, None, kw.get("causal", False), kw.get("window_size", (-1, -1)))
    ref_o, ref_l = attention_ref(q, k, v, **kw)
    for r in range(R):
        qpos = ring_positions(variant, r, R, S // R)
        qi = q[:, pos_tensor(qpos)]
        acc_o = torch.zeros(1, S // R, H, D)
        acc_l = torch.full((1, H, S // R), float("-inf"))
        for src in range(R):
            kpos = ring_positions(variant, src, R, S // R)
            idx = pos_tensor(kpos)
            bo, bl = attn_block_fwd(qi, k[:, idx], v[:, idx], qpos, kpos, p, "torch")
            merge_out_lse_(acc_o, acc_l, bo, bl)
        torch.testing.assert_close(acc_o, ref_o[:, pos_tensor(qpos)], atol=1e-5, rtol=1e-5)
        torch.testing.assert_close(acc_l, ref_l[:, :, pos_tensor(qpos)], atol=1e-5, rtol=1e-5)


def test_fully_masked_rows_are_zero_and_minus_inf():
    q, k, v = _qkv(1, 4, 4, 1, 1, 8)
    out, lse = attn_block_fwd_ref(q, k, v, torch.arange(4), torch.arange(4) + 10, 1.0, causal=True)
    assert torch.all(out == 0) and torch.all(torch.isinf(lse) & (lse < 0))
    acc_o, acc_l = torch.zeros(1, 4, 1, 8), torch.full((1, 1, 4), float("-inf"))
    merge_out_lse_(acc_o, acc_l, out, lse)
    assert not torch.isnan(acc_o).any() and torch.all(torch.isinf(acc_l))


def test_groups_block_diagonal():
    q, k, v = _qkv(1, 10, 10, 2, 2, 8)
    spec = (Seg(0, 4, 1, 0), Seg(0, 6, 1, 1))
    p = AttnParams.make(q, None, True)
    out, lse = attn_block_fwd(q, k, v, spec, spec, p, "torch")
    o0, _ = attention_ref(q[:, :4], k[:, :4], v[:, :4], causal=True)
    o1, _ = attention_ref(q[:, 4:], k[:, 4:], v[:, 4:], causal=True)
    torch.testing.assert_close(out, torch.cat([o0, o1], 1), atol=1e-5, rtol=1e-5)


def test_triple_merge_matches_lse_merge():
    """update_npu_out((out,max,sum) triples) == LSE merge (C9 parity)."""
    from lca_b200.ring import update_npu_out
    torch.manual_seed(0)
    B, S, N, D = 2, 5, 3, 4
    s1, s2 = torch.randn(B, N, S, 7), torch.randn(B, N, S, 6)
    v1, v2 = torch.randn(B, N, 7, D), torch.randn(B, N, 6, D)
    def part(s, v):
        m = s.max(-1, keepdim=True).values
        e = torch.exp(s - m)
        l = e.sum(-1, keepdim=True)
        return (e / l @ v).transpose(1, 2), m.expand(-1, -1, -1, 8), l.expand(-1, -1, -1, 8)
    o1, m1, l1 = part(s1, v1)
    o2, m2, l2 = part(s2, v2)
    out, _, _ = update_npu_out(o2, m2, l2, o1, m1, l1)
    ref = (torch.softmax(torch.cat([s1, s2], -1), -1) @ torch.cat([v1, v2], 2)).transpose(1, 2)
    torch.testing.assert_close(out, ref, atol=1e-5, rtol=1e-5)


def test_api_surface_matches_reference_names():
    import lca_b200 as y
    for name in ["LongContextAttention", "LongContextAttentionQKVPacked", "AsyncLongContextAttention", "UlyssesAttention",
                 "set_seq_parallel_pg", "EXTRACT_FUNC_DICT", "RING_IMPL_QKVPACKED_DICT", "ring_flash_attn_func",
                 "ring_flash_attn_kvpacked_func", "ring_flash_attn_qkvpacked_func", "zigzag_ring_flash_attn_func",
                 "zigzag_ring_flash_attn_kvpacked_func", "zigzag_ring_flash_attn_qkvpacked_func", "stripe_flash_attn_func",
                 "stripe_flash_attn_kvpacked_func", "stripe_flash_attn_qkvpacked_func", "ring_flash_attn_varlen_func",
                 "ring_flash_attn_varlen_kvpacked_func", "ring_flash_attn_varlen_qkvpacked_func",
                 "zigzag_ring_flash_attn_varlen_func", "zigzag_ring_flash_attn_varlen_qkvpacked_func",
                 "ring_pytorch_attn_func", "ring_flashinfer_attn_func", "ring_flashinfer_attn_kvpacked_func",
                 "ring_flashinfer_attn_qkvpacked_func", "ring_npu_flash_attn_func", "basic_extract_local",
                 "stripe_extract_local", "zigzag_extract_local", "__version__"]:
        assert hasattr(y, name), name
    from lca_b200.kernels import AttnType, select_flash_attn_impl
    assert {m.name for m in AttnType} >= {"AITER", "FA", "FA3", "FLASHINFER", "TORCH_MATH", "TORCH_FLASH", "TORCH_EFFICIENT",
                                          "TORCH_CUDNN", "SAGE_AUTO", "SAGE_FP16", "SAGE_FP16_TRITON", "SAGE_FP8",
                                          "SAGE_FP8_SM90", "SPARSE_SAGE", "NPU"}
    assert AttnType.from_string("torch") is AttnType.TORCH
    for stage in ("fwd-only", "bwd-only", "fwd-bwd"):
        assert callable(select_flash_attn_impl(AttnType.FA, stage))
        assert callable(select_flash_attn_impl(AttnType.TORCH_MATH, stage))
    import pytest
    with pytest.raises(ValueError):
        select_flash_attn_impl(AttnType.NPU)
    from lca_b200.comm import SeqAllToAll4D, SeqAllToAll5D, all_to_all_4D, all_to_all_5D   # noqa: F401
    from lca_b200.ring.utils import RingComm, update_out_and_lse, flatten_varlen_lse, unflatten_varlen_lse  # noqa: F401
    from lca_b200.globals import PROCESS_GROUP, HAS_FLASH_ATTN, HAS_FLASH_ATTN_HOPPER, HAS_FLASHINFER, HAS_NPU  # noqa: F401


# ------------------------------------------------------------------------------------------ dropout specification
def test_dropout_mask_statistics_and_keying():
    from lca_b200.ops import dropout as d
    assert d.p8_of(0.0) == 0 and d.p8_of(0.1) == 26 and d.p8_of(1.0) == 255
    assert abs(d.keep_scale(0.25) - 1.0 / 0.75) < 1e-12
    qp, kp = torch.arange(256), torch.arange(1024)
    m = d.keep_mask(7, 2, 3, qp, kp, 0.3)
    assert m.shape == (2, 3, 256, 1024)
    assert abs(m.float().mean().item() - (1 - d.p_eff(0.3))) < 5e-3
    f = m.float() - m.float().mean()
    for a, b in [(f[..., :-1], f[..., 1:]), (f[:, :, :-1], f[:, :, 1:]), (f[:, :-1], f[:, 1:]), (f[:1], f[1:])]:
        assert abs((a * b).mean().item()) < 3e-3                     # neighbours in k, q, head, batch are uncorrelated
    # pure function of the GLOBAL coordinates: any sub-block, any order, any head shard reproduces the same bits
    sub = d.keep_mask(7, 2, 1, qp[64:128], kp[512:700], 0.3, head_offset=2)
    assert torch.equal(sub[:, 0], m[:, 2, 64:128, 512:700])
    perm = torch.randperm(1024, generator=torch.Generator().manual_seed(0))
    assert torch.equal(d.keep_mask(7, 2, 3, qp, kp[perm], 0.3), m[..., perm])
    assert not torch.equal(d.keep_mask(8, 2, 3, qp, kp, 0.3), m)     # the seed matters
    assert bool(d.keep_mask(7, 1, 1, qp, kp, 0.0).all())


@pytest.mark.parametrize("causal", [False, True])
def test_single_device_dropout_matches_autograd_of_dense_formula(causal):
    """``flash_attn_func(dropout_p>0)`` (the reference's single-device entry point raised here before): forward and the
    hand-written backward against autograd through softmax -> mask -> rescale -> @V with the same keep mask."""
    from lca_b200.kernels.attention import pytorch_attn_func
    from lca_b200.ops import dropout as d
    B, S, H, D, pd, seed = 2, 40, 3, 8, 0.4, 99
    g = torch.Generator().manual_seed(3)
    q, k, v, do = (torch.randn(B, S, H, D, generator=g, dtype=torch.float64) for _ in range(4))
    q1, k1, v1 = (t.clone().requires_grad_() for t in (q, k, v))
    out = pytorch_attn_func(q1, k1, v1, dropout_p=pd, causal=causal, dropout_seed=seed)
    out.backward(do)
    q2, k2, v2 = (t.clone().requires_grad_() for t in (q, k, v))
    s = torch.einsum("bqhd,bkhd->bhqk", q2, k2) * D ** -0.5
    if causal:
        s = s.masked_fill(torch.ones(S, S, dtype=torch.bool).triu(1), float("-inf"))
    keep = d.keep_mask(seed, B, H, torch.arange(S), torch.arange(S), pd)
    pr = torch.softmax(s, dim=-1) * keep * d.keep_scale(pd)
    ref = torch.einsum("bhqk,bkhd->bqhd", pr, v2)
    ref.backward(do)
    torch.testing.assert_close(out, ref, atol=1e-5, rtol=1e-4)          # the engine computes in fp32
    for a, b in [(q1.grad, q2.grad), (k1.grad, k2.grad), (v1.grad, v2.grad)]:
        torch.testing.assert_close(a, b, atol=1e-5, rtol=1e-4)
