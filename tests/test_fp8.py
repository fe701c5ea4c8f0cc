"""FP8 path: the PyTorch emulation (CPU) tracks the fp32 oracle within fp8 tolerance; the CUDA kernel must match the
emulation; ``AttnType.SAGE_FP8*`` reach it through the modules and ``select_flash_attn_impl``."""
import os

import pytest
import torch

from lca_b200.ops.attention import AttnParams
from lca_b200.ops.fp8 import attn_fp8_emulated, attn_fp8_fwd, quantize_blockwise
from lca_b200.ops.ref_attention import attention_ref
from lca_b200.parallel.layout import Seg


def test_quantize_blockwise_roundtrip():
    torch.manual_seed(0)
    x = torch.randn(2, 300, 3, 32) * torch.linspace(0.1, 5, 300)[None, :, None, None]
    y, s = quantize_blockwise(x)
    assert y.dtype == torch.float8_e4m3fn and s.shape == (2, 3, 3)
    rows = s.repeat_interleave(128, dim=2)[:, :, :300].permute(0, 2, 1)
    back = y.to(torch.float32) * rows[..., None]
    assert ((back - x).abs() / (x.abs() + 1e-3)).median() < 0.04          # e4m3: 3 mantissa bits
    yh, sh = quantize_blockwise(x, per_head=True)
    assert sh.shape == (2, 3) and yh.to(torch.float32).abs().max() <= 448


@pytest.mark.parametrize("causal", [False, True])
def test_fp8_emulation_tracks_fp32(causal):
    torch.manual_seed(1)
    q, k, v = (torch.randn(1, 256, 4, 128) for _ in range(3))
    p = AttnParams.make(q, None, causal)
    pos = (Seg(0, 256, 1),)
    out, lse = attn_fp8_emulated(q, k[:, :, :2], v[:, :, :2], pos, pos, p)
    ro, rl = attention_ref(q, k[:, :, :2], v[:, :, :2], causal=causal)
    # rows with one or two visible keys expose the raw e4m3 rounding of V (~6 %), hence the loose max bound
    assert (out.float() - ro.float()).abs().max() < 0.25 and (out.float() - ro.float()).abs().mean() < 0.02
    torch.testing.assert_close(lse, rl, atol=0.08, rtol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("causal", [False, True])
def test_fp8_kernel_matches_emulation(causal):
    torch.manual_seed(2)
    q, k, v = (torch.randn(2, 1024, h, 128, device="cuda", dtype=torch.bfloat16) for h in (8, 2, 2))
    p = AttnParams.make(q, None, causal)
    pos = (Seg(0, 1024, 1),)
    out, lse = attn_fp8_fwd(q, k, v, pos, pos, p)
    eo, el = attn_fp8_emulated(q, k, v, pos, pos, p)
    assert (out.float() - eo.float()).abs().max() < 0.06
    torch.testing.assert_close(lse, el, atol=0.02, rtol=0)


@pytest.mark.gpu
def test_fp8_attn_type_runs_the_fp8_kernel_and_backpropagates_in_bf16():
    """SAGE_FP8 through the public single-device entry point: e4m3 forward (close to, but not identical with, the bf16
    forward), 16-bit backward on the saved operands."""
    from lca_b200.kernels import AttnType, select_flash_attn_impl
    torch.manual_seed(3)
    q, k, v = (torch.randn(1, 2048, h, 128, device="cuda", dtype=torch.bfloat16, requires_grad=True) for h in (4, 2, 2))
    fn8 = select_flash_attn_impl(AttnType.SAGE_FP8, stage="fwd-bwd")
    fn16 = select_flash_attn_impl(AttnType.FA, stage="fwd-bwd")
    o8 = fn8(q, k, v, causal=True)
    o16 = fn16(q, k, v, causal=True)
    d = (o8.float() - o16.float()).abs()
    assert 0 < d.max() < 0.25 and d.mean() < 0.02
    do = torch.randn_like(o8)
    g8 = torch.autograd.grad(o8, (q, k, v), do)
    g16 = torch.autograd.grad(o16, (q, k, v), do)
    for a, b in zip(g8, g16):
        assert (a.float() - b.float()).abs().max() / b.float().abs().max() < 0.1
    f8 = select_flash_attn_impl(AttnType.SAGE_FP8_SM90, stage="fwd-only")
    out, lse = f8(q.detach(), k.detach(), v.detach(), causal=True)
    assert lse.shape == (1, 4, 2048) and torch.equal(out, o8.detach())


def test_fp8_attn_type_on_cpu_uses_the_emulation():
    from lca_b200.kernels import AttnType, select_flash_attn_impl
    torch.manual_seed(4)
    q, k, v = (torch.randn(1, 256, h, 128, requires_grad=True) for h in (4, 2, 2))
    fn8 = select_flash_attn_impl(AttnType.SAGE_AUTO, stage="fwd-bwd")
    o8 = fn8(q, k, v, causal=True)
    ro, _ = attention_ref(q.detach(), k.detach(), v.detach(), causal=True)
    d = (o8.float() - ro.float()).abs()
    assert 0 < d.max() < 0.25 and d.mean() < 0.02          # quantised: close to, not equal to, the fp32 result
    g = torch.autograd.grad(o8, (q, k, v), torch.randn_like(o8))
    assert all(torch.isfinite(t).all() for t in g)
    os.environ["LCA_B200_FP8"] = "0"                          # switch: the same types fall back to the 16-bit engine
    try:
        o16 = select_flash_attn_impl(AttnType.SAGE_AUTO, stage="fwd-bwd")(q, k, v, causal=True)
    finally:
        del os.environ["LCA_B200_FP8"]
    assert (o16.float() - ro.float()).abs().max() < 1e-4


def test_sparse_sage_type_calls_the_users_processor():
    from lca_b200.kernels import AttnType, select_flash_attn_impl

    class Proc(torch.nn.Module):
        def forward(self, q, k, v, is_causal=False, scale=None, tensor_layout="NHD"):
            assert tensor_layout == "NHD"
            self.seen = (tuple(q.shape), is_causal, scale)
            return attention_ref(q, k, v, causal=is_causal, softmax_scale=scale)[0]

    proc = Proc()
    q, k, v = (torch.randn(1, 64, 2, 32) for _ in range(3))
    fn = select_flash_attn_impl(AttnType.SPARSE_SAGE, stage="fwd-only", attn_processor=proc)
    out, lse = fn(q, k, v, causal=True, softmax_scale=0.25)
    assert lse is None and proc.seen == ((1, 64, 2, 32), True, 0.25)
    torch.testing.assert_close(out, attention_ref(q, k, v, causal=True, softmax_scale=0.25)[0])
    with pytest.raises(ImportError):
        select_flash_attn_impl(AttnType.SPARSE_SAGE, stage="fwd-only")
    with pytest.raises(ValueError):
        select_flash_attn_impl(AttnType.SPARSE_SAGE, stage="fwd-bwd", attn_processor=proc)
    # through the module (world of one process: ring degree 1)
    from lca_b200 import LongContextAttention, set_seq_parallel_pg
    set_seq_parallel_pg(1, 1, 0, 1)
    attn = LongContextAttention(ring_impl_type="zigzag", attn_type=AttnType.SPARSE_SAGE, attn_processor=proc)
    torch.testing.assert_close(attn(q, k, v, causal=True), attention_ref(q, k, v, causal=True)[0])
