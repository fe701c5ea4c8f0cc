"""Multi-GPU tests of the fused NVLink USP kernel (needs >= 2 B200s: `gpurun --gpus 2|4|8`)."""
import os

import pytest
import torch

from dist_utils import run_distributed

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def _ngpu():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _worker(rank, world, U, R, variant, kw, H, Hkv, S, D, module, check_bwd):
    import lca_b200
    from lca_b200 import EXTRACT_FUNC_DICT, LongContextAttention, UlyssesAttention, set_seq_parallel_pg
    from lca_b200.kernels.attention import pytorch_attn_func
    dev = torch.device("cuda", rank)
    g = torch.Generator().manual_seed(11)
    B = 1 if kw.pop("_b1", True) else 2
    q = torch.randn(B, S, H, D, generator=g).to(dev, torch.bfloat16)
    k = torch.randn(B, S, Hkv, D, generator=g).to(dev, torch.bfloat16)
    v = torch.randn(B, S, Hkv, D, generator=g).to(dev, torch.bfloat16)
    do = torch.randn(B, S, H, D, generator=g).to(dev, torch.bfloat16)
    q1, k1, v1 = (t.clone().requires_grad_() for t in (q, k, v))
    ref = pytorch_attn_func(q1, k1, v1, **kw)
    ref.backward(do)
    set_seq_parallel_pg(U, R, rank, world)
    key = {"basic": "basic", "zigzag": "zigzag", "stripe": "strip"}[variant]
    sh = lambda t: EXTRACT_FUNC_DICT[key](t, rank, world, rd=R, ud=U).detach().clone()
    lq, lk, lv = (sh(t).requires_grad_() for t in (q, k, v))
    if module == "ulysses":
        attn = UlyssesAttention(None, backend="fused")
    else:
        attn = LongContextAttention(ring_impl_type=key, backend="fused")
    for it in range(3):                       # several calls: epochs / staging reuse / o_done accumulation
        out = attn(lq, lk, lv, **kw)
        torch.testing.assert_close(out.float(), sh(ref.detach()).float(), atol=2e-2, rtol=0, msg=f"call {it}")
    if check_bwd:
        out.backward(sh(do))
        for a, b, n in ((lq.grad, q1.grad, "dq"), (lk.grad, k1.grad, "dk"), (lv.grad, v1.grad, "dv")):
            ref_g = sh(b).float()
            err = (a.float() - ref_g).abs().max().item()
            assert err / (ref_g.abs().max().item() + 1e-6) < 3e-2, f"{n}: {err}"
    torch.cuda.synchronize()


CASES2 = [
    # U, R, variant, kwargs, H, Hkv, S, D, module, check_bwd
    (2, 1, "basic", dict(causal=True), 4, 4, 1024, 128, "hybrid", True),
    (1, 2, "zigzag", dict(causal=True), 4, 2, 2048, 128, "hybrid", True),
    (1, 2, "stripe", dict(causal=True, window_size=(300, 0)), 2, 2, 1024, 64, "hybrid", False),
    (1, 2, "basic", dict(causal=False), 2, 2, 1024, 128, "hybrid", False),
    (2, 1, "basic", dict(causal=True), 8, 2, 1024, 128, "ulysses", True),
    (2, 1, "basic", dict(causal=True, softcap=10.0), 4, 1, 512, 128, "hybrid", False),   # MQA: kv heads < U
]


@pytest.mark.parametrize("U,R,variant,kw,H,Hkv,S,D,module,check_bwd", CASES2)
def test_fused_2gpu(U, R, variant, kw, H, Hkv, S, D, module, check_bwd):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    run_distributed(_worker, 2, U, R, variant, dict(kw), H, Hkv, S, D, module, check_bwd, backend="nccl")


CASES4 = [
    (2, 2, "zigzag", dict(causal=True), 4, 2, 2048, 128, "hybrid", True),
    (4, 1, "basic", dict(causal=True), 8, 8, 2048, 128, "hybrid", False),
    (1, 4, "zigzag", dict(causal=True, window_size=(700, 0)), 2, 2, 4096, 128, "hybrid", False),
    (2, 2, "stripe", dict(causal=True), 4, 4, 2048, 64, "hybrid", False),
]


@pytest.mark.parametrize("U,R,variant,kw,H,Hkv,S,D,module,check_bwd", CASES4)
def test_fused_4gpu(U, R, variant, kw, H, Hkv, S, D, module, check_bwd):
    if _ngpu() < 4:
        pytest.skip("needs 4 GPUs")
    run_distributed(_worker, 4, U, R, variant, dict(kw), H, Hkv, S, D, module, check_bwd, backend="nccl")


CASES8 = [
    (8, 1, "basic", dict(causal=True), 32, 8, 8192, 128, "hybrid", False),
    (1, 8, "zigzag", dict(causal=True), 8, 8, 16384, 128, "hybrid", True),
    (2, 4, "zigzag", dict(causal=True, window_size=(3000, 0)), 8, 4, 8192, 128, "hybrid", False),
    (4, 2, "zigzag", dict(causal=True), 16, 16, 8192, 128, "hybrid", False),
]


@pytest.mark.parametrize("U,R,variant,kw,H,Hkv,S,D,module,check_bwd", CASES8)
def test_fused_8gpu(U, R, variant, kw, H, Hkv, S, D, module, check_bwd):
    if _ngpu() < 8:
        pytest.skip("needs 8 GPUs")
    run_distributed(_worker, 8, U, R, variant, dict(kw), H, Hkv, S, D, module, check_bwd, backend="nccl")


def _collective_worker(rank, world, U, R, variant):
    """backend="collective": NCCL all-to-all + P2P ring around the native kernels (multi-node capable path)."""
    import lca_b200
    from lca_b200 import EXTRACT_FUNC_DICT, LongContextAttention, set_seq_parallel_pg
    from lca_b200.kernels.attention import pytorch_attn_func
    dev = torch.device("cuda", rank)
    g = torch.Generator().manual_seed(5)
    B, S, H, Hkv, D = 1, 1024, 4, 2, 128
    q, k, v, do = (torch.randn(B, S, h, D, generator=g).to(dev, torch.bfloat16) for h in (H, Hkv, Hkv, H))
    q1, k1, v1 = (t.clone().requires_grad_() for t in (q, k, v))
    ref = pytorch_attn_func(q1, k1, v1, causal=True)
    ref.backward(do)
    set_seq_parallel_pg(U, R, rank, world)
    sh = lambda t: EXTRACT_FUNC_DICT[variant](t, rank, world, rd=R, ud=U).detach().clone()
    lq, lk, lv = (sh(t).requires_grad_() for t in (q, k, v))
    attn = LongContextAttention(ring_impl_type=variant, backend="collective")
    for _ in range(2):
        lq.grad = lk.grad = lv.grad = None
        out = attn(lq, lk, lv, causal=True)
        out.backward(sh(do))
    torch.testing.assert_close(out.float(), sh(ref.detach()).float(), atol=2e-2, rtol=0)
    for a, b in ((lq.grad, q1.grad), (lk.grad, k1.grad), (lv.grad, v1.grad)):
        rg = sh(b).float()
        assert (a.float() - rg).abs().max().item() / (rg.abs().max().item() + 1e-6) < 3e-2
    torch.cuda.synchronize()


@pytest.mark.parametrize("U,R,variant", [(1, 2, "zigzag"), (2, 1, "basic")])
def test_collective_backend_2gpu(U, R, variant):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    run_distributed(_collective_worker, 2, U, R, variant, backend="nccl", timeout=90)


@pytest.mark.parametrize("U,R,variant", [(2, 1, "basic"), (1, 2, "zigzag")])
def test_fused_batch2_2gpu(U, R, variant):
    """B = 2 through the fused path (staging batch strides, counters scale with B)."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    run_distributed(_worker, 2, U, R, variant, dict(causal=True, _b1=False), 4, 2, 1024, 128, "hybrid", True,
                    backend="nccl", timeout=90)


def _dropout_worker(rank, world, U, R, variant):
    """Fused USP kernels with native dropout == single-device dropout with the same seed (coordinate-keyed masks)."""
    from lca_b200 import EXTRACT_FUNC_DICT, LongContextAttention, set_seq_parallel_pg
    from lca_b200.kernels.attention import pytorch_attn_func
    dev = torch.device("cuda", rank)
    B, S, H, Hkv, D = 1, 1024, 4, 2, 128
    g = torch.Generator().manual_seed(5)
    q, k, v, do = (torch.randn(B, S, h, D, generator=g).to(dev, torch.bfloat16) for h in (H, Hkv, Hkv, H))
    kw = dict(causal=True, dropout_p=0.2)
    q1, k1, v1 = (t.clone().requires_grad_() for t in (q, k, v))
    torch.manual_seed(777)                               # the dropout seed is drawn from torch's CPU generator
    ref = pytorch_attn_func(q1, k1, v1, **kw)
    ref.backward(do)
    set_seq_parallel_pg(U, R, rank, world)
    key = {"basic": "basic", "zigzag": "zigzag", "stripe": "strip"}[variant]
    sh = lambda t: EXTRACT_FUNC_DICT[key](t, rank, world, rd=R, ud=U).detach().clone()
    lq, lk, lv = (sh(t).requires_grad_() for t in (q, k, v))
    attn = LongContextAttention(ring_impl_type=key, backend="fused")
    torch.manual_seed(777)
    out = attn(lq, lk, lv, **kw)
    torch.testing.assert_close(out.float(), sh(ref.detach()).float(), atol=3e-2, rtol=0)
    out.backward(sh(do))
    for a, b, n in ((lq.grad, q1.grad, "dq"), (lk.grad, k1.grad, "dk"), (lv.grad, v1.grad, "dv")):
        ref_g = sh(b).float()
        err = (a.float() - ref_g).abs().max().item()
        assert err / (ref_g.abs().max().item() + 1e-6) < 4e-2, f"{n}: {err}"
    torch.cuda.synchronize()


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
@pytest.mark.skipif(os.environ.get("LCA_B200_NATIVE_DROPOUT", "0") != "1", reason="native dropout kernels are opt-in")
@pytest.mark.parametrize("U,R,variant", [(1, 2, "zigzag"), (2, 1, "basic")])
def test_fused_dropout_2gpu(U, R, variant):
    run_distributed(_dropout_worker, 2, U, R, variant, backend="nccl")
