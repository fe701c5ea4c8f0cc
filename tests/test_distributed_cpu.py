"""gloo multi-process tests of every distributed code path (BASELINE config 1 lives here).
Pattern follows the reference's tests (test/test_hybrid_attn.py: broadcast global tensors, shard with
EXTRACT_FUNC_DICT, run the module, compare against the single-device result sharded the same way) but
with hard forward AND backward assertions."""
import pytest
import torch
import torch.distributed as dist

from dist_utils import run_distributed


def _global_inputs(B, S, H, Hkv, D, seed=0, packed=False):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(B, S, H, D, generator=g)
    k = torch.randn(B, S, Hkv, D, generator=g)
    v = torch.randn(B, S, Hkv, D, generator=g)
    do = torch.randn(B, S, H, D, generator=g)
    return q, k, v, do


def _reference(q, k, v, do, **kw):
    from lca_b200.ops.ref_attention import attention_ref
    from lca_b200.kernels.attention import pytorch_attn_func
    q, k, v = (t.clone().requires_grad_() for t in (q, k, v))
    out = pytorch_attn_func(q, k, v, **kw)
    out.backward(do)
    return out.detach(), q.grad, k.grad, v.grad


# ------------------------------------------------------------------------------------------ a2a
def _a2a_worker(rank, world):
    from lca_b200.parallel.all_to_all import SeqAllToAll4D, SeqAllToAll5D
    B, Sl, H, D = 2, 3, 2 * world, 4
    full = torch.arange(B * Sl * world * H * D, dtype=torch.float32).view(B, Sl * world, H, D)
    x = full[:, rank * Sl:(rank + 1) * Sl].clone().requires_grad_()
    y = SeqAllToAll4D.apply(None, x, 2, 1)
    hl = H // world
    assert torch.equal(y, full[:, :, rank * hl:(rank + 1) * hl])            # token order + head ownership
    z = SeqAllToAll4D.apply(None, y, 1, 2)
    assert torch.equal(z, x)
    z.sum().backward()
    assert torch.equal(x.grad, torch.ones_like(x))
    full5 = torch.arange(B * Sl * world * 3 * H * D, dtype=torch.float32).view(B, Sl * world, 3, H, D)
    x5 = full5[:, rank * Sl:(rank + 1) * Sl].clone()
    y5 = SeqAllToAll5D.apply(None, x5, 3, 1)
    assert torch.equal(y5, full5[:, :, :, rank * hl:(rank + 1) * hl])
    assert torch.equal(SeqAllToAll5D.apply(None, y5, 1, 3), x5)


@pytest.mark.parametrize("world", [2, 4])
def test_all_to_all(world):
    run_distributed(_a2a_worker, world)


# ------------------------------------------------------------------------------------------ hybrid
def _hybrid_worker(rank, world, U, R, variant, kw, module, H, Hkv):
    import lca_b200
    from lca_b200 import (AsyncLongContextAttention, EXTRACT_FUNC_DICT, LongContextAttention,
                          LongContextAttentionQKVPacked, UlyssesAttention, set_seq_parallel_pg)
    from lca_b200.kernels import AttnType
    B, S, D = 1, 16 * world, 8
    q, k, v, do = _global_inputs(B, S, H, Hkv, D, seed=1)
    torch.manual_seed(4242)              # dropout seeds are drawn from torch's generator: same draw on every rank
    ro, rdq, rdk, rdv = _reference(q, k, v, do, **kw)
    torch.manual_seed(4242)
    set_seq_parallel_pg(U, R, rank, world)
    ex = EXTRACT_FUNC_DICT[variant]
    sh = lambda t: ex(t, rank, world, rd=R, ud=U).detach().clone()
    lq, lk, lv, ldo = sh(q), sh(k), sh(v), sh(do)
    for t in (lq, lk, lv):
        t.requires_grad_()
    if module == "hybrid":
        out = LongContextAttention(ring_impl_type=variant, attn_type=AttnType.TORCH)(lq, lk, lv, **kw)
    elif module == "hybrid_pack":
        out = LongContextAttention(ring_impl_type=variant, attn_type=AttnType.TORCH, use_pack_qkv=True)(lq, lk, lv, **kw)
    elif module == "async":
        out = AsyncLongContextAttention(ring_impl_type=variant, attn_type=AttnType.TORCH)(lq, lk, lv, **kw)
    elif module == "ulysses":
        out = UlyssesAttention(None, attn_type=AttnType.TORCH)(lq, lk, lv, **kw)
    elif module == "qkvpacked":
        qkv = torch.stack([lq, lk, lv], dim=2)
        out = LongContextAttentionQKVPacked(ring_impl_type=variant, attn_type=AttnType.TORCH)(qkv, **kw)
    out.backward(ldo)
    torch.testing.assert_close(out, sh(ro), atol=2e-5, rtol=1e-4)
    torch.testing.assert_close(lq.grad, sh(rdq), atol=2e-5, rtol=1e-4)
    torch.testing.assert_close(lk.grad, sh(rdk), atol=2e-5, rtol=1e-4)
    torch.testing.assert_close(lv.grad, sh(rdv), atol=2e-5, rtol=1e-4)


CAUSAL = dict(causal=True)


@pytest.mark.parametrize("U,R,variant,kw,module,H,Hkv", [
    (1, 2, "basic", CAUSAL, "hybrid", 8, 8),                 # BASELINE config 1 shape class (ring=2, ulysses=1)
    (1, 4, "zigzag", CAUSAL, "hybrid", 4, 4),
    (1, 4, "strip", CAUSAL, "hybrid", 4, 2),
    (2, 2, "zigzag", CAUSAL, "hybrid", 4, 2),                # GQA through Ulysses + ring
    (2, 2, "strip", dict(causal=True, window_size=(11, 0)), "hybrid", 4, 4),   # exact window across blocks
    (2, 2, "basic", dict(causal=False, window_size=(5, 9)), "hybrid", 4, 4),
    (4, 1, "basic", dict(causal=True, softcap=5.0), "hybrid", 4, 4),
    (2, 2, "zigzag", CAUSAL, "hybrid_pack", 4, 4),
    (2, 2, "zigzag", CAUSAL, "qkvpacked", 4, 4),
    (2, 2, "basic", dict(causal=False), "qkvpacked", 4, 4),
    (2, 2, "zigzag", CAUSAL, "async", 4, 2),
    (4, 1, "basic", CAUSAL, "ulysses", 8, 4),
])
def test_usp_modules_match_single_device(U, R, variant, kw, module, H, Hkv):
    run_distributed(_hybrid_worker, U * R, U, R, variant, kw, module, H, Hkv)


DROP = dict(causal=True, dropout_p=0.25)


@pytest.mark.parametrize("U,R,variant,kw,module,H,Hkv", [
    (1, 2, "basic", DROP, "hybrid", 4, 4),
    (1, 4, "zigzag", DROP, "hybrid", 4, 2),
    (2, 2, "strip", DROP, "hybrid", 4, 4),
    (2, 2, "zigzag", dict(causal=False, dropout_p=0.5, window_size=(9, 9)), "hybrid", 4, 2),
    (2, 2, "zigzag", DROP, "qkvpacked", 4, 4),
    (2, 2, "zigzag", DROP, "async", 4, 2),
    (4, 1, "basic", DROP, "ulysses", 8, 4),
])
def test_usp_dropout_equals_single_device_dropout(U, R, variant, kw, module, H, Hkv):
    """Dropout masks are functions of (seed, batch, global head, global q/k position): with the same seed every
    Ulysses x Ring layout reproduces the single-device result exactly, forward and backward (the reference draws an
    unrelated Philox stream per ring step and cannot re-create its masks in the ring backward)."""
    run_distributed(_hybrid_worker, U * R, U, R, variant, kw, module, H, Hkv)


def _baseline_config1_worker(rank, world):
    """BASELINE.json config 1: LongContextAttention basic, TORCH attn, ulysses=1 ring=2, gloo, S=1024 h=8 d=64."""
    from lca_b200 import EXTRACT_FUNC_DICT, LongContextAttention, set_seq_parallel_pg
    from lca_b200.kernels import AttnType
    q, k, v, do = _global_inputs(1, 1024, 8, 8, 64, seed=3)
    ro, rdq, rdk, rdv = _reference(q, k, v, do, causal=True)
    set_seq_parallel_pg(1, 2, rank, world)
    sh = lambda t: EXTRACT_FUNC_DICT["basic"](t, rank, world, rd=2, ud=1).detach().clone()
    lq, lk, lv = (sh(t).requires_grad_() for t in (q, k, v))
    out = LongContextAttention(ring_impl_type="basic", attn_type=AttnType.TORCH)(lq, lk, lv, causal=True)
    out.backward(sh(do))
    torch.testing.assert_close(out, sh(ro), atol=2e-5, rtol=1e-4)
    torch.testing.assert_close(lq.grad, sh(rdq), atol=5e-5, rtol=1e-4)
    torch.testing.assert_close(lk.grad, sh(rdk), atol=5e-5, rtol=1e-4)


def test_baseline_config1_cpu_gloo():
    run_distributed(_baseline_config1_worker, 2)


def _alibi_worker(rank, world):
    from lca_b200 import EXTRACT_FUNC_DICT, LongContextAttention, set_seq_parallel_pg
    from lca_b200.kernels import AttnType
    U, R = 2, 2
    q, k, v, do = _global_inputs(1, 64, 4, 4, 8, seed=5)
    slopes = torch.tensor([0.5, 0.25, 0.125, 0.0625])
    ro, rdq, _, _ = _reference(q, k, v, do, causal=True, alibi_slopes=slopes)
    set_seq_parallel_pg(U, R, rank, world)
    sh = lambda t: EXTRACT_FUNC_DICT["zigzag"](t, rank, world, rd=R, ud=U).detach().clone()
    lq, lk, lv = (sh(t).requires_grad_() for t in (q, k, v))
    out = LongContextAttention(ring_impl_type="zigzag", attn_type=AttnType.TORCH)(lq, lk, lv, causal=True, alibi_slopes=slopes)
    out.backward(sh(do))
    torch.testing.assert_close(out, sh(ro), atol=2e-5, rtol=1e-4)
    torch.testing.assert_close(lq.grad, sh(rdq), atol=2e-5, rtol=1e-4)


def test_alibi_through_usp():
    run_distributed(_alibi_worker, 4)


def _varlen_worker(rank, world, variant):
    from lca_b200 import ring_flash_attn_varlen_func, zigzag_ring_flash_attn_varlen_func
    from lca_b200.kernels import AttnType
    from lca_b200.kernels.attention import pytorch_attn_func
    from lca_b200.parallel.layout import local_token_index
    R, H, D = world, 2, 8
    glens = [8 * R, 4 * R, 12 * R]
    g = torch.Generator().manual_seed(7)
    seqs = [tuple(torch.randn(1, L, H, D, generator=g) for _ in range(4)) for L in glens]
    fn = ring_flash_attn_varlen_func if variant == "basic" else zigzag_ring_flash_attn_varlen_func
    loc, refs = [], []
    for (q, k, v, do) in seqs:
        q1, k1, v1 = (t.clone().requires_grad_() for t in (q, k, v))
        o = pytorch_attn_func(q1, k1, v1, causal=True)
        o.backward(do)
        idx = local_token_index(variant, q.shape[1], 0, rank, 1, R)
        loc.append(tuple(t[0, idx] for t in (q, k, v, do)))
        refs.append(tuple(t[0, idx] for t in (o.detach(), q1.grad, k1.grad, v1.grad)))
    lq, lk, lv, ldo = (torch.cat([l[i] for l in loc]).requires_grad_() for i in range(4))
    cu = torch.tensor([0] + list(torch.tensor([l[0].shape[0] for l in loc]).cumsum(0)), dtype=torch.int32)
    out = fn(lq, lk, lv, cu, max(l[0].shape[0] for l in loc), causal=True, attn_type=AttnType.TORCH)
    out.backward(ldo.detach())
    for i, name in enumerate(["out", "dq", "dk", "dv"]):
        got = [out, lq.grad, lk.grad, lv.grad][i]
        torch.testing.assert_close(got, torch.cat([r[i] for r in refs]), atol=2e-5, rtol=1e-4, msg=name)


@pytest.mark.parametrize("variant", ["basic", "zigzag"])
def test_varlen_ring(variant):
    run_distributed(_varlen_worker, 2, variant)


def _dp_worker(rank, world):
    """dp_degree > 1: two replicas of a 1x2 mesh shard identically (reference never tests this)."""
    from lca_b200 import EXTRACT_FUNC_DICT, LongContextAttention, PROCESS_GROUP, set_seq_parallel_pg
    from lca_b200.kernels import AttnType
    set_seq_parallel_pg(1, 2, rank, world)
    assert PROCESS_GROUP.mesh.dp_degree == 2
    q, k, v, do = _global_inputs(1, 32, 2, 2, 8, seed=9)
    ro, *_ = _reference(q, k, v, do, causal=True)
    sh = lambda t: EXTRACT_FUNC_DICT["strip"](t, rank, world, rd=2, ud=1)
    out = LongContextAttention(ring_impl_type="strip", attn_type=AttnType.TORCH)(sh(q), sh(k), sh(v), causal=True)
    torch.testing.assert_close(out, sh(ro), atol=2e-5, rtol=1e-4)


def test_dp_replicas():
    run_distributed(_dp_worker, 4)


def _ulysses_high_worker(rank, world):
    """use_ulysses_low=False: ring on the contiguous (low) ranks, Ulysses strided -- layouts follow mesh coordinates."""
    from lca_b200 import EXTRACT_FUNC_DICT, LongContextAttention, PROCESS_GROUP, set_seq_parallel_pg
    from lca_b200.kernels import AttnType
    U, R = 2, 2
    set_seq_parallel_pg(U, R, rank, world, use_ulysses_low=False)
    m = PROCESS_GROUP.mesh
    assert m.ring_group == ((0, 1) if rank < 2 else (2, 3)) and m.ulysses_group == ((0, 2) if rank % 2 == 0 else (1, 3))
    q, k, v, do = _global_inputs(1, 64, 4, 4, 8, seed=13)
    ro, rdq, _, _ = _reference(q, k, v, do, causal=True)
    sh = lambda t: EXTRACT_FUNC_DICT["zigzag"](t, rank, world, rd=R, ud=U, use_ulysses_low=False).detach().clone()
    lq, lk, lv = (sh(t).requires_grad_() for t in (q, k, v))
    out = LongContextAttention(ring_impl_type="zigzag", attn_type=AttnType.TORCH)(lq, lk, lv, causal=True)
    out.backward(sh(do))
    torch.testing.assert_close(out, sh(ro), atol=2e-5, rtol=1e-4)
    torch.testing.assert_close(lq.grad, sh(rdq), atol=2e-5, rtol=1e-4)


def test_use_ulysses_low_false():
    run_distributed(_ulysses_high_worker, 4)


def _dropout_worker(rank, world):
    """Dropout through the ring (PyTorch engine): the backward regenerates the forward's per-block masks, so the
    analytic gradient matches a finite-difference-free check: d(sum(out*w))/dv computed twice with the same seed agrees
    and differs from the no-dropout result (the reference's ring backward passes rng_state=None)."""
    from lca_b200 import EXTRACT_FUNC_DICT, set_seq_parallel_pg
    from lca_b200.kernels import AttnType
    from lca_b200.ring import ring_flash_attn_func
    from lca_b200.globals import PROCESS_GROUP
    set_seq_parallel_pg(1, 2, rank, world)
    q, k, v, do = _global_inputs(1, 32, 2, 2, 8, seed=21)
    sh = lambda t: EXTRACT_FUNC_DICT["basic"](t, rank, world, rd=2, ud=1).detach().clone()
    outs, grads = [], []
    for _ in range(2):
        torch.manual_seed(1234)          # same dropout seed draw on every repetition
        lq, lk, lv = (sh(t).requires_grad_() for t in (q, k, v))
        out = ring_flash_attn_func(lq, lk, lv, dropout_p=0.3, causal=True, group=PROCESS_GROUP.RING_PG,
                                   attn_type=AttnType.TORCH)
        out.backward(sh(do))
        outs.append(out.detach()); grads.append(lv.grad.clone())
    torch.testing.assert_close(outs[0], outs[1])
    torch.testing.assert_close(grads[0], grads[1])
    lq, lk, lv = (sh(t).requires_grad_() for t in (q, k, v))
    plain = ring_flash_attn_func(lq, lk, lv, causal=True, group=PROCESS_GROUP.RING_PG, attn_type=AttnType.TORCH)
    assert (plain - outs[0]).abs().max() > 1e-3
    # linearity in v under a fixed mask: out(v1 + v2) == out(v1) + out(v2)
    torch.manual_seed(1234)
    a = ring_flash_attn_func(sh(q), sh(k), sh(v) * 2.0, dropout_p=0.3, causal=True, group=PROCESS_GROUP.RING_PG,
                             attn_type=AttnType.TORCH)
    torch.testing.assert_close(a, outs[0] * 2.0, atol=1e-5, rtol=1e-5)


def test_ring_dropout_reproducible():
    run_distributed(_dropout_worker, 2)


def _packed_varlen_worker(rank, world):
    from lca_b200 import (ring_flash_attn_varlen_func, ring_flash_attn_varlen_kvpacked_func,
                          ring_flash_attn_varlen_qkvpacked_func)
    from lca_b200.kernels import AttnType
    g = torch.Generator().manual_seed(3)
    tot, H, D = 24, 2, 8
    q, k, v = (torch.randn(tot, H, D, generator=g) for _ in range(3))
    cu = torch.tensor([0, 8, 24], dtype=torch.int32)
    a = ring_flash_attn_varlen_func(q, k, v, cu, 16, causal=True, attn_type=AttnType.TORCH)
    b = ring_flash_attn_varlen_kvpacked_func(q, torch.stack([k, v], 1), cu, 16, causal=True, attn_type=AttnType.TORCH)
    c = ring_flash_attn_varlen_qkvpacked_func(torch.stack([q, k, v], 1), cu, 16, causal=True, attn_type=AttnType.TORCH)
    torch.testing.assert_close(a, b)
    torch.testing.assert_close(a, c)


def test_varlen_packed_wrappers():
    run_distributed(_packed_varlen_worker, 2)


# ------------------------------------------------------------------------------------------ reference-path modules
def _lowlevel_worker(rank, world):
    """``from yunchang.ring.<file> import <variant>_forward/_backward`` style entry points (positional reference
    signatures) against the single-device result, dense and varlen."""
    from lca_b200.comm.all_to_all import SeqAllToAll4D  # noqa: F401  (module paths of the reference must import)
    from lca_b200.comm.extract_local import zigzag_extract_local  # noqa: F401
    from lca_b200.kernels import AttnType
    from lca_b200.parallel.layout import local_token_index
    from lca_b200.ring.ring_flash_attn import ring_flash_attn_backward, ring_flash_attn_forward
    from lca_b200.ring.ring_flash_attn_varlen import ring_flash_attn_varlen_backward, ring_flash_attn_varlen_forward
    from lca_b200.ring.stripe_flash_attn import stripe_flash_attn_backward, stripe_flash_attn_forward
    from lca_b200.ring.zigzag_ring_flash_attn import zigzag_ring_flash_attn_backward, zigzag_ring_flash_attn_forward
    from lca_b200.ring.zigzag_ring_flash_attn_varlen import (get_half_index, zigzag_ring_flash_attn_varlen_backward,
                                                            zigzag_ring_flash_attn_varlen_forward)
    B, S, H, D = 1, 16 * world, 2, 8
    q, k, v, do = _global_inputs(B, S, H, H, D, seed=3)
    scale = D ** -0.5
    ro, rdq, rdk, rdv = _reference(q, k, v, do, causal=True)
    for variant, fwd, bwd in [("basic", ring_flash_attn_forward, ring_flash_attn_backward),
                              ("zigzag", zigzag_ring_flash_attn_forward, zigzag_ring_flash_attn_backward),
                              ("stripe", stripe_flash_attn_forward, stripe_flash_attn_backward)]:
        idx = local_token_index(variant, S, 0, rank, 1, world)
        lq, lk, lv, ldo = (t[:, idx].contiguous() for t in (q, k, v, do))
        out, lse = fwd(None, lq, lk, lv, scale, 0, True, (-1, -1), 0.0, None, False, AttnType.TORCH)
        assert lse.shape == (B, H, S // world)
        dq, dk, dv = bwd(None, ldo, lq, lk, lv, out, lse, scale, 0, True, (-1, -1), 0.0, None, False, AttnType.TORCH)
        for got, ref, name in [(out, ro, "out"), (dq, rdq, "dq"), (dk, rdk, "dk"), (dv, rdv, "dv")]:
            torch.testing.assert_close(got, ref[:, idx], atol=2e-5, rtol=1e-4, msg=f"{variant}:{name}")
    # varlen: one packed sequence per call keeps the bookkeeping short; cu_seqlens are LOCAL cumulative lengths
    for variant, fwd, bwd, extra in [
            ("basic", ring_flash_attn_varlen_forward, ring_flash_attn_varlen_backward, ()),
            ("zigzag", zigzag_ring_flash_attn_varlen_forward, zigzag_ring_flash_attn_varlen_backward, (None, None))]:
        idx = local_token_index(variant, S, 0, rank, 1, world)
        lq, lk, lv, ldo = (t[0, idx].contiguous() for t in (q, k, v, do))
        cu = torch.tensor([0, S // world], dtype=torch.int32)
        out, lse = fwd(None, lq, lk, lv, cu, S // world, *extra, scale, 0, True)
        assert out.shape == lq.shape and lse.shape == (H, S // world)
        dq, dk, dv = bwd(None, ldo, lq, lk, lv, out, lse, cu, S // world, *extra, scale, 0, True)
        for got, ref, name in [(out, ro, "out"), (dq, rdq, "dq"), (dk, rdk, "dk"), (dv, rdv, "dv")]:
            torch.testing.assert_close(got, ref[0, idx], atol=2e-5, rtol=1e-4, msg=f"varlen-{variant}:{name}")
    hi = get_half_index(torch.tensor([0, 8, 12]), front=False)
    assert hi.tolist() == [False] * 4 + [True] * 4 + [False] * 2 + [True] * 2
    assert get_half_index([0, 8], front=True) == slice(None, 4)


def test_reference_module_paths_and_lowlevel_ring_functions():
    run_distributed(_lowlevel_worker, 2)


def _slab_plan_worker(rank, world):
    """Ranks with DIFFERENT memory caps must take the same plan (the smallest), and a call that fits nowhere must be
    refused on every rank (collective fallback), decided once per call shape."""
    import types
    from lca_b200.parallel.fused_engine import FusedUSPEngine, _SlabDoesNotFit
    e = object.__new__(FusedUSPEngine)
    e.U, e.R, e.P, e.u, e.r, e.me = 1, world, world, 0, rank, rank
    e.group, e.device = None, torch.device("cpu")
    e.with_bwd, e._plan, e.slab, e.slab_bytes = True, {}, None, 0
    ensured = []
    e._ensure = types.MethodType(lambda self, *key: ensured.append(key), e)
    B, rows, H, Hkv, D = 1, 4096, 16, 8, 128
    q, k = torch.empty(B, rows, H, D, dtype=torch.bfloat16), torch.empty(B, rows, Hkv, D, dtype=torch.bfloat16)
    full = e.staging_bytes(B, rows, H, Hkv, D, 2, True)
    caps = [full + 1, full // 3][rank % 2]                     # rank 0 could take the whole call, rank 1 only a quarter of the heads
    e._cap_bytes = types.MethodType(lambda self: caps, e)
    c = e.reserve(q, k, True)
    assert c == 2, c                                            # 8 kv heads -> groups of 2 fit under a third of the slab
    assert ensured == [(B, rows, H // Hkv * 2, 2, D, 2, True)]
    assert e.reserve(q, k, True) == 2 and len(ensured) == 1     # cached: no second collective, no second layout
    # forward-only calls are planned separately (smaller layout: K/V + my ring block's Q)
    assert e.reserve(q, k, False) in (8, 4, 2)
    # nothing fits on one rank -> refused everywhere
    e._plan.clear()
    tiny = [full, 1024][rank % 2]
    e._cap_bytes = types.MethodType(lambda self: tiny, e)
    assert e.reserve(q, k, True) == 0
    # allocation failure after a positive plan -> refused as well
    e._plan.clear()
    e._cap_bytes = types.MethodType(lambda self: full + 1, e)

    def boom(self, *key):
        raise _SlabDoesNotFit("no memory")
    e._ensure = types.MethodType(boom, e)
    assert e.reserve(q, k, True) == 0


def test_fused_slab_plan_is_agreed_across_ranks():
    run_distributed(_slab_plan_worker, 2)
