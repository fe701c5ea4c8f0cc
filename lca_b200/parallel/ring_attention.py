"""Ring attention over a process group -- collective (NCCL/gloo P2P) path, all variants.

Parity: ``yunchang/ring/ring_flash_attn.py`` (basic), ``zigzag_ring_flash_attn.py``,
``stripe_flash_attn.py`` -- forward loops, backward loops with the travelling fp32 dK/dV ring,
and the ``autograd.Function`` wrappers (``ring_flash_attn.py:7-224`` etc.).

Design difference: the three reference files hard-code their masking cases per ring step
(``step <= rank`` / half-slicing / shifted slicing).  Here ONE loop serves every variant because
each block is described by the *global positions* of its tokens
(:func:`lca_b200.parallel.layout.ring_positions`) and the attention op masks on positions.
Consequences: sliding windows and ALiBi are exact across blocks (the reference is only right
for ``window=(-1,-1)`` and asserts ``alibi_slopes is None``), non-causal zigzag/stripe work, and
fully-masked blocks are skipped by a position-range test instead of variant-specific rules.
The zigzag/stripe load balance is preserved because the kernel skips fully masked tiles.
"""
from __future__ import annotations

from dataclasses import replace
from typing import Optional

import torch

from ..globals import group_rank, group_size
from ..ops import native
from ..ops.attention import (AttnParams, attn_block_bwd, attn_block_fwd, block_is_visible,
                             merge_out_lse_)
from .layout import canonical_variant, ring_positions, varlen_positions
from .ring_comm import RingComm


def _engine_for(attn_type) -> Optional[str]:
    """AttnType -> engine name (None = auto: native on Blackwell, torch elsewhere)."""
    if attn_type is None:
        return None
    name = getattr(attn_type, "value", attn_type)
    if isinstance(name, str) and name.startswith("torch"):
        return "torch"
    if name in ("sage_fp8", "sage_fp8_sm90", "sage_auto"):      # the quantised forward family -> e4m3 tcgen05 forward
        return "fp8"
    return None


def _pos_builders(variant, R, Lq, Lk, cu_seqlens_q=None, cu_seqlens_k=None):
    """-> (qpos(rank), kpos(rank)) closures; dense shards or packed varlen shards."""
    if cu_seqlens_q is None:
        return (lambda rr: ring_positions(variant, rr, R, Lq)), (lambda rr: ring_positions(variant, rr, R, Lk))
    cq = [int(x) for x in cu_seqlens_q]
    ck = cq if cu_seqlens_k is None else [int(x) for x in cu_seqlens_k]
    return (lambda rr: varlen_positions(variant, rr, R, cq)), (lambda rr: varlen_positions(variant, rr, R, ck))


def ring_attn_forward(group, q, k, v, variant: str, p: AttnParams, engine=None, dropout_seed: int = 0,
                      cu_seqlens_q=None, cu_seqlens_k=None):
    """-> out (B,Sq,H,D) q.dtype, lse (B,H,Sq) fp32.  With ``cu_seqlens`` the tensors are packed
    ``(1, total, H, D)`` shards and every sequence is its own attention group."""
    R, r = group_size(group), group_rank(group)
    B, Lq, H, D = q.shape
    Lk = k.shape[1]
    qpos_of, kpos_of = _pos_builders(variant, R, Lq, Lk, cu_seqlens_q, cu_seqlens_k)
    q_pos = qpos_of(r)
    if p.dropout_p > 0.0 and dropout_seed:
        p = replace(p, dropout_seed=int(dropout_seed))       # masks are functions of global coordinates (ops/dropout.py)
    if R == 1:
        return attn_block_fwd(q, k, v, q_pos, kpos_of(0), p, engine)
    comm = RingComm(group)
    k, v = k.contiguous(), v.contiguous()      # P2P payloads must be dense (packed-QKV views are not)
    out_acc = lse_acc = None
    next_k = next_v = None
    for step in range(R):
        if step + 1 != R:
            next_k, next_v = comm.send_recv(k), comm.send_recv(v)
            comm.commit()
        src = (r - step) % R
        k_pos = kpos_of(src)
        if block_is_visible(q_pos, k_pos, p):
            bo, bl = attn_block_fwd(q, k, v, q_pos, k_pos, p, engine)
            if out_acc is None:
                out_acc, lse_acc = bo.to(torch.float32), bl
            else:
                merge_out_lse_(out_acc, lse_acc, bo, bl)
        if step + 1 != R:
            comm.wait()
            k, v = next_k, next_v
    if out_acc is None:  # nothing visible at all (e.g. window excludes every block)
        return (torch.zeros_like(q), torch.full((B, H, Lq), float("-inf"), dtype=torch.float32, device=q.device))
    out = native.finalize_out(out_acc, q.dtype) if out_acc.is_cuda else out_acc.to(q.dtype)
    return out, lse_acc


def ring_attn_backward(group, dout, q, k, v, out, lse, variant: str, p: AttnParams, engine=None,
                       dropout_seed: int = 0, cu_seqlens_q=None, cu_seqlens_k=None):
    """-> (dq, dk, dv) in the input dtypes.  dK/dV partial sums travel with their K/V block in fp32
    and arrive at the owner after R hops (same scheme as ``ring_flash_attn.py:65-147``)."""
    R, r = group_size(group), group_rank(group)
    B, Lq, H, D = q.shape
    Lk = k.shape[1]
    qpos_of, kpos_of = _pos_builders(variant, R, Lq, Lk, cu_seqlens_q, cu_seqlens_k)
    q_pos = qpos_of(r)
    lse2 = None
    if q.is_cuda and native.available() and dout.dtype in (torch.bfloat16, torch.float16):
        delta, lse2 = native.attn_delta(out, dout, lse)     # one fused pass: rowsum(dO o O) + log2-domain LSE
    else:
        delta = (dout.to(torch.float32) * out.to(torch.float32)).sum(-1).permute(0, 2, 1).contiguous()
    if p.dropout_p > 0.0 and dropout_seed:
        p = replace(p, dropout_seed=int(dropout_seed))
    if R == 1:
        dq, dk, dv = attn_block_bwd(dout, q, k, v, out, lse, q_pos, kpos_of(0), p, engine, None, delta, lse2)
        return dq.to(q.dtype), dk.to(k.dtype), dv.to(v.dtype)
    kv_comm, dkv_comm = RingComm(group), RingComm(group)
    k, v = k.contiguous(), v.contiguous()
    f32 = dict(dtype=torch.float32, device=q.device)
    dq = torch.empty(q.shape, **f32)
    dq_live = False
    dk_acc = dv_acc = next_dk = next_dv = next_k = next_v = None
    for step in range(R):
        if step + 1 != R:
            next_k, next_v = kv_comm.send_recv(k), kv_comm.send_recv(v)
            kv_comm.commit()
        src = (r - step) % R
        k_pos = kpos_of(src)
        if step > 0:
            dkv_comm.wait()                       # partial dK/dV of the block we now hold (from the previous rank)
            dk_acc, dv_acc = next_dk, next_dv
        if block_is_visible(q_pos, k_pos, p):
            if step == 0:
                dk_acc, dv_acc = torch.empty(k.shape, **f32), torch.empty(v.shape, **f32)
            # the kernels write / accumulate straight into the fp32 buffers (no extra passes)
            attn_block_bwd(dout, q, k, v, out, lse, q_pos, k_pos, p, engine, None, delta, lse2,
                           into=(dq, dk_acc, dv_acc, dq_live, step > 0))
            dq_live = True
        elif step == 0:
            dk_acc, dv_acc = torch.zeros(k.shape, **f32), torch.zeros(v.shape, **f32)
        if step + 1 != R:
            kv_comm.wait()
            k, v = next_k, next_v
        next_dk, next_dv = dkv_comm.send_recv(dk_acc), dkv_comm.send_recv(dv_acc)
        dkv_comm.commit()
    dkv_comm.wait()
    if not dq_live:
        dq.zero_()
    return dq.to(q.dtype), next_dk.to(k.dtype), next_dv.to(v.dtype)


class RingAttnFunc(torch.autograd.Function):
    """One autograd node for every ring variant (reference: RingFlashAttnFunc /
    ZigZagRingFlashAttnFunc / StripeFlashAttnFunc)."""

    @staticmethod
    def forward(ctx, q, k, v, variant, dropout_p, softmax_scale, causal, window_size, softcap, alibi_slopes,
                deterministic, return_softmax, group, attn_type, attn_processor, head_offset=0, dropout_seed=None):
        p = AttnParams.make(q, softmax_scale, causal, window_size, softcap, alibi_slopes, dropout_p, deterministic)
        engine = _engine_for(attn_type)
        seed = 0
        if p.dropout_p > 0:
            # drawn from torch's CPU generator: ranks that share torch.manual_seed() share the dropout seed, which
            # makes a sequence-parallel run reproduce the single-device masks exactly (ops/dropout.py)
            seed = int(dropout_seed) if dropout_seed is not None else int(torch.randint(1, 2**31 - 1, (1,)).item())
            p = replace(p, head_offset=int(head_offset))
        out, lse = ring_attn_forward(group, q, k, v, variant, p, engine, seed)
        ctx.save_for_backward(q, k, v, out, lse)
        ctx.p, ctx.variant, ctx.group, ctx.engine, ctx.seed = p, variant, group, engine, seed
        if return_softmax:
            ctx.mark_non_differentiable(lse)
            return out, lse, None
        return out

    @staticmethod
    def backward(ctx, dout, *args):
        q, k, v, out, lse = ctx.saved_tensors
        dq, dk, dv = ring_attn_backward(ctx.group, dout, q, k, v, out, lse, ctx.variant, ctx.p, ctx.engine, ctx.seed)
        return (dq, dk, dv) + (None,) * 14


def _processor_attention(attn_processor, group, q, k, v, softmax_scale, causal, return_attn_probs):
    """``AttnType.SPARSE_SAGE``: the user's sparse-attention module IS the block kernel (reference
    ``kernels/__init__.py:255-277``: ``attn_processor(q, k, v, is_causal=, scale=, tensor_layout="NHD")`` -> out, no
    LSE).  Without an LSE partial results cannot be merged, hence ring degree 1 only (reference guard
    ``hybrid/attn_layer.py:51-54``); Ulysses parallelism composes freely (heads are independent)."""
    if attn_processor is None or not callable(attn_processor):
        raise ImportError("AttnType.SPARSE_SAGE needs a sparse attention processor module passed as attn_processor")
    if group_size(group) > 1:
        raise RuntimeError("Sparse Sage attention does not support ring degree > 1.")
    out = attn_processor(q, k, v, is_causal=causal, scale=softmax_scale, tensor_layout="NHD")
    return (out, None, None) if return_attn_probs else out


def _make_funcs(variant: str):
    variant = canonical_variant(variant)

    def func(q, k, v, dropout_p=0.0, softmax_scale=None, causal=False, window_size=(-1, -1), softcap=0.0,
             alibi_slopes=None, deterministic=False, return_attn_probs=False, group=None, attn_type=None,
             attn_processor=None, head_offset=0, dropout_seed=None, backend=None):
        """``backend``: "auto" (default, ``LCA_B200_BACKEND``) runs the ring on the fused NVLink engine when the group
        is P2P-reachable on one node (push CTAs + tcgen05 attention in one kernel per rank), else -- and always with
        "collective" -- the NCCL/gloo P2P ring below."""
        if getattr(attn_type, "value", attn_type) == "sparse_sage":
            return _processor_attention(attn_processor, group, q, k, v, softmax_scale, causal, return_attn_probs)
        if head_offset == 0 and dropout_seed is None:
            from .fused import resolve_backend, try_fused
            res = try_fused("ring", group, resolve_backend(backend), attn_type, q, k, v, variant, dropout_p, softmax_scale,
                            causal, window_size, softcap, alibi_slopes, deterministic, None, return_attn_probs)
            if res is not None:
                return (res[0], res[1], None) if return_attn_probs else res
        return RingAttnFunc.apply(q, k, v, variant, dropout_p, softmax_scale, causal, window_size, softcap,
                                  alibi_slopes, deterministic, return_attn_probs, group, attn_type, attn_processor,
                                  head_offset, dropout_seed)

    def kvpacked(q, kv, dropout_p=0.0, softmax_scale=None, causal=False, window_size=(-1, -1), softcap=0.0,
                 alibi_slopes=None, deterministic=False, return_attn_probs=False, group=None, attn_type=None,
                 attn_processor=None, **kw):
        return func(q, kv[:, :, 0], kv[:, :, 1], dropout_p, softmax_scale, causal, window_size, softcap,
                    alibi_slopes, deterministic, return_attn_probs, group, attn_type, attn_processor, **kw)

    def qkvpacked(qkv, dropout_p=0.0, softmax_scale=None, causal=False, window_size=(-1, -1), softcap=0.0,
                  alibi_slopes=None, deterministic=False, return_attn_probs=False, group=None, attn_type=None,
                  attn_processor=None, **kw):
        return func(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], dropout_p, softmax_scale, causal, window_size,
                    softcap, alibi_slopes, deterministic, return_attn_probs, group, attn_type, attn_processor, **kw)

    return func, kvpacked, qkvpacked


ring_flash_attn_func, ring_flash_attn_kvpacked_func, ring_flash_attn_qkvpacked_func = _make_funcs("basic")
(zigzag_ring_flash_attn_func, zigzag_ring_flash_attn_kvpacked_func,
 zigzag_ring_flash_attn_qkvpacked_func) = _make_funcs("zigzag")
stripe_flash_attn_func, stripe_flash_attn_kvpacked_func, stripe_flash_attn_qkvpacked_func = _make_funcs("stripe")
