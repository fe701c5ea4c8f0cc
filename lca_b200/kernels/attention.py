"""Uniform forward / backward attention wrappers (contract of ``yunchang/kernels/attention.py:165-250``).

forward : ``(q,k,v,dropout_p,softmax_scale,causal,window_size,softcap,alibi_slopes,return_softmax)
          -> (out, lse)``
backward: ``(dout,q,k,v,out,softmax_lse,dq,dk,dv,dropout_p,softmax_scale,causal,window_size,softcap,
          alibi_slopes,deterministic,rng_state)`` writing into ``dq/dk/dv`` in place.

Unlike the reference's torch wrappers, the pure-PyTorch ones here return a real LSE and have a
working backward (reference: fake zeros LSE ``:108,130`` and ``pytorch_attn_backward`` raises ``:138-159``).
These whole-sequence wrappers assume positions ``0..S-1`` (bottom-right aligned causal when
``Sq != Sk``); ring code uses the position-aware block ops directly.
"""
from __future__ import annotations

from dataclasses import replace

import torch

from ..ops.attention import AttnParams, attn_block_bwd, attn_block_fwd
from ..parallel.layout import Seg


def _pos(q, k):
    Sq, Sk = q.shape[1], k.shape[1]
    return (Seg(max(Sk - Sq, 0), Sq, 1),), (Seg(0, Sk, 1),)


def _with_dropout(p: AttnParams, dropout_seed, head_offset) -> AttnParams:
    """Dropout masks are functions of (seed, batch, head, positions) (``ops/dropout.py``): a forward / backward pair
    must be given the same ``dropout_seed`` (default 0) -- there is no RNG state to carry."""
    if p.dropout_p <= 0.0:
        return p
    return replace(p, dropout_seed=int(dropout_seed or 0), head_offset=int(head_offset or 0))


def _fwd(engine, q, k, v, dropout_p=0.0, softmax_scale=None, causal=False, window_size=(-1, -1), softcap=0.0,
         alibi_slopes=None, return_softmax=False, op_type=None, dropout_seed=0, head_offset=0):
    p = _with_dropout(AttnParams.make(q, softmax_scale, causal, window_size, softcap, alibi_slopes, dropout_p),
                      dropout_seed, head_offset)
    qp, kp = _pos(q, k)
    return attn_block_fwd(q, k, v, qp, kp, p, engine)


def _bwd(engine, dout, q, k, v, out, softmax_lse, dq=None, dk=None, dv=None, dropout_p=0.0, softmax_scale=None,
         causal=False, window_size=(-1, -1), softcap=0.0, alibi_slopes=None, deterministic=False, rng_state=None,
         dropout_seed=None, head_offset=0):
    if dropout_seed is None:          # the reference threads flash-attn's rng_state here; an int seed is accepted too
        dropout_seed = rng_state if isinstance(rng_state, int) else 0
    p = _with_dropout(AttnParams.make(q, softmax_scale, causal, window_size, softcap, alibi_slopes, dropout_p,
                                      deterministic), dropout_seed, head_offset)
    qp, kp = _pos(q, k)
    gq, gk, gv = attn_block_bwd(dout, q, k, v, out, softmax_lse, qp, kp, p, engine)
    if dq is not None:
        dq.copy_(gq); dk.copy_(gk); dv.copy_(gv)
        return dq, dk, dv
    return gq.to(q.dtype), gk.to(k.dtype), gv.to(v.dtype)


def flash_attn_forward(q, k, v, *a, **kw):
    return _fwd(None, q, k, v, *a, **kw)


def flash_attn_backward(dout, q, k, v, out, softmax_lse, *a, **kw):
    return _bwd(None, dout, q, k, v, out, softmax_lse, *a, **kw)


def pytorch_attn_forward(q, k, v, *a, **kw):
    return _fwd("torch", q, k, v, *a, **kw)


def pytorch_attn_backward(dout, q, k, v, out, softmax_lse, *a, **kw):
    return _bwd("torch", dout, q, k, v, out, softmax_lse, *a, **kw)


class _LocalAttnFunc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, p, engine, return_lse):
        qp, kp = _pos(q, k)
        out, lse = attn_block_fwd(q, k, v, qp, kp, p, engine)
        ctx.save_for_backward(q, k, v, out, lse)
        ctx.p, ctx.engine = p, engine
        if return_lse:
            ctx.mark_non_differentiable(lse)
            return out, lse
        return out

    @staticmethod
    def backward(ctx, dout, *args):
        q, k, v, out, lse = ctx.saved_tensors
        qp, kp = _pos(q, k)
        dq, dk, dv = attn_block_bwd(dout, q, k, v, out, lse, qp, kp, ctx.p, ctx.engine)
        return dq.to(q.dtype), dk.to(k.dtype), dv.to(v.dtype), None, None, None


def _func(engine, q, k, v, dropout_p=0.0, softmax_scale=None, causal=False, window_size=(-1, -1), softcap=0.0,
          alibi_slopes=None, deterministic=False, return_attn_probs=False, dropout_seed=None, head_offset=0):
    p = AttnParams.make(q, softmax_scale, causal, window_size, softcap, alibi_slopes, dropout_p, deterministic)
    if p.dropout_p > 0.0:
        if dropout_seed is None:      # torch's CPU generator: torch.manual_seed() makes the masks reproducible
            dropout_seed = int(torch.randint(1, 2**31 - 1, (1,)).item())
        p = _with_dropout(p, dropout_seed, head_offset)
    if return_attn_probs:
        out, lse = _LocalAttnFunc.apply(q, k, v, p, engine, True)
        return out, lse, None
    return _LocalAttnFunc.apply(q, k, v, p, engine, False)


def flash_attn_func(q, k, v, *a, **kw):
    """Autograd-aware single-device attention (role of ``flash_attn.flash_attn_func``)."""
    return _func(None, q, k, v, *a, **kw)


def pytorch_attn_func(q, k, v, *a, **kw):
    return _func("torch", q, k, v, *a, **kw)


# ------------------------------------------------------------------------------------------------
# Names of the reference's other per-library wrappers (``kernels/attention.py:253-457``).  On B200 the
# FA3 / flashinfer entry points run the native tcgen05 engine; the ROCm / Ascend ones raise.
# ------------------------------------------------------------------------------------------------
def flash_attn3_func_forward(q, k, v, dropout_p=0.0, softmax_scale=None, causal=False, window_size=(-1, -1), softcap=0.0,
                             alibi_slopes=None, return_softmax=False):
    """Unlike the reference's FA3 wrapper (hard-codes ``causal=False`` and ``softcap=0``, ``:283-286``) every
    argument is honoured."""
    return _fwd(None, q, k, v, dropout_p, softmax_scale, causal, window_size, softcap, alibi_slopes, return_softmax)


def flash_attn3_func_backward(dout, q, k, v, out, softmax_lse, *a, **kw):
    return _bwd(None, dout, q, k, v, out, softmax_lse, *a, **kw)


def flashinfer_attn_forward(q, k, v, *a, **kw):
    """Returns a natural-log LSE directly (the reference converts flashinfer's base-2 LSE, ``:393``)."""
    return _fwd(None, q, k, v, *a, **kw)


def flashinfer_attn_backbward(dout, q, k, v, out, softmax_lse, *a, **kw):   # (sic) reference spelling
    return _bwd(None, dout, q, k, v, out, softmax_lse, *a, **kw)


flashinfer_attn_backward = flashinfer_attn_backbward


def _foreign(name, hw):
    def fn(*a, **kw):
        raise RuntimeError(f"{name} targets {hw}; not available in the B200 build (use flash_attn_forward/backward)")
    fn.__name__ = name
    return fn


flash_attn_forward_aiter = _foreign("flash_attn_forward_aiter", "AMD ROCm (aiter)")
npu_fused_attn_forward = _foreign("npu_fused_attn_forward", "Ascend NPU")
npu_fused_attn_backward = _foreign("npu_fused_attn_backward", "Ascend NPU")
