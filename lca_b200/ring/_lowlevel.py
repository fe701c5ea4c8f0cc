"""Reference-shaped low-level ring entry points.

The reference exposes, next to the autograd functions, a ``<variant>_forward(process_group, q, k, v, softmax_scale,
...) -> (out, lse)`` / ``<variant>_backward(process_group, dout, q, k, v, out, softmax_lse, softmax_scale, ...) ->
(dq, dk, dv)`` pair per ring flavour (``ring_flash_attn.py:7-147``, ``zigzag_ring_flash_attn.py``,
``stripe_flash_attn.py``, the two varlen files).  Third-party code imports them to build its own autograd nodes.
Here they are thin adapters onto the single position-aware ring loop (``parallel/ring_attention.py``); dropout
masks are derived from ``dropout_seed`` (default 0) so a forward/backward pair built from these functions sees the
same mask, which the reference cannot guarantee.
"""
from __future__ import annotations

from ..ops.attention import AttnParams
from ..parallel.ring_attention import _engine_for, ring_attn_backward, ring_attn_forward


def _params(q4, softmax_scale, causal, window_size, softcap, alibi_slopes, dropout_p, deterministic):
    return AttnParams.make(q4, softmax_scale, causal, tuple(window_size), softcap, alibi_slopes, dropout_p,
                           deterministic)


def make_dense(variant: str):
    def forward(process_group, q, k, v, softmax_scale, dropout_p=0, causal=True, window_size=(-1, -1), softcap=0.0,
                alibi_slopes=None, deterministic=False, attn_type=None, attn_processor=None, dropout_seed=0):
        p = _params(q, softmax_scale, causal, window_size, softcap, alibi_slopes, dropout_p, deterministic)
        return ring_attn_forward(process_group, q, k, v, variant, p, _engine_for(attn_type), dropout_seed)

    def backward(process_group, dout, q, k, v, out, softmax_lse, softmax_scale, dropout_p=0, causal=True,
                 window_size=(-1, -1), softcap=0.0, alibi_slopes=None, deterministic=False, attn_type=None,
                 dropout_seed=0):
        p = _params(q, softmax_scale, causal, window_size, softcap, alibi_slopes, dropout_p, deterministic)
        return ring_attn_backward(process_group, dout, q, k, v, out, softmax_lse, variant, p, _engine_for(attn_type),
                                  dropout_seed)

    forward.__name__ = backward.__name__ = f"{variant}_ring"
    return forward, backward


def _cu_list(cu_seqlens):
    return [int(x) for x in (cu_seqlens.tolist() if hasattr(cu_seqlens, "tolist") else cu_seqlens)]


def make_varlen(variant: str, with_half_index: bool):
    """Varlen flavours take packed ``(total_local, H, D)`` shards; LSE is ``(H, total_local)``.  The zigzag variant of
    the reference threads two precomputed index tensors through (``half_index0/1``); they are accepted and ignored --
    positions make them unnecessary."""

    def parse(q, rest, kw):
        """Positional tail / keywords of the reference's low-level signature -> (q as a batch of one, AttnParams, args)."""
        if with_half_index:
            rest = rest[2:] if len(rest) >= 2 else rest
            kw.pop("half_index0", None), kw.pop("half_index1", None)
        names = ["softmax_scale", "dropout_p", "causal", "window_size", "softcap", "alibi_slopes", "deterministic"]
        a = dict(dropout_p=0, causal=True, window_size=(-1, -1), softcap=0.0, alibi_slopes=None, deterministic=False)
        a.update(dict(zip(names, rest)))
        a.update(kw)
        q4 = q.unsqueeze(0)
        p = _params(q4, a["softmax_scale"], a["causal"], a["window_size"], a["softcap"], a["alibi_slopes"],
                    a["dropout_p"], a["deterministic"])
        return q4, p, a

    def forward(process_group, q, k, v, cu_seqlens, max_seqlen, *rest, **kw):
        q4, p, a = parse(q, rest, kw)
        if p.dropout_p > 0:
            raise NotImplementedError("dropout is not supported on the varlen ring path")
        cu = _cu_list(cu_seqlens)
        out, lse = ring_attn_forward(process_group, q4, k.unsqueeze(0), v.unsqueeze(0), variant, p,
                                     _engine_for(a.get("attn_type")), 0, cu, cu)
        return out.squeeze(0), lse.squeeze(0)

    def backward(process_group, dout, q, k, v, out, softmax_lse, cu_seqlens, max_seqlen, *rest, **kw):
        q4, p, a = parse(q, rest, kw)
        cu = _cu_list(cu_seqlens)
        dq, dk, dv = ring_attn_backward(process_group, dout.unsqueeze(0), q4, k.unsqueeze(0), v.unsqueeze(0),
                                        out.unsqueeze(0), softmax_lse.unsqueeze(0), variant, p,
                                        _engine_for(a.get("attn_type")), 0, cu, cu)
        return dq.squeeze(0), dk.squeeze(0), dv.squeeze(0)

    return forward, backward
