"""Block-attention dispatch: one op signature, two engines.

``attn_block_fwd`` / ``attn_block_bwd`` are what every parallel algorithm in this package
calls (ring steps, Ulysses local attention, varlen).  Engines:

* ``"native"`` -- the in-tree sm_100a tcgen05/TMEM/TMA kernels (``ops/csrc/fmha_*.cu``);
* ``"torch"``  -- :mod:`lca_b200.ops.ref_attention` (CPU/gloo, oracle, AttnType.TORCH_*).

This replaces the reference's per-library wrappers with their uniform fwd/bwd contracts
(``yunchang/kernels/attention.py:165-250``).  The contract here differs on purpose: blocks
carry *global positions* (:class:`~lca_b200.parallel.layout.Seg` lists) so causal, sliding
window and ALiBi are exact across ring blocks (the reference applies them per block with
local offsets, ``ring_flash_attn.py:44``, which is only right for ``window=(-1,-1)``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, replace
from typing import Optional, Tuple

import torch

from ..parallel.layout import PosSpec, group_tensor, has_groups, pos_min_max, pos_tensor
from . import dropout, native, ref_attention


@dataclass(frozen=True)
class AttnParams:
    softmax_scale: float
    causal: bool = False
    window_size: Tuple[int, int] = (-1, -1)
    softcap: float = 0.0
    alibi_slopes: Optional[torch.Tensor] = None
    dropout_p: float = 0.0
    deterministic: bool = False
    dropout_seed: int = 0          # dropout is a pure function of (seed, batch, head, q position, k position): ops/dropout.py
    head_offset: int = 0           # global index of local head 0 (Ulysses head shards), part of the dropout key

    @staticmethod
    def make(q, softmax_scale=None, causal=False, window_size=(-1, -1), softcap=0.0,
             alibi_slopes=None, dropout_p=0.0, deterministic=False) -> "AttnParams":
        if softmax_scale is None:
            softmax_scale = 1.0 / math.sqrt(q.shape[-1])
        ws = (-1, -1) if window_size is None else (int(window_size[0]), int(window_size[1]))
        return AttnParams(float(softmax_scale), bool(causal), ws, float(softcap or 0.0),
                          alibi_slopes, float(dropout_p or 0.0), bool(deterministic))


def block_is_visible(q_pos: PosSpec, k_pos: PosSpec, p: AttnParams) -> bool:
    """False when every (q, k) pair of the block is masked -> the whole ring step is skipped.

    This is the position-space generalisation of the reference's ``step <= rank`` test
    (``ring_flash_attn.py:35``).
    """
    wl, wr = p.window_size
    for g in {s.group for s in q_pos}:
        qs = tuple(s for s in q_pos if s.group == g and s.count > 0)
        ks = tuple(s for s in k_pos if s.group == g and s.count > 0)
        if not qs or not ks:
            continue
        qlo, qhi = pos_min_max(qs)
        klo, khi = pos_min_max(ks)
        if p.causal and klo > qhi:
            continue
        if wl >= 0 and qlo - khi > wl:
            continue
        if wr >= 0 and not p.causal and klo - qhi > wr:
            continue
        return True
    return False


def pick_engine(q: torch.Tensor, engine: Optional[str]) -> str:
    if engine == "fp8":      # quantised forward (ops/fp8.py); everything it cannot take runs on the 16-bit engines
        engine = None
    if engine in ("torch", "native"):
        if engine == "native" and not native.supports(q):
            raise RuntimeError(
                "engine='native' requested but the sm_100a extension cannot run this input "
                f"(device={q.device}, dtype={q.dtype}, head_dim={q.shape[-1]}): {native.why_not(q)}"
            )
        return engine
    if q.is_cuda:
        if native.supports(q):
            return "native"
        if native.must_be_native() and not native.extension_loaded():
            # a Blackwell box without the extension is a build problem, not something to paper over
            raise RuntimeError(f"native sm_100a kernels required but unusable: {native.why_not(q)}")
        _warn_once(f"lca_b200: using the PyTorch attention engine on CUDA ({native.why_not(q)})")
    return "torch"


_WARNED = set()


def _warn_once(msg: str) -> None:
    if msg not in _WARNED:
        _WARNED.add(msg)
        import warnings
        warnings.warn(msg, stacklevel=3)


def _dropout_args(p: AttnParams, B, H, qpt, kpt, q_grp, dropout_mask):
    """-> (effective drop probability, keep mask) for the PyTorch engine; the mask is regenerated from the global
    coordinates unless the caller supplies one (oracle tests)."""
    if p.dropout_p <= 0.0:
        return 0.0, None
    if dropout_mask is None:
        dropout_mask = dropout.keep_mask(p.dropout_seed, B, H, qpt, kpt, p.dropout_p, p.head_offset, 0, q_grp)
    return dropout.p_eff(p.dropout_p), dropout_mask


def attn_block_fwd(q, k, v, q_pos: PosSpec, k_pos: PosSpec, p: AttnParams,
                   engine: Optional[str] = None, dropout_mask=None):
    """-> out (B,Sq,H,D) q.dtype, lse (B,H,Sq) fp32.  ``engine="fp8"``: block-scaled e4m3 forward (``ops/fp8.py``:
    tcgen05 ``kind::f8f6f4`` kernel on B200, bit-faithful PyTorch emulation elsewhere); the backward of such a
    forward runs on the 16-bit engine against the saved 16-bit q/k/v (quantisation error is treated as noise, as in
    FP8 training recipes that keep the backward of attention in bf16)."""
    if engine == "fp8" and p.dropout_p == 0.0 and dropout_mask is None:
        from . import fp8
        if fp8.takes(q, k):
            return fp8.attn_fp8_fwd(q, k, v, q_pos, k_pos, p)
    eng = pick_engine(q, engine)
    if eng == "native" and (p.dropout_p == 0.0 or (dropout_mask is None and native.dropout_supported(p))):
        return native.fmha_fwd(q, k, v, q_pos, k_pos, p)
    grp = has_groups(q_pos) or has_groups(k_pos)
    qpt, kpt = pos_tensor(q_pos, q.device), pos_tensor(k_pos, q.device)
    q_grp = group_tensor(q_pos, q.device) if grp else None
    pe, dm = _dropout_args(p, q.shape[0], q.shape[2], qpt, kpt, q_grp, dropout_mask)
    return ref_attention.attn_block_fwd_ref(
        q, k, v, qpt, kpt, p.softmax_scale, p.causal, p.window_size, p.softcap, p.alibi_slopes, pe, dm,
        q_grp=q_grp, k_grp=group_tensor(k_pos, q.device) if grp else None)


def attn_block_bwd(dout, q, k, v, out, lse, q_pos: PosSpec, k_pos: PosSpec, p: AttnParams,
                   engine: Optional[str] = None, dropout_mask=None, delta=None, lse2=None, into=None):
    """-> (dq, dk, dv) contributions of this block (see ref_attention.attn_block_bwd_ref).

    ``into = (dq, dk, dv, acc_dq, acc_dkv)``: fp32 buffers to write (``acc_*`` False) or accumulate into
    (``acc_*`` True); the native kernels do the ``+=`` in their epilogue, so a ring step costs no
    extra pass over the gradients."""
    eng = pick_engine(q, engine)
    if eng == "native" and native.has_bwd() and (p.dropout_p == 0.0 or
                                                  (dropout_mask is None and native.dropout_supported(p))):
        if into is not None:
            dq, dk, dv, acc_dq, acc_dkv = into
            return native.fmha_bwd(dout, q, k, v, out, lse, q_pos, k_pos, p, delta=delta, lse2=lse2, dq=dq, dk=dk,
                                   dv=dv, acc_dq=acc_dq, acc_dkv=acc_dkv)
        return native.fmha_bwd(dout, q, k, v, out, lse, q_pos, k_pos, p, delta=delta, lse2=lse2)
    if into is not None:
        dq, dk, dv, acc_dq, acc_dkv = into
        gq, gk, gv = attn_block_bwd(dout, q, k, v, out, lse, q_pos, k_pos, p, eng, dropout_mask, delta, lse2)
        dq.add_(gq) if acc_dq else dq.copy_(gq)
        dk.add_(gk) if acc_dkv else dk.copy_(gk)
        dv.add_(gv) if acc_dkv else dv.copy_(gv)
        return dq, dk, dv
    grp = has_groups(q_pos) or has_groups(k_pos)
    qpt, kpt = pos_tensor(q_pos, q.device), pos_tensor(k_pos, q.device)
    q_grp = group_tensor(q_pos, q.device) if grp else None
    pe, dm = _dropout_args(p, q.shape[0], q.shape[2], qpt, kpt, q_grp, dropout_mask)
    return ref_attention.attn_block_bwd_ref(
        dout, q, k, v, out, lse, qpt, kpt, p.softmax_scale, p.causal, p.window_size, p.softcap, p.alibi_slopes, pe,
        dm, delta=delta, q_grp=q_grp, k_grp=group_tensor(k_pos, q.device) if grp else None)


# ------------------------------------------------------------------------------------------
# online-softmax merge of partial results
# ------------------------------------------------------------------------------------------
def merge_out_lse_(out_acc: torch.Tensor, lse_acc: torch.Tensor, block_out: torch.Tensor,
                   block_lse: torch.Tensor) -> None:
    """In-place ``(out_acc, lse_acc) <- merge((out_acc, lse_acc), (block_out, block_lse))``.

    out_acc ``(B,S,H,D)`` fp32, lse_acc ``(B,H,S)`` fp32; block_* same shapes (block_out any
    float dtype).  Exact log-sum-exp merge (reference: ``yunchang/ring/utils.py:10-51``), made
    safe for ``-inf`` rows (no visible keys so far), which the reference's sigmoid form turns
    into NaN.
    """
    if out_acc.is_cuda and native.available():
        native.merge_out_lse_(out_acc, lse_acc, block_out, block_lse)
        return
    new = torch.logaddexp(lse_acc, block_lse)
    safe = torch.where(torch.isinf(new) & (new < 0), torch.zeros_like(new), new)
    w_old = torch.exp(lse_acc - safe).transpose(1, 2).unsqueeze(-1)    # (B,S,H,1)
    w_new = torch.exp(block_lse - safe).transpose(1, 2).unsqueeze(-1)
    out_acc.mul_(w_old).add_(block_out.to(torch.float32) * w_new)
    lse_acc.copy_(new)


def with_scale(p: AttnParams, q) -> AttnParams:
    return p if p.softmax_scale is not None else replace(p, softmax_scale=1.0 / math.sqrt(q.shape[-1]))
