"""Plain-PyTorch block attention with a TRUE log-sum-exp and global-position masks.

This is (a) the CPU/gloo backend (BASELINE config 1), (b) the numerical oracle every CUDA
kernel is tested against, and (c) the backend behind ``AttnType.TORCH_*``.

The reference has no equivalent: its ``pytorch_attn_forward`` fabricates ``lse = zeros`` for
the math/cudnn modes (``yunchang/kernels/attention.py:108,130``), has no CPU dispatch for the
flash/efficient modes, and ``pytorch_attn_backward`` raises (``:138-159``).  The oracle in its
test-suite (``test/test_utils.py:1-130``) has the masks but no LSE and no block/offset form.

Conventions: q ``(B, Sq, H, D)``, k/v ``(B, Sk, Hkv, D)``, out ``(B, Sq, H, D)`` in q.dtype,
lse ``(B, H, Sq)`` fp32.  ``q_pos`` / ``k_pos`` are int64 vectors of *global* token positions;
all masks are functions of positions only, so any ring/zigzag/stripe block is exact.
Rows with no visible key produce ``out = 0`` and ``lse = -inf``.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

NEG_INF = float("-inf")


def _bias_and_mask(q_pos, k_pos, causal, window_size, alibi_slopes, H, device, q_grp=None, k_grp=None):
    """-> (mask (Sq,Sk) bool or None [True = masked], bias (H,Sq,Sk) fp32 or None)."""
    wl, wr = window_size
    mask = None
    rel = None
    if q_grp is not None and k_grp is not None:
        mask = q_grp.view(-1, 1) != k_grp.view(1, -1)
    if causal or wl >= 0 or wr >= 0 or alibi_slopes is not None:
        rel = k_pos.view(1, -1) - q_pos.view(-1, 1)  # (Sq, Sk), k - q
    if causal:
        m = rel > 0
        mask = m if mask is None else (mask | m)
    if wl >= 0:
        m = rel < -wl
        mask = m if mask is None else (mask | m)
    if wr >= 0 and not causal:
        m = rel > wr
        mask = m if mask is None else (mask | m)
    bias = None
    if alibi_slopes is not None:
        slopes = alibi_slopes.to(device=device, dtype=torch.float32)
        if slopes.dim() == 2:  # (B, H) -> handled by caller broadcasting; keep (B,H,1,1)
            bias = -slopes[:, :, None, None] * rel.abs().to(torch.float32)[None, None]
        else:
            bias = (-slopes[:, None, None] * rel.abs().to(torch.float32)[None])[None]
    return mask, bias


def _expand_kv(x: torch.Tensor, H: int) -> torch.Tensor:
    Hkv = x.shape[2]
    if Hkv == H:
        return x
    if H % Hkv:
        raise ValueError(f"query heads {H} not a multiple of kv heads {Hkv}")
    return x.repeat_interleave(H // Hkv, dim=2)


def _bounded_chunk(q_chunk: int, B: int, H: int, Sk: int, budget_elems: int = 1 << 28) -> int:
    """Query rows per step such that one (B, H, rows, Sk) fp32 temporary stays under ~1 GiB: the engine's memory is
    O(rows * Sk), never O(Sq * Sk) -- head dims the tcgen05 kernels do not take (> 128) stay usable at long context."""
    return max(16, min(q_chunk, budget_elems // max(1, B * H * Sk)))


def attn_block_fwd_ref(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    q_pos: torch.Tensor,
    k_pos: torch.Tensor,
    softmax_scale: float,
    causal: bool = False,
    window_size: Tuple[int, int] = (-1, -1),
    softcap: float = 0.0,
    alibi_slopes: Optional[torch.Tensor] = None,
    dropout_p: float = 0.0,
    dropout_mask: Optional[torch.Tensor] = None,
    q_chunk: int = 2048,
    q_grp: Optional[torch.Tensor] = None,
    k_grp: Optional[torch.Tensor] = None,
) -> Tuple[torch.Tensor, torch.Tensor]:
    B, Sq, H, D = q.shape
    Sk = k.shape[1]
    dev = q.device
    q_chunk = _bounded_chunk(q_chunk, B, H, Sk)
    kf = _expand_kv(k, H).to(torch.float32).permute(0, 2, 3, 1)  # (B,H,D,Sk)
    vf = _expand_kv(v, H).to(torch.float32).permute(0, 2, 1, 3)  # (B,H,Sk,D)
    out = torch.empty(B, Sq, H, D, dtype=q.dtype, device=dev)
    lse = torch.empty(B, H, Sq, dtype=torch.float32, device=dev)
    for s0 in range(0, Sq, q_chunk):
        s1 = min(Sq, s0 + q_chunk)
        qf = q[:, s0:s1].to(torch.float32).permute(0, 2, 1, 3)  # (B,H,sq,D)
        s = torch.matmul(qf, kf) * softmax_scale                # (B,H,sq,Sk)
        if softcap > 0:
            s = softcap * torch.tanh(s / softcap)
        mask, bias = _bias_and_mask(q_pos[s0:s1], k_pos, causal, window_size, alibi_slopes, H, dev,
                                    None if q_grp is None else q_grp[s0:s1], k_grp)
        if bias is not None:
            s = s + bias
        if mask is not None:
            s = s.masked_fill(mask[None, None], NEG_INF)
        l = torch.logsumexp(s, dim=-1)                          # (B,H,sq); -inf for empty rows
        p = torch.exp(s - torch.where(torch.isinf(l), torch.zeros_like(l), l)[..., None])
        if dropout_p > 0.0:
            if dropout_mask is None:
                raise ValueError("dropout needs an explicit keep-mask in the reference op")
            p = p * dropout_mask[:, :, s0:s1].to(p.dtype) / (1.0 - dropout_p)
        o = torch.matmul(p, vf)                                 # (B,H,sq,D)
        out[:, s0:s1] = o.permute(0, 2, 1, 3).to(q.dtype)
        lse[:, :, s0:s1] = l
    return out, lse


def attn_block_bwd_ref(
    dout: torch.Tensor,
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    out: torch.Tensor,
    lse: torch.Tensor,
    q_pos: torch.Tensor,
    k_pos: torch.Tensor,
    softmax_scale: float,
    causal: bool = False,
    window_size: Tuple[int, int] = (-1, -1),
    softcap: float = 0.0,
    alibi_slopes: Optional[torch.Tensor] = None,
    dropout_p: float = 0.0,
    dropout_mask: Optional[torch.Tensor] = None,
    delta: Optional[torch.Tensor] = None,
    q_chunk: int = 2048,
    q_grp: Optional[torch.Tensor] = None,
    k_grp: Optional[torch.Tensor] = None,
) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Block backward given the FINAL (merged) ``out``/``lse`` of the rows -> fp32 (dq, dk, dv).

    ``lse`` may be the global log-sum-exp over many KV blocks; the returned dk/dv are then this
    block's exact contribution and dq this block's partial sum (ring backward adds them up).
    """
    B, Sq, H, D = q.shape
    Sk, Hkv = k.shape[1], k.shape[2]
    dev = q.device
    g = H // Hkv
    q_chunk = _bounded_chunk(q_chunk, B, H, Sk)
    kf = _expand_kv(k, H).to(torch.float32).permute(0, 2, 1, 3)   # (B,H,Sk,D)
    vf = _expand_kv(v, H).to(torch.float32).permute(0, 2, 1, 3)
    dq = torch.zeros(B, Sq, H, D, dtype=torch.float32, device=dev)
    dk = torch.zeros(B, H, Sk, D, dtype=torch.float32, device=dev)
    dv = torch.zeros(B, H, Sk, D, dtype=torch.float32, device=dev)
    if delta is None:
        delta = (dout.to(torch.float32) * out.to(torch.float32)).sum(-1).permute(0, 2, 1)  # (B,H,Sq)
    for s0 in range(0, Sq, q_chunk):
        s1 = min(Sq, s0 + q_chunk)
        qf = q[:, s0:s1].to(torch.float32).permute(0, 2, 1, 3)      # (B,H,sq,D)
        dof = dout[:, s0:s1].to(torch.float32).permute(0, 2, 1, 3)
        raw = torch.matmul(qf, kf.transpose(-1, -2)) * softmax_scale
        if softcap > 0:
            t = torch.tanh(raw / softcap)
            s = softcap * t
        else:
            s = raw
        mask, bias = _bias_and_mask(q_pos[s0:s1], k_pos, causal, window_size, alibi_slopes, H, dev,
                                    None if q_grp is None else q_grp[s0:s1], k_grp)
        if bias is not None:
            s = s + bias
        if mask is not None:
            s = s.masked_fill(mask[None, None], NEG_INF)
        l = lse[:, :, s0:s1]
        l_safe = torch.where(torch.isinf(l), torch.zeros_like(l), l)
        p = torch.exp(s - l_safe[..., None])                        # masked -> exp(-inf) = 0
        p = torch.where(torch.isinf(l)[..., None], torch.zeros_like(p), p)
        if dropout_p > 0.0:
            keep = dropout_mask[:, :, s0:s1].to(p.dtype) / (1.0 - dropout_p)
            p_drop = p * keep
        else:
            keep, p_drop = None, p
        dv += torch.matmul(p_drop.transpose(-1, -2), dof)
        dp = torch.matmul(dof, vf.transpose(-1, -2))
        if keep is not None:
            dp = dp * keep
        ds = p * (dp - delta[:, :, s0:s1, None])
        if softcap > 0:
            ds = ds * (1.0 - t * t)
        ds = ds * softmax_scale
        dq[:, s0:s1] = torch.matmul(ds, kf).permute(0, 2, 1, 3)
        dk += torch.matmul(ds.transpose(-1, -2), qf)
    if g > 1:
        dk = dk.view(B, Hkv, g, Sk, D).sum(2)
        dv = dv.view(B, Hkv, g, Sk, D).sum(2)
    return dq, dk.permute(0, 2, 1, 3).contiguous(), dv.permute(0, 2, 1, 3).contiguous()


def attention_ref(
    q, k, v, causal=False, window_size=(-1, -1), softcap=0.0, alibi_slopes=None,
    softmax_scale=None, upcast=True,
):
    """Whole-sequence oracle (positions 0..S-1), same role as ``test/test_utils.py:attention_ref``.

    For ``Sq != Sk`` the causal diagonal is bottom-right aligned like flash-attn.
    """
    B, Sq, H, D = q.shape
    Sk = k.shape[1]
    if softmax_scale is None:
        softmax_scale = 1.0 / math.sqrt(D)
    q_pos = torch.arange(Sq, device=q.device) + (Sk - Sq)
    k_pos = torch.arange(Sk, device=q.device)
    if not upcast:
        q, k, v = (t.to(torch.float32).to(t.dtype) for t in (q, k, v))
    return attn_block_fwd_ref(q, k, v, q_pos, k_pos, softmax_scale, causal, window_size, softcap, alibi_slopes)


def construct_local_mask(seqlen_q, seqlen_k, window_size=(-1, -1), query_padding_mask=None, key_padding_mask=None,
                         device=None):
    """Boolean (Sq, Sk) mask, True = masked, for a sliding window with bottom-right aligned diagonal (role of
    ``test/test_utils.py:construct_local_mask``), expressed with the same position algebra as the kernels."""
    q_pos = torch.arange(seqlen_q, device=device) + (seqlen_k - seqlen_q)
    k_pos = torch.arange(seqlen_k, device=device)
    mask, _ = _bias_and_mask(q_pos, k_pos, False, window_size, None, 1, device)
    if mask is None:
        mask = torch.zeros(seqlen_q, seqlen_k, dtype=torch.bool, device=device)
    return mask
