"""Block-scaled FP8 (e4m3) attention forward.

Role: the reference reaches fp8 only through third-party forward-only kernels (``AttnType.SAGE_FP8*``,
``kernels/__init__.py:177-254``; FA3 with caller-supplied fp8 tensors, ``kernels/attention.py:258-292``), none of
which run on sm_100.  Here: Q and K are quantised to e4m3 with one fp32 scale per (batch, head, 128-row block), V
with one scale per (batch, kv head); the tcgen05 ``kind::f8f6f4`` kernel (``csrc/fmha_fwd_fp8_sm100.cu``) folds the
Q/K block scales into the softmax argument and the V scale into the final normalisation.

Status: validated on B200 in round 2 (``tests/test_fp8.py`` against :func:`attn_fp8_emulated`, the bit-faithful
PyTorch model of the kernel's arithmetic: quantise -> fp32 matmuls -> e4m3 P; timings and errors in
``profiles/r2/``).  Selected by ``AttnType.SAGE_FP8 / SAGE_FP8_SM90 / SAGE_AUTO`` (the reference's quantised
forward-only family) through the ``"fp8"`` engine of :func:`lca_b200.ops.attention.attn_block_fwd`; ``LCA_B200_FP8=0``
sends those types to the bf16 kernel instead.  Forward only, like the reference's fp8 paths: under autograd the
backward runs in 16 bit on the saved 16-bit operands.
"""
from __future__ import annotations

import os
from typing import Tuple

import torch

from ..parallel.layout import PosSpec, pos_tensor
from . import native
from .attention import AttnParams
from .ref_attention import _bias_and_mask, _expand_kv

E4M3_MAX = 448.0
BLOCK = 128


def enabled() -> bool:
    """The CUDA fp8 kernel may be used (``LCA_B200_FP8=0`` / legacy ``LCA_B200_EXPERIMENTAL_FP8=0`` turn it off)."""
    return os.environ.get("LCA_B200_FP8", os.environ.get("LCA_B200_EXPERIMENTAL_FP8", "1")) == "1"


def takes(q: torch.Tensor, k: torch.Tensor) -> bool:
    """Inputs the fp8 forward handles: the CUDA kernel needs head_dim 128 and 16-bit inputs; the emulation (CPU, other
    head dims) is O(Sq*Sk) memory and only meant for tests, so it is limited to small blocks."""
    if q.is_cuda and native.available():
        return enabled() and q.shape[-1] == 128 and q.dtype in (torch.bfloat16, torch.float16)
    return enabled() and q.shape[1] * k.shape[1] <= 4096 * 4096


def quantize_blockwise(x: torch.Tensor, per_head: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """x (B,S,H,D) -> (x8 float8_e4m3fn (B,S,H,D), scale): per-(b,h,128-row block) ``(B,H,nblk)`` or per-head ``(B,H)``."""
    if x.is_cuda and native.available() and x.dtype in (torch.bfloat16, torch.float16):
        y, s = native.ext().quantize_e4m3(x if x.stride(-1) == 1 else x.contiguous(), per_head)
        return y.view(torch.float8_e4m3fn), s
    B, S, H, D = x.shape
    xf = x.to(torch.float32)
    if per_head:
        scale = (xf.abs().amax(dim=(1, 3)) / E4M3_MAX).clamp_min(1e-12)                     # (B,H)
        y = (xf / scale[:, None, :, None]).clamp(-E4M3_MAX, E4M3_MAX).to(torch.float8_e4m3fn)
        return y, scale
    nblk = (S + BLOCK - 1) // BLOCK
    pad = nblk * BLOCK - S
    xp = torch.nn.functional.pad(xf, (0, 0, 0, 0, 0, pad)) if pad else xf
    amax = xp.view(B, nblk, BLOCK, H, D).abs().amax(dim=(2, 4))                              # (B,nblk,H)
    scale = torch.where(amax > 0, amax / E4M3_MAX, torch.ones_like(amax)).permute(0, 2, 1).contiguous()   # (B,H,nblk)
    rows = scale.repeat_interleave(BLOCK, dim=2)[:, :, :S].permute(0, 2, 1)                  # (B,S,H)
    y = (xf / rows[..., None]).clamp(-E4M3_MAX, E4M3_MAX).to(torch.float8_e4m3fn)
    return y, scale


def attn_fp8_emulated(q, k, v, q_pos: PosSpec, k_pos: PosSpec, p: AttnParams):
    """PyTorch model of the fp8 kernel: e4m3 Q/K/V with block scales, fp32 accumulation, P rounded to e4m3."""
    dev = q.device
    B, Sq, H, D = q.shape
    q8, sq = quantize_blockwise(q)
    k8, sk = quantize_blockwise(k)
    v8, sv = quantize_blockwise(v, per_head=True)
    Hkv = k.shape[2]
    g = H // Hkv
    qf = q8.to(torch.float32).permute(0, 2, 1, 3)                                         # (B,H,Sq,D)
    kf = _expand_kv(k8.to(torch.float32), H).permute(0, 2, 3, 1)                          # (B,H,D,Sk)
    vf = _expand_kv(v8.to(torch.float32), H).permute(0, 2, 1, 3)
    sq_r = sq.repeat_interleave(BLOCK, dim=2)[:, :, :Sq]                                   # (B,H,Sq)
    sk_r = sk.repeat_interleave(g, dim=1).repeat_interleave(BLOCK, dim=2)[:, :, : k.shape[1]]
    s = torch.matmul(qf, kf) * sq_r[..., None] * sk_r[:, :, None, :] * p.softmax_scale
    if p.softcap > 0:
        s = p.softcap * torch.tanh(s / p.softcap)
    mask, bias = _bias_and_mask(pos_tensor(q_pos, dev), pos_tensor(k_pos, dev), p.causal, p.window_size, p.alibi_slopes, H, dev)
    if bias is not None:
        s = s + bias
    if mask is not None:
        s = s.masked_fill(mask[None, None], float("-inf"))
    lse = torch.logsumexp(s, dim=-1)
    m = s.amax(dim=-1, keepdim=True)
    m = torch.where(torch.isinf(m), torch.zeros_like(m), m)
    pe = torch.exp(s - m)                                                                  # in [0, 1]
    l = pe.sum(-1, keepdim=True)
    p8 = pe.to(torch.float8_e4m3fn).to(torch.float32)                                      # the kernel feeds e4m3 P to the PV MMA
    o = torch.matmul(p8, vf) * sv.repeat_interleave(g, dim=1)[:, :, None, None] / l.clamp_min(1e-30)
    o = torch.where(l > 0, o, torch.zeros_like(o))
    return o.permute(0, 2, 1, 3).to(torch.bfloat16 if q.dtype == torch.float32 else q.dtype), lse


def attn_fp8_fwd(q, k, v, q_pos: PosSpec, k_pos: PosSpec, p: AttnParams):
    """Block-scaled fp8 forward of one block -> (out bf16/fp16-as-input, lse).  CUDA kernel when enabled and
    supported, otherwise the PyTorch emulation."""
    use_cuda = (enabled() and q.is_cuda and native.available() and q.shape[-1] == 128
                and q.dtype in (torch.bfloat16, torch.float16) and p.dropout_p == 0.0)
    if not use_cuda:
        return attn_fp8_emulated(q, k, v, q_pos, k_pos, p)
    C = native.ext()
    q8, sq = quantize_blockwise(q)
    k8, sk = quantize_blockwise(k)
    v8, sv = quantize_blockwise(v, per_head=True)
    B, Sq, H, D = q.shape
    out = torch.empty((B, Sq, H, D), dtype=torch.bfloat16, device=q.device)
    lse = torch.empty((B, H, Sq), dtype=torch.float32, device=q.device)
    qrows = sorted(native._rows(q_pos), key=lambda r: -r[2])
    qsegs = [[r0, n, pos0, -1, r0, 0, 0, g] for (r0, n, pos0, g) in qrows]
    ksegs = [[r0, n, pos0, -1, g] for (r0, n, pos0, g) in native._rows(k_pos)]
    wl, wr = native.window_bounds(p)
    alibi = p.alibi_slopes
    if alibi is not None:
        alibi = alibi.to(device=q.device, dtype=torch.float32).contiguous()
    C.fmha_fwd_fp8(q8.view(torch.uint8), k8.view(torch.uint8), v8.view(torch.uint8), sq, sk, sv, qsegs, ksegs,
                   native._common_stride(q_pos), native._common_stride(k_pos), out, lse, float(p.softmax_scale), wl, wr,
                   float(p.softcap), alibi)
    return out.to(q.dtype), lse
