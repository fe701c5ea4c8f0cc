"""Merge / varlen-LSE helpers with the reference's names (``yunchang/ring/utils.py:10-115``,
``ring/triton_utils.py``).  CUDA tensors go through the in-tree kernels (ops/csrc/util_kernels.cu)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from ..ops import native
from ..ops.attention import merge_out_lse_
from ..parallel.ring_comm import RingComm

__all__ = ["update_out_and_lse", "update_npu_out", "RingComm", "flatten_varlen_lse", "unflatten_varlen_lse"]


def update_out_and_lse(out: Optional[torch.Tensor], lse: Optional[torch.Tensor], block_out: torch.Tensor,
                       block_lse: torch.Tensor, slice_=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Reference-shaped API: ``out`` fp32 ``(B,S,H,D)``, ``lse`` ``(B,S,H,1)``; ``block_lse`` ``(B,H,S)``."""
    if out is None:
        if slice_ is not None:
            raise RuntimeError("first update_out_and_lse should not pass slice_ args")
        return block_out.to(torch.float32), block_lse.transpose(-2, -1).unsqueeze(-1).contiguous()
    if slice_ is not None:
        so, sl = out[slice_].contiguous(), lse[slice_].squeeze(-1).transpose(1, 2).contiguous()
        merge_out_lse_(so, sl, block_out, block_lse.contiguous())
        out[slice_], lse[slice_] = so, sl.transpose(1, 2).unsqueeze(-1)
        return out, lse
    acc_lse = lse.squeeze(-1).transpose(1, 2).contiguous()
    out = out.contiguous()
    merge_out_lse_(out, acc_lse, block_out, block_lse.contiguous())
    return out, acc_lse.transpose(1, 2).unsqueeze(-1).contiguous()


def flatten_varlen_lse(lse: torch.Tensor, cu_seqlens: torch.Tensor) -> torch.Tensor:
    """(B, H, max_s) -> (H, total)."""
    total = int(cu_seqlens[-1])
    if lse.is_cuda and native.available():
        return native.ext().flatten_varlen_lse(lse.contiguous(), cu_seqlens.to(torch.int32).contiguous(), total)
    parts = [lse[i, :, : int(cu_seqlens[i + 1] - cu_seqlens[i])] for i in range(len(cu_seqlens) - 1)]
    return torch.cat(parts, dim=1)


def unflatten_varlen_lse(lse: torch.Tensor, cu_seqlens: torch.Tensor, max_seqlen: int) -> torch.Tensor:
    """(H, total) -> (B, H, max_s), padded with -inf."""
    if lse.dim() == 3 and lse.shape[-1] == 1:     # reference layout (total, H, 1)
        lse = lse.squeeze(-1).transpose(0, 1)
    if lse.is_cuda and native.available():
        return native.ext().unflatten_varlen_lse(lse.contiguous(), cu_seqlens.to(torch.int32).contiguous(), int(max_seqlen))
    B = len(cu_seqlens) - 1
    out = torch.full((B, lse.shape[0], max_seqlen), float("-inf"), dtype=lse.dtype, device=lse.device)
    for i in range(B):
        s, e = int(cu_seqlens[i]), int(cu_seqlens[i + 1])
        out[i, :, : e - s] = lse[:, s:e]
    return out


def update_npu_out(cur_attn_out, cur_softmax_max, cur_softmax_sum, prev_attn_out, prev_softmax_max, prev_softmax_sum,
                   layout="BSND"):
    """Merge two partial results given as ``(out, row_max, row_sum)`` triples -- the statistic format of fused
    attention ops that return max/sum instead of an LSE (reference: ``ring/utils.py:54-93``, written for Ascend's
    ``npu_fusion_attention``).  ``out`` is ``(B, S, N, D)``; ``max``/``sum`` are ``(B, N, S, k)`` with the value
    replicated along the last axis.  Device-agnostic; equivalent to the LSE merge with ``lse = max + log(sum)``."""
    assert layout == "BSND", "only the BSND layout is supported"
    if prev_attn_out is None:
        return cur_attn_out, cur_softmax_max, cur_softmax_sum
    new_max = torch.maximum(prev_softmax_max, cur_softmax_max)
    prev_scaled = prev_softmax_sum * torch.exp(prev_softmax_max - new_max)
    cur_scaled = cur_softmax_sum * torch.exp(cur_softmax_max - new_max)
    new_sum = prev_scaled + cur_scaled
    w_prev = (prev_scaled / new_sum)[..., 0].transpose(1, 2).unsqueeze(-1)      # (B, S, N, 1)
    w_cur = (cur_scaled / new_sum)[..., 0].transpose(1, 2).unsqueeze(-1)
    out = (prev_attn_out.to(torch.float32) * w_prev + cur_attn_out.to(torch.float32) * w_cur).to(prev_attn_out.dtype)
    return out, new_max, new_sum
