"""A small decoder-only transformer whose attention is sequence-parallel through this library.

Role: the integration contract the reference documents with its Megatron-DeepSpeed patch
(``patches/Megatron-DeepSpeed.patch:1-327``: call ``set_seq_parallel_pg`` when model-parallel groups are
built, swap the attention core for ``LongContextAttention``) and its convergence check
(``README.md:157-162``: loss curves of DP vs Ulysses2 x Ring2 overlap).  Everything that is NOT attention
is token-local, so sequence parallelism only needs (a) tokens/labels sharded with
``EXTRACT_FUNC_DICT[ring_impl_type]`` and (b) RoPE position ids permuted the same way
(:func:`lca_b200.parallel.layout.local_token_index`).  Gradients of replicated weights are averaged
over the sequence-parallel group by :func:`allreduce_sp_grads` (the reference leaves this to Megatron).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from ..globals import PROCESS_GROUP
from ..hybrid import LongContextAttention
from ..kernels import AttnType
from ..parallel.layout import local_token_index


@dataclass
class SPTransformerConfig:
    vocab_size: int = 256
    dim: int = 64
    n_layers: int = 2
    n_heads: int = 4
    n_kv_heads: int = 2
    ffn_mult: int = 2
    rope_theta: float = 10000.0
    ring_impl_type: str = "zigzag"
    attn_type: AttnType = AttnType.FA
    backend: Optional[str] = None


def rope(x: torch.Tensor, pos: torch.Tensor, theta: float) -> torch.Tensor:
    """x (B, S, H, D), pos (S,) global positions."""
    D = x.shape[-1]
    inv = 1.0 / (theta ** (torch.arange(0, D, 2, device=x.device, dtype=torch.float32) / D))
    ang = pos.to(torch.float32)[:, None] * inv[None]                       # (S, D/2)
    cos, sin = ang.cos()[None, :, None, :], ang.sin()[None, :, None, :]
    x1, x2 = x.float()[..., 0::2], x.float()[..., 1::2]
    out = torch.stack([x1 * cos - x2 * sin, x1 * sin + x2 * cos], dim=-1).flatten(-2)
    return out.to(x.dtype)


class SPBlock(nn.Module):
    def __init__(self, cfg: SPTransformerConfig):
        super().__init__()
        self.cfg = cfg
        hd = cfg.dim // cfg.n_heads
        self.hd = hd
        self.wq = nn.Linear(cfg.dim, cfg.n_heads * hd, bias=False)
        self.wk = nn.Linear(cfg.dim, cfg.n_kv_heads * hd, bias=False)
        self.wv = nn.Linear(cfg.dim, cfg.n_kv_heads * hd, bias=False)
        self.wo = nn.Linear(cfg.n_heads * hd, cfg.dim, bias=False)
        self.n1, self.n2 = nn.LayerNorm(cfg.dim), nn.LayerNorm(cfg.dim)
        self.up = nn.Linear(cfg.dim, cfg.ffn_mult * cfg.dim, bias=False)
        self.down = nn.Linear(cfg.ffn_mult * cfg.dim, cfg.dim, bias=False)
        self.attn = LongContextAttention(ring_impl_type=cfg.ring_impl_type, attn_type=cfg.attn_type,
                                         backend=cfg.backend)

    def forward(self, x, pos):
        B, S, _ = x.shape
        h = self.n1(x)
        q = rope(self.wq(h).view(B, S, self.cfg.n_heads, self.hd), pos, self.cfg.rope_theta)
        k = rope(self.wk(h).view(B, S, self.cfg.n_kv_heads, self.hd), pos, self.cfg.rope_theta)
        v = self.wv(h).view(B, S, self.cfg.n_kv_heads, self.hd)
        a = self.attn(q, k, v, causal=True)
        x = x + self.wo(a.reshape(B, S, -1))
        return x + self.down(F.gelu(self.up(self.n2(x))))


class SPTransformerLM(nn.Module):
    """forward(tokens_local (B, S/P), global_seqlen) -> logits for the local tokens."""

    def __init__(self, cfg: SPTransformerConfig):
        super().__init__()
        self.cfg = cfg
        self.emb = nn.Embedding(cfg.vocab_size, cfg.dim)
        self.blocks = nn.ModuleList([SPBlock(cfg) for _ in range(cfg.n_layers)])
        self.norm = nn.LayerNorm(cfg.dim)
        self.head = nn.Linear(cfg.dim, cfg.vocab_size, bias=False)

    def local_positions(self, global_seqlen: int, device) -> torch.Tensor:
        m = PROCESS_GROUP.mesh
        if m is None:
            return torch.arange(global_seqlen, device=device)
        return local_token_index(self.cfg.ring_impl_type, global_seqlen, m.ulysses_rank, m.ring_rank,
                                 m.ulysses_degree, m.ring_degree).to(device)

    def forward(self, tokens_local: torch.Tensor, global_seqlen: int) -> torch.Tensor:
        pos = self.local_positions(global_seqlen, tokens_local.device)
        x = self.emb(tokens_local)
        for b in self.blocks:
            x = b(x, pos)
        return self.head(self.norm(x))


def allreduce_sp_grads(model: nn.Module, group=None, average: bool = False) -> None:
    """Sum (or average) parameter gradients over the sequence-parallel group: every rank saw a different
    slice of the same sequence, so the full-sequence gradient is the sum of the per-rank gradients."""
    group = group if group is not None else PROCESS_GROUP.SP_PG
    if group is None or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    n = dist.get_world_size(group)
    for p in model.parameters():
        if p.grad is not None:
            dist.all_reduce(p.grad, group=group)
            if average:
                p.grad.div_(n)
