"""Megatron / DeepSpeed style integration shims.

Role of the reference's ``patches/Megatron-DeepSpeed.patch`` (adds ``--ds-ring-sequence-parallel-size``, calls
``set_seq_parallel_pg`` inside ``initialize_model_parallel`` and replaces DeepSpeed's ``DistributedAttention`` by
``LongContextAttention()``).  ``patches/apply_megatron_deepspeed.py`` (+ the generated ``Megatron-DeepSpeed.patch``)
wires a Megatron-DeepSpeed checkout to the hooks below:

* :func:`initialize_sequence_parallel` -- call it where the framework builds its model-parallel groups;
* :class:`DistributedAttention` -- drop-in for ``deepspeed.sequence.layer.DistributedAttention``: it is
  constructed with a *local* attention module (ignored: the fused kernels do the local attention) and called
  as ``forward(query, key, value, *args)`` on ``(S/P, B, H, D)`` or ``(B, S/P, H, D)`` shards.
* :func:`shard_batch` -- the token layout of the ring dimension (zigzag / stripe load balancing) applied to a
  ``(B, S, ...)`` batch in ``get_batch`` (the reference's patch keeps ``ring_impl_type="basic"`` and so never reorders).
See docs/megatron_integration.md.
"""
from __future__ import annotations

from typing import Any, Optional

import torch
import torch.distributed as dist

from ..globals import set_seq_parallel_pg
from ..hybrid import LongContextAttention
from ..kernels import AttnType


def initialize_sequence_parallel(sequence_parallel_size: int, ring_sequence_parallel_size: int = 1,
                                 use_ulysses_low: bool = True) -> None:
    """``sequence_parallel_size`` = U * R (Megatron's ``--ds-sequence-parallel-size``), of which
    ``ring_sequence_parallel_size`` = R (the reference's ``--ds-ring-sequence-parallel-size``)."""
    if sequence_parallel_size % ring_sequence_parallel_size:
        raise ValueError("sequence_parallel_size must be a multiple of ring_sequence_parallel_size")
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    set_seq_parallel_pg(sequence_parallel_size // ring_sequence_parallel_size, ring_sequence_parallel_size, rank, world,
                        use_ulysses_low)


def shard_batch(x: torch.Tensor, rank: Optional[int] = None, world_size: Optional[int] = None,
                ulysses_degree: Optional[int] = None, ring_degree: Optional[int] = None,
                ring_impl_type: str = "zigzag", dim: int = 1) -> torch.Tensor:
    """This rank's shard of a global ``(B, S, ...)`` tensor (tokens, labels, position ids, loss mask) in the layout the
    attention expects: contiguous for ``basic``, two mirrored chunks for ``zigzag``, round-robin for ``stripe``.
    Defaults come from the mesh built by :func:`initialize_sequence_parallel`."""
    from ..globals import PROCESS_GROUP
    from ..parallel.layout import EXTRACT_FUNC_DICT
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    ud = ulysses_degree if ulysses_degree is not None else PROCESS_GROUP.ulysses_degree
    rd = ring_degree if ring_degree is not None else PROCESS_GROUP.ring_degree
    key = {"stripe": "strip"}.get(ring_impl_type, ring_impl_type)
    if dim != 1:
        return EXTRACT_FUNC_DICT[key](x.transpose(1, dim), rank, world_size, rd=rd, ud=ud).transpose(1, dim)
    return EXTRACT_FUNC_DICT[key](x, rank, world_size, rd=rd, ud=ud)


class DistributedAttention(torch.nn.Module):
    def __init__(self, local_attention: Optional[torch.nn.Module] = None, sequence_process_group=None,
                 scatter_idx: int = 2, gather_idx: int = 0, ring_impl_type: str = "zigzag", causal: bool = True,
                 attn_type: AttnType = AttnType.FA, seq_first: bool = True) -> None:
        super().__init__()
        self.seq_first = seq_first        # Megatron activations are (S, B, H, D)
        self.causal = causal
        self.attn = LongContextAttention(ring_impl_type=ring_impl_type, attn_type=attn_type)

    def forward(self, query: torch.Tensor, key: torch.Tensor, value: torch.Tensor, *args: Any, **kwargs: Any) -> torch.Tensor:
        if self.seq_first:
            query, key, value = (t.transpose(0, 1) for t in (query, key, value))
        out = self.attn(query, key, value, causal=kwargs.pop("causal", self.causal), **kwargs)
        return out.transpose(0, 1) if self.seq_first else out
