"""``yunchang.ring.zigzag_ring_flash_attn_varlen`` module path (reference ``ring/zigzag_ring_flash_attn_varlen.py``).

``get_half_index`` / ``get_half_lse`` are the reference's helpers for slicing the "second half" of every packed
sequence (``:27-59``); the ring here never needs them (positions carry that information) but they are provided
for callers that use them directly."""
import torch

from ..parallel.ring_varlen import (RingVarlenAttnFunc as ZigZagRingFlashAttnVarlenFunc,  # noqa: F401
                                    zigzag_ring_flash_attn_varlen_func, zigzag_ring_flash_attn_varlen_kvpacked_func,
                                    zigzag_ring_flash_attn_varlen_qkvpacked_func)
from ._lowlevel import make_varlen as _make

zigzag_ring_flash_attn_varlen_forward, zigzag_ring_flash_attn_varlen_backward = _make("zigzag", True)


def get_half_index(cu_seqlens, *, front: bool):
    """Index of the front / back half of every packed sequence: one ``slice`` when there is a single sequence,
    otherwise a boolean mask over the packed token dimension."""
    cu = [int(x) for x in (cu_seqlens.tolist() if hasattr(cu_seqlens, "tolist") else cu_seqlens)]
    if len(cu) == 2:
        half = cu[1] // 2
        return slice(None, half) if front else slice(half, None)
    mask = torch.zeros(cu[-1], dtype=torch.bool)
    for a, b in zip(cu[:-1], cu[1:]):
        mid = (a + b) // 2
        if front:
            mask[a:mid] = True
        else:
            mask[mid:b] = True
    return mask


def get_half_lse(lse, cu_seqlens, *, front: bool):
    """Rows of a padded ``(num_seq, H, max_seqlen)`` LSE that belong to the front / back half of each sequence,
    packed to ``(num_seq, H, max_seqlen // 2)``."""
    cu = [int(x) for x in (cu_seqlens.tolist() if hasattr(cu_seqlens, "tolist") else cu_seqlens)]
    n, H, L = lse.shape
    out = torch.zeros(n, H, L // 2, dtype=lse.dtype, device=lse.device)
    for i, (a, b) in enumerate(zip(cu[:-1], cu[1:])):
        ln = b - a
        lo, hi = (0, ln // 2) if front else (ln // 2, ln)
        out[i, :, :hi - lo] = lse[i, :, lo:hi]
    return out
