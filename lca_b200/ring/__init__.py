"""``yunchang.ring``-compatible namespace: every ring function of the reference
(``ring/__init__.py:1-38``) plus the varlen kvpacked variant it forgets to export."""
from functools import partial as _partial

from ..kernels import AttnType as _AttnType
from ..parallel.ring_attention import (RingAttnFunc, ring_attn_backward, ring_attn_forward,
                                       ring_flash_attn_func, ring_flash_attn_kvpacked_func,
                                       ring_flash_attn_qkvpacked_func, stripe_flash_attn_func,
                                       stripe_flash_attn_kvpacked_func, stripe_flash_attn_qkvpacked_func,
                                       zigzag_ring_flash_attn_func, zigzag_ring_flash_attn_kvpacked_func,
                                       zigzag_ring_flash_attn_qkvpacked_func)
from ..parallel.ring_comm import RingComm
from ..parallel.ring_varlen import (ring_flash_attn_varlen_func, ring_flash_attn_varlen_kvpacked_func,
                                    ring_flash_attn_varlen_qkvpacked_func, zigzag_ring_flash_attn_varlen_func,
                                    zigzag_ring_flash_attn_varlen_kvpacked_func,
                                    zigzag_ring_flash_attn_varlen_qkvpacked_func)
from .utils import flatten_varlen_lse, unflatten_varlen_lse, update_npu_out, update_out_and_lse


def ring_pytorch_attn_func(q, k, v, dropout_p=0.0, softmax_scale=None, causal=False, window_size=(-1, -1),
                           softcap=0.0, alibi_slopes=None, deterministic=False, return_attn_probs=False, group=None,
                           attn_type=None, attn_processor=None, **kw):
    """Basic ring on the pure-PyTorch engine, forward AND backward (reference: ``ring_pytorch_attn.py``,
    whose backward is unreachable)."""
    if attn_type is None or not getattr(attn_type, "value", "").startswith("torch"):
        attn_type = _AttnType.TORCH
    return ring_flash_attn_func(q, k, v, dropout_p, softmax_scale, causal, window_size, softcap, alibi_slopes,
                                deterministic, return_attn_probs, group, attn_type, attn_processor, **kw)


# flashinfer-flavoured entry points of the reference (``ring_flashinfer_attn.py``) -- on B200 they run the
# native engine; kept so call sites keep working.
ring_flashinfer_attn_func = ring_flash_attn_func
ring_flashinfer_attn_kvpacked_func = ring_flash_attn_kvpacked_func
ring_flashinfer_attn_qkvpacked_func = ring_flash_attn_qkvpacked_func


def ring_npu_flash_attn_func(*args, **kwargs):
    raise RuntimeError("ring_npu_flash_attn_func targets Ascend NPUs; use ring_flash_attn_func on B200")
