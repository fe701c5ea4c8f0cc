"""``yunchang.ring.ring_flashinfer_attn`` module path (reference ``ring/ring_flashinfer_attn.py``): on B200 the
flashinfer-flavoured entry points run the native engine."""
from ..parallel.ring_attention import RingAttnFunc as RingFlashInferAttnFunc  # noqa: F401
from . import (ring_flashinfer_attn_func, ring_flashinfer_attn_kvpacked_func,  # noqa: F401
               ring_flashinfer_attn_qkvpacked_func)
from ._lowlevel import make_dense as _make

ring_flashinfer_attn_forward, ring_flashinfer_attn_backward = _make("basic")
