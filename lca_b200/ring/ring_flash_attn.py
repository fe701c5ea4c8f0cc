"""``yunchang.ring.ring_flash_attn`` module path (basic ring; reference ``ring/ring_flash_attn.py``)."""
from ..parallel.ring_attention import (RingAttnFunc as RingFlashAttnFunc, ring_flash_attn_func,  # noqa: F401
                                       ring_flash_attn_kvpacked_func, ring_flash_attn_qkvpacked_func)
from ._lowlevel import make_dense as _make

ring_flash_attn_forward, ring_flash_attn_backward = _make("basic")
