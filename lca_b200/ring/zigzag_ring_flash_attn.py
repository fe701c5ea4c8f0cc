"""``yunchang.ring.zigzag_ring_flash_attn`` module path (reference ``ring/zigzag_ring_flash_attn.py``)."""
from ..parallel.ring_attention import (RingAttnFunc as ZigZagRingFlashAttnFunc,  # noqa: F401
                                       zigzag_ring_flash_attn_func, zigzag_ring_flash_attn_kvpacked_func,
                                       zigzag_ring_flash_attn_qkvpacked_func)
from ._lowlevel import make_dense as _make

zigzag_ring_flash_attn_forward, zigzag_ring_flash_attn_backward = _make("zigzag")
