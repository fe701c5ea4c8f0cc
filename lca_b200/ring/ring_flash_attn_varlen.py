"""``yunchang.ring.ring_flash_attn_varlen`` module path (reference ``ring/ring_flash_attn_varlen.py``)."""
from ..parallel.ring_varlen import (RingVarlenAttnFunc as RingFlashAttnVarlenFunc,  # noqa: F401
                                    ring_flash_attn_varlen_func, ring_flash_attn_varlen_kvpacked_func,
                                    ring_flash_attn_varlen_qkvpacked_func)
from ._lowlevel import make_varlen as _make

ring_flash_attn_varlen_forward, ring_flash_attn_varlen_backward = _make("basic", False)
