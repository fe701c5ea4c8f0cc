"""``yunchang.ring.stripe_flash_attn`` module path (reference ``ring/stripe_flash_attn.py``)."""
from ..parallel.ring_attention import (RingAttnFunc as StripeFlashAttnFunc, stripe_flash_attn_func,  # noqa: F401
                                       stripe_flash_attn_kvpacked_func, stripe_flash_attn_qkvpacked_func)
from ._lowlevel import make_dense as _make

stripe_flash_attn_forward, stripe_flash_attn_backward = _make("stripe")
