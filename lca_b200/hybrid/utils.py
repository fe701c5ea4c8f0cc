"""Ring implementation registries (``yunchang/hybrid/utils.py:1-28``).

Keys are the reference's (including the ``"strip"`` spelling); ``"stripe"`` is accepted too.  The
vendor-specific entries (``basic_pytorch`` / ``basic_flashinfer`` / ``basic_npu``) resolve to the
same position-aware ring loop -- the engine is chosen by ``attn_type``, not by the registry key.
Unlike the reference, every key is usable with packed QKV.
"""
from ..parallel.ring_attention import (ring_flash_attn_func, ring_flash_attn_qkvpacked_func,
                                       stripe_flash_attn_func, stripe_flash_attn_qkvpacked_func,
                                       zigzag_ring_flash_attn_func, zigzag_ring_flash_attn_qkvpacked_func)
from ..ring import (ring_flashinfer_attn_func, ring_flashinfer_attn_qkvpacked_func, ring_npu_flash_attn_func,
                    ring_pytorch_attn_func)

RING_IMPL_DICT = {
    "basic": ring_flash_attn_func,
    "zigzag": zigzag_ring_flash_attn_func,
    "strip": stripe_flash_attn_func,
    "stripe": stripe_flash_attn_func,
    "basic_pytorch": ring_pytorch_attn_func,
    "basic_flashinfer": ring_flashinfer_attn_func,
    "basic_npu": ring_npu_flash_attn_func,       # raises: Ascend-only in the reference
}

RING_IMPL_QKVPACKED_DICT = {
    "basic": ring_flash_attn_qkvpacked_func,
    "zigzag": zigzag_ring_flash_attn_qkvpacked_func,
    "strip": stripe_flash_attn_qkvpacked_func,
    "stripe": stripe_flash_attn_qkvpacked_func,
    "basic_flashinfer": ring_flashinfer_attn_qkvpacked_func,
}
