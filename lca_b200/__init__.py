"""lca_b200 -- Blackwell-native unified sequence-parallel attention (USP: Ulysses x Ring).

Drop-in API of ``yunchang`` 0.6.4 (``yunchang/__init__.py:1-12``): ``import lca_b200 as yunchang``.
"""
from .globals import (HAS_AITER, HAS_FLASH_ATTN, HAS_FLASH_ATTN_HOPPER, HAS_FLASHINFER, HAS_NPU,
                      HAS_SAGE_ATTENTION, HAS_SPARSE_SAGE_ATTENTION, PROCESS_GROUP, has_native_kernels,
                      set_seq_parallel_pg)
from .kernels import AttnType, select_flash_attn_impl
from .comm import (EXTRACT_FUNC_DICT, SeqAllToAll4D, SeqAllToAll5D, basic_extract_local, gather_global,
                   local_token_index, stripe_extract_local, zigzag_extract_local)
from .ring import *  # noqa: F401,F403  (the reference star-exports every ring function)
from .ring import (ring_flash_attn_func, ring_flash_attn_kvpacked_func, ring_flash_attn_qkvpacked_func,
                   ring_flash_attn_varlen_func, ring_flash_attn_varlen_kvpacked_func,
                   ring_flash_attn_varlen_qkvpacked_func, ring_flashinfer_attn_func,
                   ring_flashinfer_attn_kvpacked_func, ring_flashinfer_attn_qkvpacked_func,
                   ring_npu_flash_attn_func, ring_pytorch_attn_func, stripe_flash_attn_func,
                   stripe_flash_attn_kvpacked_func, stripe_flash_attn_qkvpacked_func,
                   zigzag_ring_flash_attn_func, zigzag_ring_flash_attn_kvpacked_func,
                   zigzag_ring_flash_attn_qkvpacked_func, zigzag_ring_flash_attn_varlen_func,
                   zigzag_ring_flash_attn_varlen_kvpacked_func, zigzag_ring_flash_attn_varlen_qkvpacked_func)
from .hybrid import (AsyncLongContextAttention, LongContextAttention, LongContextAttentionQKVPacked,
                     RING_IMPL_QKVPACKED_DICT)
from .ulysses import UlyssesAttention

__version__ = "0.1.0"
