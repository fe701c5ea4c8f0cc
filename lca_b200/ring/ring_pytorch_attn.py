"""``yunchang.ring.ring_pytorch_attn`` module path (reference ``ring/ring_pytorch_attn.py``)."""
from ..parallel.ring_attention import RingAttnFunc as RingAttentionFunc  # noqa: F401
from . import ring_pytorch_attn_func  # noqa: F401
