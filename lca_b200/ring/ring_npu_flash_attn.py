"""``yunchang.ring.ring_npu_flash_attn`` module path (reference ``ring/ring_npu_flash_attn.py``, Ascend only)."""
from . import ring_npu_flash_attn_func  # noqa: F401


def _ascend_only(*args, **kwargs):
    raise RuntimeError("the NPU ring targets Ascend hardware; use lca_b200.ring.ring_flash_attn on B200")


ring_npu_flash_attn_forward = ring_npu_flash_attn_backward = _ascend_only


class RingNpuFlashAttnFunc:
    apply = staticmethod(_ascend_only)
