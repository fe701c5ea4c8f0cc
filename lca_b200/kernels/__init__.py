"""Kernel selection API, name-compatible with ``yunchang.kernels`` (``kernels/__init__.py:38-295``).

On B200 there is exactly one production attention engine -- the in-tree tcgen05 kernels -- so the
15-member ``AttnType`` enum of the reference is kept for source compatibility and mapped as:

=====================  ==============================================================
AttnType               engine here
=====================  ==============================================================
FA, FA3, FLASHINFER    native sm_100a tcgen05 kernel (torch oracle on CPU)
SAGE_* / SPARSE_SAGE   native kernel (bf16/fp16 math; quantised variants are future work)
TORCH / TORCH_*        pure-PyTorch engine with a TRUE log-sum-exp (CPU capable)
AITER, NPU             other vendors' hardware -> ValueError
=====================  ==============================================================

``AttnType.TORCH`` (value ``"torch"``) exists because the reference's own benchmark offers
``--attn_type torch`` although its enum has no such member.
"""
from __future__ import annotations

from enum import Enum

import torch

from .attention import (flash_attn3_func_backward, flash_attn3_func_forward, flash_attn_backward,
                        flash_attn_forward, flash_attn_forward_aiter, flash_attn_func, flashinfer_attn_backbward,
                        flashinfer_attn_backward, flashinfer_attn_forward, npu_fused_attn_backward,
                        npu_fused_attn_forward, pytorch_attn_backward, pytorch_attn_forward, pytorch_attn_func)


class AttnType(Enum):
    AITER = "aiter"
    FA = "fa"
    FA3 = "fa3"
    FLASHINFER = "flashinfer"
    TORCH = "torch"
    TORCH_MATH = "torch_math"
    TORCH_FLASH = "torch_flash"
    TORCH_EFFICIENT = "torch_efficient"
    TORCH_CUDNN = "torch_cudnn"
    SAGE_AUTO = "sage_auto"
    SAGE_FP16 = "sage_fp16"
    SAGE_FP16_TRITON = "sage_fp16_triton"
    SAGE_FP8 = "sage_fp8"
    SAGE_FP8_SM90 = "sage_fp8_sm90"
    SPARSE_SAGE = "sparse_sage"
    NPU = "npu"

    @classmethod
    def from_string(cls, s: str):
        for member in cls:
            if member.value == s:
                return member
        raise ValueError(f"'{s}' is not a valid {cls.__name__}")


def flash_attn_forward_fp8(q, k, v, dropout_p=0.0, softmax_scale=None, causal=False, window_size=(-1, -1), softcap=0.0,
                           alibi_slopes=None, return_softmax=False):
    """Uniform-contract forward on block-scaled e4m3 operands (quantised on the fly): ``-> (out, lse)``."""
    from ..ops.attention import AttnParams
    from ..ops.fp8 import attn_fp8_fwd
    from ..parallel.layout import Seg
    p = AttnParams.make(q, softmax_scale, causal, window_size, softcap, alibi_slopes)
    Sq, Sk = q.shape[1], k.shape[1]
    return attn_fp8_fwd(q, k, v, (Seg(max(Sk - Sq, 0), Sq, 1),), (Seg(0, Sk, 1),), p)


_FOREIGN = {AttnType.AITER: "AMD ROCm (aiter)", AttnType.NPU: "Ascend NPU"}
_QUANT_FWD = (AttnType.SAGE_AUTO, AttnType.SAGE_FP16, AttnType.SAGE_FP16_TRITON, AttnType.SAGE_FP8, AttnType.SAGE_FP8_SM90)
_FP8_FWD = (AttnType.SAGE_AUTO, AttnType.SAGE_FP8, AttnType.SAGE_FP8_SM90)


def is_torch_type(t) -> bool:
    return isinstance(t, AttnType) and t.value.startswith("torch")


def select_flash_attn_impl(impl_type: AttnType, stage: str = "fwd-bwd", attn_processor: torch.nn.Module = None):
    """Return a callable for ``stage`` in {"fwd-only", "bwd-only", "fwd-bwd"} (``kernels/__init__.py:63-65``)."""
    if not isinstance(impl_type, AttnType):
        # the reference's last resort (``kernels/__init__.py:292-295``): an unknown implementation tag with a
        # user-supplied attention module returns that module
        if attn_processor is not None:
            return attn_processor
        raise ValueError(f"Unknown flash attention implementation: {impl_type}")
    if impl_type in _FOREIGN:
        raise ValueError(f"AttnType.{impl_type.name} targets {_FOREIGN[impl_type]}; not available in the B200 build")
    if stage not in ("fwd-only", "bwd-only", "fwd-bwd"):
        raise ValueError(f"Unknown stage: {stage}")
    torch_like = is_torch_type(impl_type)
    if impl_type == AttnType.SPARSE_SAGE:
        # the user's sparse-attention module is the kernel (reference :255-277); forward only, no LSE
        if attn_processor is None or not callable(attn_processor):
            raise ImportError("SparseSageAttention is only available with a sparse attention processor passed in")
        if stage != "fwd-only":
            raise ValueError(f"Unknown/Unsupported stage: {stage}")

        def fn(q, k, v, causal=False, softmax_scale=None, *args, **kwargs):
            return attn_processor(q, k, v, is_causal=causal, scale=softmax_scale, tensor_layout="NHD"), None

        return fn
    if impl_type in _QUANT_FWD:
        # the reference's quantised family is forward-only third-party kernels (SageAttention int8/fp8, none of which
        # has an sm_100 build).  B200-native equivalent: block-scaled e4m3 Q/K/V on tcgen05 kind::f8f6f4
        # (ops/fp8.py).  SAGE_FP16 / SAGE_FP16_TRITON (int8 QK^T with a 16-bit PV product) run the 16-bit kernel:
        # B200 has no int8 tensor-core path worth taking and 16-bit QK^T is the higher-precision superset.
        from ..ops import fp8
        from .attention import _func
        if impl_type in _FP8_FWD and fp8.enabled():
            if stage == "fwd-only":
                return flash_attn_forward_fp8
            if stage == "fwd-bwd":       # e4m3 forward, 16-bit backward on the saved operands
                return lambda q, k, v, *a, **kw: _func("fp8", q, k, v, *a, **kw)
    if stage == "fwd-only":
        return pytorch_attn_forward if torch_like else flash_attn_forward
    if stage == "bwd-only":
        return pytorch_attn_backward if torch_like else flash_attn_backward
    return pytorch_attn_func if torch_like else flash_attn_func


# the reference's kernels package also re-exports the feature probes and the torch ring function
from ..globals import (HAS_AITER, HAS_FLASH_ATTN, HAS_FLASH_ATTN_HOPPER, HAS_FLASHINFER, HAS_NPU,  # noqa: E402,F401
                       HAS_SAGE_ATTENTION, HAS_SPARSE_SAGE_ATTENTION)


def __getattr__(name):          # lazy: lca_b200.ring imports this package
    if name == "ring_pytorch_attn_func":
        from ..ring import ring_pytorch_attn_func
        return ring_pytorch_attn_func
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")


__all__ = ["AttnType", "select_flash_attn_impl", "flash_attn_forward", "flash_attn_backward", "flash_attn_func",
           "pytorch_attn_forward", "pytorch_attn_backward", "pytorch_attn_func", "flash_attn3_func_forward",
           "flash_attn3_func_backward", "flashinfer_attn_forward", "flashinfer_attn_backbward", "flashinfer_attn_backward",
           "flash_attn_forward_aiter", "npu_fused_attn_forward", "npu_fused_attn_backward"]
