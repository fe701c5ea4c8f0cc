"""Loader and thin Python wrappers for the in-tree sm_100a extension (``lca_b200/ops/_C*.so``).

Policy: on a machine with a Blackwell GPU the CUDA path is THE path -- if the extension is
missing or fails to import we raise (no silent eager fallback), unless the user explicitly opts
out with ``LCA_B200_ALLOW_FALLBACK=1``.  On CPU-only hosts everything routes to the torch engine.
"""
from __future__ import annotations

import importlib
import os
from typing import List, Optional, Tuple

import torch

from . import dropout as _dropout

from ..parallel.layout import PosSpec

_C = None
_LOAD_ERROR: Optional[str] = None
SUPPORTED_HEAD_DIMS = (64, 128)


def _load():
    global _C, _LOAD_ERROR
    if _C is not None or _LOAD_ERROR is not None:
        return _C
    try:
        _C = importlib.import_module("lca_b200.ops._C")
    except Exception as e:  # noqa: BLE001
        _LOAD_ERROR = f"{type(e).__name__}: {e}"
        _C = None
    return _C


LAUNCHES = 0     # number of in-tree CUDA kernels launched through this module (bench.py reports it)
_KERNEL_FUNCS = {"fmha_fwd", "fmha_bwd_pass", "merge_out_lse", "finalize_out", "flatten_varlen_lse",
                 "unflatten_varlen_lse", "permute_group", "attn_delta", "usp_fwd", "usp_bwd_pass", "symm_wait",
                 "fmha_fwd_drop", "fmha_bwd_pass_drop", "fmha_fwd_fp8", "quantize_e4m3"}


class _CountingExt:
    """Proxy over the extension module that counts kernel launches."""

    def __init__(self, mod):
        self._mod = mod

    def __getattr__(self, name):
        fn = getattr(self._mod, name)
        if name not in _KERNEL_FUNCS:
            return fn

        def wrapped(*a, **kw):
            global LAUNCHES
            LAUNCHES += 1
            return fn(*a, **kw)

        return wrapped


_PROXY = None


def ext():
    global _PROXY
    c = _load()
    if c is None:
        raise RuntimeError(
            f"lca_b200 native extension not loadable ({_LOAD_ERROR}); build it with "
            "`python -m lca_b200.ops.build`"
        )
    if _PROXY is None:
        _PROXY = _CountingExt(c)
    return _PROXY


def _is_blackwell(device) -> bool:
    try:
        major, _ = torch.cuda.get_device_capability(device)
    except Exception:  # noqa: BLE001
        return False
    return major == 10


def available() -> bool:
    """Extension importable and a sm_100 device visible."""
    if not torch.cuda.is_available():
        return False
    return _load() is not None and _is_blackwell(torch.cuda.current_device())


def extension_loaded() -> bool:
    return _load() is not None


def must_be_native() -> bool:
    """On a Blackwell box the native path is mandatory unless explicitly waived."""
    if os.environ.get("LCA_B200_ALLOW_FALLBACK", "0") == "1":
        return False
    return torch.cuda.is_available() and _is_blackwell(torch.cuda.current_device())


def why_not(q: torch.Tensor) -> str:
    if not q.is_cuda:
        return "tensor is not on CUDA"
    if _load() is None:
        return f"extension not loadable: {_LOAD_ERROR}"
    if not _is_blackwell(q.device):
        return "device is not sm_100"
    if q.dtype not in (torch.bfloat16, torch.float16):
        return f"dtype {q.dtype} (need bf16/fp16)"
    if q.shape[-1] > max(SUPPORTED_HEAD_DIMS) or q.shape[-1] % 8:
        return f"head_dim {q.shape[-1]} (native tiles: {SUPPORTED_HEAD_DIMS}; others are zero-padded up to 128)"
    return ""


def _padded_dim(D: int) -> int:
    return next(d for d in SUPPORTED_HEAD_DIMS if d >= D)


def _pad_last(t: torch.Tensor, Dp: int) -> torch.Tensor:
    return t if t.shape[-1] == Dp else torch.nn.functional.pad(t, (0, Dp - t.shape[-1]))


def supports(q: torch.Tensor) -> bool:
    return why_not(q) == ""


def has_bwd() -> bool:
    c = _load()
    return c is not None and hasattr(c, "fmha_bwd_pass") and os.environ.get("LCA_B200_TORCH_BWD", "0") != "1"


# ------------------------------------------------------------------------------------------
def _common_stride(spec: PosSpec) -> int:
    strides = {s.stride for s in spec if s.count > 1}
    if len(strides) > 1:
        raise ValueError(f"segments with different position strides: {spec}")
    return strides.pop() if strides else 1


def _rows(spec: PosSpec) -> List[Tuple[int, int, int, int]]:
    """-> [(row0, nrows, pos0, group)]"""
    out, off = [], 0
    for s in spec:
        if s.count > 0:
            out.append((off, s.count, s.start, s.group))
        off += s.count
    return out


MAX_SEG = 128


def window_bounds(p) -> Tuple[int, int]:
    wl, wr = p.window_size
    if p.causal:
        wr = 0 if wr < 0 else min(wr, 0)
    return int(wl), int(wr)


def _tma_ready(t: torch.Tensor) -> torch.Tensor:
    """TMA needs: last dim contiguous, other strides multiples of 8 elements, 16B-aligned base."""
    ok = t.stride(-1) == 1 and all(st % 8 == 0 for st in t.stride()[:-1]) and t.data_ptr() % 16 == 0
    return t if ok else t.contiguous()


def fmha_fwd(q, k, v, q_pos: PosSpec, k_pos: PosSpec, p, out=None, lse=None, sm_limit: int = 0):
    """Forward attention of one block with global-position masks.  -> (out, lse)."""
    C = ext()
    D0 = q.shape[-1]
    if D0 not in SUPPORTED_HEAD_DIMS:
        # head dims between the native tile widths run zero-padded (QK^T and PV are unchanged by zero columns;
        # the softmax scale is explicit in `p`)
        Dp = _padded_dim(D0)
        o, l = fmha_fwd(_pad_last(q, Dp), _pad_last(k, Dp), _pad_last(v, Dp), q_pos, k_pos, p, None, lse, sm_limit)
        if out is not None:
            out.copy_(o[..., :D0])
            return out, l
        return o[..., :D0].contiguous(), l
    q, k, v = _tma_ready(q), _tma_ready(k), _tma_ready(v)
    B, Sq, H, D = q.shape
    if out is None:
        out = torch.empty((B, Sq, H, D), dtype=q.dtype, device=q.device)
    if lse is None:
        lse = torch.empty((B, H, Sq), dtype=torch.float32, device=q.device)
    # heaviest (latest) segments first: the kernel walks q segments in the given order
    qrows = sorted(_rows(q_pos), key=lambda r: -r[2])
    krows = _rows(k_pos)
    wl, wr = window_bounds(p)
    alibi = p.alibi_slopes
    if alibi is not None:
        alibi = alibi.to(device=q.device, dtype=torch.float32).contiguous()
    qs, ks = _common_stride(q_pos), _common_stride(k_pos)
    # the kernel takes <= MAX_SEG segments per side; many-sequence varlen batches are issued in
    # chunks of whole groups (each launch writes disjoint output rows)
    for qchunk, kchunk in _chunk_by_group(qrows, krows):
        qsegs = [[r0, n, pos0, -1, r0, 0, 0, g] for (r0, n, pos0, g) in qchunk]
        ksegs = [[r0, n, pos0, -1, g] for (r0, n, pos0, g) in kchunk]
        drop = _drop_args(p)
        if drop is None:
            C.fmha_fwd(q, k, v, qsegs, ksegs, qs, ks, out, 0, lse, float(p.softmax_scale), wl, wr,
                       float(p.softcap), alibi, 0, 0, int(sm_limit))
        else:
            C.fmha_fwd_drop(q, k, v, qsegs, ksegs, qs, ks, out, 0, lse, float(p.softmax_scale), wl, wr,
                            float(p.softcap), alibi, 0, 0, int(sm_limit), drop)
    return out, lse


def dropout_supported(p) -> bool:
    """The ``kDrop`` kernel instantiations regenerate the coordinate-keyed keep mask of ``ops/dropout.py`` in registers.
    Default since round 2 (validated on hardware against the PyTorch engine, ``tests/test_dropout_gpu.py``;
    ``LCA_B200_NATIVE_DROPOUT=0`` opts out); softcap + dropout stays on the PyTorch engine."""
    return (os.environ.get("LCA_B200_NATIVE_DROPOUT", "1") == "1" and float(getattr(p, "softcap", 0.0)) == 0.0
            and _dropout.p8_of(p.dropout_p) > 0)


def _drop_args(p):
    if float(getattr(p, "dropout_p", 0.0)) <= 0.0 or _dropout.p8_of(p.dropout_p) == 0:
        return None
    return [_dropout.p8_of(p.dropout_p), int(p.dropout_seed) & 0xFFFFFFFF, int(p.head_offset)]


def _chunk_by_group(qrows, krows):
    if len(qrows) <= MAX_SEG and len(krows) <= MAX_SEG:
        yield qrows, krows
        return
    groups = sorted({r[3] for r in qrows})
    cur_q, cur_k = [], []
    for g in groups:
        gq = [r for r in qrows if r[3] == g]
        gk = [r for r in krows if r[3] == g]
        if len(gq) > MAX_SEG or len(gk) > MAX_SEG:
            raise ValueError("a single attention group has more than 32 segments")
        if not gk:
            gk = []
        if len(cur_q) + len(gq) > MAX_SEG or len(cur_k) + len(gk) > MAX_SEG:
            if cur_q and cur_k:
                yield cur_q, cur_k
            cur_q, cur_k = [], []
        cur_q += gq
        cur_k += gk
    if cur_q and cur_k:
        yield cur_q, cur_k


def attn_delta(out, dout, lse=None):
    """-> delta (B,H,S) fp32 [, lse2 (B,H,S) log2-domain LSE with +inf for key-less rows]."""
    dout = dout if dout.stride(-1) == 1 else dout.contiguous()
    out = out if out.stride(-1) == 1 else out.contiguous()
    r = ext().attn_delta(out, dout, None if lse is None else lse.contiguous())
    return r[0] if lse is None else (r[0], r[1])


def fmha_bwd(dout, q, k, v, out, lse, q_pos: PosSpec, k_pos: PosSpec, p, delta=None, lse2=None,
             dq=None, dk=None, dv=None, accumulate: bool = False, out_dtype=None, sm_limit: int = 0,
             acc_dq: Optional[bool] = None, acc_dkv: Optional[bool] = None):
    """Backward of one block: two tcgen05 passes (dQ, then dK/dV).  ``lse`` is the FINAL LSE of the
    query rows (may cover more keys than this block: ring steps).  Returns (dq, dk, dv) in
    ``out_dtype`` (default: input dtype; fp32 when ``accumulate``)."""
    C = ext()
    D0 = q.shape[-1]
    if D0 not in SUPPORTED_HEAD_DIMS:
        Dp = _padded_dim(D0)
        if delta is None or lse2 is None:
            delta, lse2 = attn_delta(out, dout, lse)
        gq, gk, gv = fmha_bwd(_pad_last(dout, Dp), _pad_last(q, Dp), _pad_last(k, Dp), _pad_last(v, Dp), None, lse,
                              q_pos, k_pos, p, delta=delta, lse2=lse2, out_dtype=out_dtype or (torch.float32 if (accumulate or acc_dq or acc_dkv) else q.dtype),
                              sm_limit=sm_limit)
        gq, gk, gv = gq[..., :D0], gk[..., :D0], gv[..., :D0]
        if dq is None:
            return gq.contiguous(), gk.contiguous(), gv.contiguous()
        a_q = accumulate if acc_dq is None else acc_dq
        a_kv = accumulate if acc_dkv is None else acc_dkv
        dq.add_(gq) if a_q else dq.copy_(gq)
        dk.add_(gk) if a_kv else dk.copy_(gk)
        dv.add_(gv) if a_kv else dv.copy_(gv)
        return dq, dk, dv
    q, k, v, dout = _tma_ready(q), _tma_ready(k), _tma_ready(v), _tma_ready(dout)
    if delta is None or lse2 is None:
        delta, lse2 = attn_delta(out, dout, lse)
    acc_dq = accumulate if acc_dq is None else acc_dq
    acc_dkv = accumulate if acc_dkv is None else acc_dkv
    if out_dtype is None:
        out_dtype = torch.float32 if (acc_dq or acc_dkv) else q.dtype
    if dq is None:
        assert not (acc_dq or acc_dkv)
        dq = torch.empty(q.shape, dtype=out_dtype, device=q.device)
        dk = torch.empty(k.shape, dtype=out_dtype, device=q.device)
        dv = torch.empty(v.shape, dtype=out_dtype, device=q.device)
    wl, wr = window_bounds(p)
    alibi = p.alibi_slopes
    if alibi is not None:
        alibi = alibi.to(device=q.device, dtype=torch.float32).contiguous()
    qs, ks = _common_stride(q_pos), _common_stride(k_pos)
    qrows, krows = _rows(q_pos), _rows(k_pos)
    for qchunk, kchunk in _chunk_by_group(qrows, krows):
        xq = [[r0, n, pos0, g, r0] for (r0, n, pos0, g) in sorted(qchunk, key=lambda r: -r[2])]
        yk = [[r0, n, pos0, -1, g] for (r0, n, pos0, g) in kchunk]
        drop = _drop_args(p)
        if drop is not None:
            C.fmha_bwd_pass_drop(False, q, dout, k, v, xq, yk, qs, ks, lse2, delta, dq, None, acc_dq,
                                 float(p.softmax_scale), wl, wr, float(p.softcap), alibi, int(sm_limit), drop)
            xk = [[r0, n, pos0, g, r0] for (r0, n, pos0, g) in kchunk]
            yq = [[r0, n, pos0, -1, g] for (r0, n, pos0, g) in qchunk]
            C.fmha_bwd_pass_drop(True, k, v, q, dout, xk, yq, ks, qs, lse2, delta, dk, dv, acc_dkv,
                                 float(p.softmax_scale), wr, wl, float(p.softcap), alibi, int(sm_limit), drop)
            continue
        C.fmha_bwd_pass(False, q, dout, k, v, xq, yk, qs, ks, lse2, delta, dq, None, acc_dq,
                        float(p.softmax_scale), wl, wr, float(p.softcap), alibi, int(sm_limit))
        xk = [[r0, n, pos0, g, r0] for (r0, n, pos0, g) in kchunk]
        yq = [[r0, n, pos0, -1, g] for (r0, n, pos0, g) in qchunk]
        # rows are keys, columns are queries: bounds on (qpos - kpos) are the mirrored window
        C.fmha_bwd_pass(True, k, v, q, dout, xk, yq, ks, qs, lse2, delta, dk, dv, acc_dkv,
                        float(p.softmax_scale), wr, wl, float(p.softcap), alibi, int(sm_limit))
    return dq, dk, dv


def merge_out_lse_(out_acc, lse_acc, block_out, block_lse) -> None:
    D = out_acc.shape[-1]
    if D % 4 == 0 and D // 4 <= 32 and 32 % (D // 4) == 0 and out_acc.is_contiguous() and lse_acc.is_contiguous():
        ext().merge_out_lse(out_acc, lse_acc, block_out.contiguous(), block_lse.contiguous())
    else:  # odd head dims: same math in torch
        new = torch.logaddexp(lse_acc, block_lse)
        safe = torch.where(torch.isinf(new) & (new < 0), torch.zeros_like(new), new)
        w_old = torch.exp(lse_acc - safe).transpose(1, 2).unsqueeze(-1)
        w_new = torch.exp(block_lse - safe).transpose(1, 2).unsqueeze(-1)
        out_acc.mul_(w_old).add_(block_out.to(torch.float32) * w_new)
        lse_acc.copy_(new)


def finalize_out(out_acc: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    if out_acc.is_cuda and available() and out_acc.is_contiguous() and out_acc.numel() % 4 == 0:
        return ext().finalize_out(out_acc, dtype)
    return out_acc.to(dtype)
