"""fp32 oracle for sequences too long to materialise S x S scores for every head: evaluates ONE (batch, head) of
global attention in query chunks and returns the exact ``out`` / ``lse`` for sampled query rows, ``dq`` for the same
rows and ``dk`` / ``dv`` for sampled key rows.  Used by ``bench.py`` (the correctness line of the driver-run JSON) and by
the long-sequence cases of ``tests/test_fused_multigpu.py``.

The reference's tests compare against ``flash_attn_func`` on one GPU at S = 3816 (``test/test_hybrid_attn.py:60-140``)
and never at the sequence lengths its benchmarks run; this oracle scales to S = 256K+ because it only ever holds a
``chunk x S`` score block.

Everything is plain PyTorch in fp32 (TF32 off).  Masks: causal and sliding window on GLOBAL positions (tensors are in
natural token order); softcap supported; no ALiBi / dropout (not part of the long-sequence configurations).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch


def _scores(qc, k, i0, causal, window, scale, softcap):
    """qc (n, D) fp32 rows at global positions i0.. , k (S, D) fp32 -> masked scaled scores (n, S) fp32."""
    s = (qc @ k.t()) * scale
    if softcap and softcap > 0:
        s = softcap * torch.tanh(s / softcap)
    n, S = s.shape
    if causal or window[0] >= 0 or window[1] >= 0:
        qi = (i0 if torch.is_tensor(i0) else torch.arange(i0, i0 + n, device=s.device)).view(-1, 1)
        kj = torch.arange(S, device=s.device).view(1, -1)
        rel = kj - qi
        mask = torch.zeros_like(s, dtype=torch.bool)
        if causal:
            mask |= rel > 0
        if window[0] >= 0:
            mask |= rel < -window[0]
        if window[1] >= 0 and not causal:
            mask |= rel > window[1]
        s = s.masked_fill(mask, float("-inf"))
    return s


@torch.no_grad()
def head_oracle(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, do: Optional[torch.Tensor], rows: torch.Tensor,
                cols: Optional[torch.Tensor] = None, causal: bool = True, window: Tuple[int, int] = (-1, -1),
                softmax_scale: Optional[float] = None, softcap: float = 0.0, chunk: int = 1024) -> Dict[str, torch.Tensor]:
    """One kv head of global attention.

    q, do : (S, G, D) the G query heads that share the kv head (any float dtype, natural token order)
    k, v  : (S, D)
    rows  : (n,) int64 global query positions to return out / lse / dq for
    cols  : (m,) int64 global key positions to return dk / dv for (needs ``do``; costs one full chunked pass)

    -> dict(out (n, G, D), lse (n, G), dq (n, G, D), dk (m, D), dv (m, D)) fp32.
    """
    old_tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        S, G, D = q.shape
        scale = float(softmax_scale) if softmax_scale is not None else D ** -0.5
        kf, vf = k.float(), v.float()
        res: Dict[str, torch.Tensor] = {}
        rows = rows.to(q.device)
        out_r = torch.zeros(rows.numel(), G, D, device=q.device)
        lse_r = torch.zeros(rows.numel(), G, device=q.device)
        dq_r = torch.zeros_like(out_r) if do is not None else None
        for g in range(G):
            qs = q[rows, g].float()
            s = _scores(qs, kf, rows, causal, window, scale, softcap)
            lse = torch.logsumexp(s, dim=-1)
            p = torch.exp(s - lse.view(-1, 1))
            p = torch.where(torch.isfinite(lse).view(-1, 1), p, torch.zeros_like(p))
            o = p @ vf
            out_r[:, g], lse_r[:, g] = o, lse
            if do is not None:
                if softcap and softcap > 0:
                    raise NotImplementedError("sampled gradients with softcap")
                dor = do[rows, g].float()
                delta = (dor * o).sum(-1, keepdim=True)
                ds = p * (dor @ vf.t() - delta)
                dq_r[:, g] = (ds @ kf) * scale
        res["out"], res["lse"] = out_r, lse_r
        if dq_r is not None:
            res["dq"] = dq_r
        if cols is not None:
            if do is None:
                raise ValueError("dk/dv need dO")
            if softcap and softcap > 0:
                raise NotImplementedError("sampled gradients with softcap")
            cols = cols.to(q.device)
            kc, vc = kf[cols], vf[cols]
            dk = torch.zeros(cols.numel(), D, device=q.device)
            dv = torch.zeros_like(dk)
            cmin = int(cols.min())
            for g in range(G):
                for i0 in range(0, S, chunk):
                    i1 = min(S, i0 + chunk)
                    if causal and i1 - 1 < cmin:
                        continue                       # no sampled key is visible to these queries
                    # keys beyond the chunk's last query are masked under causality: skip their columns
                    kend = i1 if causal and window[1] < 0 else S
                    qc = q[i0:i1, g].float()
                    s = _scores(qc, kf[:kend], i0, causal, window, scale, 0.0)
                    lse = torch.logsumexp(s, dim=-1, keepdim=True)
                    ok = torch.isfinite(lse)
                    p = torch.where(ok, torch.exp(s - lse), torch.zeros_like(s))
                    o = p @ vf[:kend]
                    doc = do[i0:i1, g].float()
                    delta = (doc * o).sum(-1, keepdim=True)
                    sel = cols < kend
                    if not bool(sel.any()):
                        continue
                    cidx = cols[sel]
                    pc = p[:, cidx]                                        # (chunk, m')
                    dv[sel] += pc.t() @ doc
                    dsc = pc * (doc @ vc[sel].t() - delta)
                    dk[sel] += (dsc.t() @ qc) * scale
            res["dk"], res["dv"] = dk, dv
        return res
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old_tf32


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a - b| / max |b| (the scale-aware error the GPU tests use for gradients)."""
    return float((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-6))
