"""Counter-based attention dropout keyed on GLOBAL coordinates.

The reference hands ``dropout_p`` to flash-attn once per ring step, so every (rank, step) block draws an unrelated
Philox stream and its ring backward passes ``rng_state=None`` (``ring_flash_attn.py:104-130``): a sequence-parallel
run can neither be reproduced on one device nor re-generate its forward masks in the backward.  Here the keep
decision of a score is a pure function of ``(seed, batch, global head, global query position, global key position)``:

    word = mix32( qpos * 0x9E3779B1  ^  (kpos >> 2) * 0x85EBCA77  ^  seed  ^  ((batch << 16 | head) * 0xC2B2AE3D) )
    keep = byte[kpos & 3] of word  >=  p8              p8 = round(256 * dropout_p)

(``mix32`` = the "lowbias32" integer finaliser; all arithmetic modulo 2**32.)  Whatever the Ulysses x Ring layout,
ring flavour or tile schedule, every rank / tile / pass regenerates the same mask, so distributed dropout equals
single-device dropout bit for bit and the backward needs no saved RNG state.  The probability is quantised to
1/256 (like flash-attn's 8-bit thresholds); ``keep_scale`` uses the quantised value, so the estimator stays unbiased.
One 32-bit word serves four consecutive key positions: a kernel thread that owns a score row hashes once per four
columns (8 integer instructions), and the key is a plain XOR of per-coordinate terms, so a thread that owns a KEY row
(the dK/dV pass of the backward) pays one hash per score without any per-query pre-hash.

This module is the executable specification: the PyTorch engine calls it directly, the CUDA kernels implement the
same integer recipe (``csrc/sm100_ptx.cuh: dropout_*``).
"""
from __future__ import annotations

from typing import Optional

import torch

_M = 0xFFFFFFFF
K_Q, K_K, K_BH = 0x9E3779B1, 0x85EBCA77, 0xC2B2AE3D


def p8_of(dropout_p: float) -> int:
    """8-bit drop threshold: a score is dropped when its byte is < p8, i.e. with probability p8/256."""
    return max(0, min(255, int(round(float(dropout_p) * 256.0))))


def p_eff(dropout_p: float) -> float:
    return p8_of(dropout_p) / 256.0


def keep_scale(dropout_p: float) -> float:
    return 256.0 / (256 - p8_of(dropout_p))


def _mix32(x: torch.Tensor) -> torch.Tensor:
    x = x ^ (x >> 16)
    x = (x * 0x7FEB352D) & _M
    x = x ^ (x >> 15)
    x = (x * 0x846CA68B) & _M
    return x ^ (x >> 16)


def keep_mask(seed: int, B: int, H: int, q_pos: torch.Tensor, k_pos: torch.Tensor, dropout_p: float,
              head_offset: int = 0, batch_offset: int = 0, q_grp: Optional[torch.Tensor] = None) -> torch.Tensor:
    """-> bool ``(B, H, Sq, Sk)``, True = keep.  ``q_pos`` / ``k_pos`` are the global token positions of the rows /
    columns (int tensors); ``head_offset`` is the global index of local head 0 (Ulysses head shards); packed varlen
    batches pass ``q_grp`` (sequence id per row), which takes the place of the batch index."""
    dev = q_pos.device
    qp = q_pos.to(torch.int64).view(1, 1, -1)
    b = torch.arange(B, device=dev, dtype=torch.int64).view(-1, 1, 1) + int(batch_offset)
    if q_grp is not None:
        b = b + q_grp.to(torch.int64).view(1, 1, -1)
    h = torch.arange(H, device=dev, dtype=torch.int64).view(1, -1, 1) + int(head_offset)
    bh = ((b << 16) | h) & _M
    row = ((qp * K_Q) & _M) ^ (int(seed) & _M) ^ ((bh * K_BH) & _M)                     # (B,H,Sq)
    kp = k_pos.to(torch.int64).view(1, 1, 1, -1)
    word = _mix32(row.unsqueeze(-1) ^ (((kp >> 2) * K_K) & _M))                          # (B,H,Sq,Sk)
    byte = (word >> ((kp & 3) * 8)) & 0xFF
    return byte >= p8_of(dropout_p)
