"""Fused NVLink USP engine: Ulysses head shuffle + ring K/V exchange + attention in ONE kernel per rank.

What replaces what (reference call sites in parentheses):

* 3x ``all_to_all_single`` + 6 staging copies before attention (``hybrid/attn_layer.py:111-119``,
  ``comm/all_to_all.py:45-65``)  ->  communication CTAs of the attention kernel push each head-slice
  straight from the caller's q/k/v tensors into the consumer rank's staging buffer with 16-byte
  ``st.global`` over NVLink (no NCCL, no intermediate copy, strided packed-QKV views accepted).
* ``R-1`` rounds of ``batch_isend_irecv(K,V)`` + ``R`` flash-attn launches + ``R`` LSE-merge launches
  (``ring/zigzag_ring_flash_attn.py:45-72``)  ->  the same kernel's compute CTAs walk ALL K/V
  segments (own shard first, then peers' in arrival order) with one online softmax carried in
  registers/TMEM; segments are gated by ``ld.acquire.sys`` arrival counters, so the transfer of
  shard ``j+1`` overlaps the tcgen05 math on shard ``j`` tile by tile.  NVSwitch makes every peer
  one hop away, so the "ring" is only a schedule, not a topology.
* output ``all_to_all_single`` + 2 copies (``hybrid/attn_layer.py:156-158``)  ->  the epilogue of every
  128-row O tile stores it into the token owner's output buffer over NVLink and bumps the owner's
  completion counter with ``red.release.sys``.

Memory: one symmetric slab per rank (cudaMalloc + CUDA IPC, handles exchanged once through the
process group's store), laid out as ``[signal pad | Q stage | K stage | V stage | out]``.
All counters are monotonic "epochs" -- nothing is reset between calls, so there is no barrier
kernel on the hot path; the only cross-rank handshake is a ready-to-receive flag written at the
top of each call.

Backward (``LCA_B200_FUSED_BWD=1``, default): dQ pass kernel = push CTAs (q, dO, k, v, delta) + tcgen05 dQ
tiles scattered to the token owners; dK/dV pass kernel reduces its partial tiles into the owners' fp32
accumulators with ``red.global.add.v4.f32`` over NVLink (replaces the R-hop fp32 dK/dV ring of
``ring_flash_attn.py:141-145``).  ``LCA_B200_FUSED_BWD=0`` falls back to the collective backward (NCCL
all-to-all + P2P ring around the same tcgen05 backward kernels).
"""
from __future__ import annotations

import os
import socket
from typing import List, Optional

import torch
import torch.distributed as dist

from ..ops import native
from ..ops.attention import AttnParams
from ..utils.logging import get_logger
from ..utils.profiling import nvtx_range
from .layout import Seg, canonical_variant, ring_positions, varlen_positions

_LOG = get_logger()

SIG_BYTES = 4096
SIG_KV, SIG_Q, SIG_RTR, SIG_ODONE, SIG_DKV, SIG_QA = 0, 16, 32, 48, 49, 64
MAX_PEERS = 16
_ALIGN = 1024


def _align(x: int) -> int:
    return (x + _ALIGN - 1) // _ALIGN * _ALIGN


class _SlabDoesNotFit(RuntimeError):
    pass


class _Slab:
    """This rank's symmetric slab + mapped views of every peer's slab (cudaMalloc + legacy CUDA IPC)."""

    def __init__(self, nbytes: int, device: torch.device):
        C = native.ext()
        self.nbytes = nbytes
        self.device = device
        self.ptr = C.symm_alloc(nbytes, device.index)       # raises on out-of-memory
        self.peer_ptrs = None
        self._me = -1

    def exchange(self, group):
        C = native.ext()
        handle = C.symm_export(self.ptr)
        world = dist.get_world_size(group)
        handles: List[Optional[bytes]] = [None] * world
        dist.all_gather_object(handles, handle, group=group)
        me = dist.get_rank(group)
        self.peer_ptrs = [self.ptr if i == me else C.symm_import(handles[i], self.device.index) for i in range(world)]
        self._me = me
        return self

    def tensor(self, offset: int, shape, dtype) -> torch.Tensor:
        return native.ext().symm_tensor(self.ptr + offset, list(shape), dtype, self.device.index)

    def free_local(self):
        native.ext().symm_free(self.ptr)

    def close(self):
        C = native.ext()
        for p in (self.peer_ptrs or []):
            if p != self.ptr:
                C.symm_unmap(p)
        C.symm_free(self.ptr)


class _VmmSlab:
    """Opt-in slab provider (``LCA_B200_SLAB=vmm``): ``torch.distributed._symmetric_memory`` (CUDA VMM API:
    ``cuMemCreate`` / ``cuMemMap`` / ``cuMemSetAccess``; only the slab itself is peer-visible, and the NVLS multicast
    address is available for the broadcast push).  Same interface as ``_Slab``; allocation and rendezvous are one
    collective step."""

    def __init__(self, nbytes: int, device: torch.device):
        import torch.distributed._symmetric_memory as symm
        self.nbytes = nbytes
        self.device = device
        self.buf = symm.empty(nbytes, dtype=torch.uint8, device=device)
        self.buf.zero_()
        torch.cuda.synchronize(device)

    def exchange(self, group):
        import torch.distributed._symmetric_memory as symm
        pg = group if group is not None else dist.group.WORLD
        self.hdl = symm.rendezvous(self.buf, pg)          # collective: also orders the zero-fill before any push
        self.ptr = int(self.buf.data_ptr())
        self.peer_ptrs = [int(x) for x in self.hdl.buffer_ptrs]
        me = dist.get_rank(pg)
        if self.peer_ptrs[me] != self.ptr:
            raise RuntimeError("symmetric memory handle does not describe the local buffer")
        try:
            self.multicast_ptr = int(self.hdl.multicast_ptr or 0)
        except Exception:   # noqa: BLE001 - no NVLS on this box / build
            self.multicast_ptr = 0
        dist.barrier(group=pg)
        return self

    def tensor(self, offset: int, shape, dtype) -> torch.Tensor:
        n = 1
        for d in shape:
            n *= int(d)
        nb = n * torch.empty((), dtype=dtype).element_size()
        return self.buf[offset:offset + nb].view(dtype).view(*shape)

    def free_local(self):
        self.buf = None

    def close(self):
        self.hdl = None
        self.buf = None


def _make_slab_local(nbytes: int, device: torch.device):
    kind = os.environ.get("LCA_B200_SLAB", "ipc")
    if kind == "vmm":
        return _VmmSlab(nbytes, device)
    if kind != "ipc":
        raise ValueError(f"LCA_B200_SLAB={kind!r}: expected 'ipc' or 'vmm'")
    return _Slab(nbytes, device)


def _make_slab(nbytes: int, group, device: torch.device):
    return _make_slab_local(nbytes, device).exchange(group)


class FusedUSPEngine:
    def __init__(self, sp_group, U: int, R: int, u: int, r: int, device: torch.device, ulysses_low: bool = True):
        self.group, self.U, self.R, self.u, self.r = sp_group, U, R, u, r
        self.P = U * R
        # LOGICAL sp index used by the kernels and every segment list: ring-major, ``d = r * U + u``.  With
        # ``use_ulysses_low=False`` (``globals.py:59-78`` of the reference: ring groups on the contiguous ranks) the
        # rank inside ``sp_group`` is ``u * R + r`` instead; only the peer-pointer tables are permuted (logical ->
        # group rank), so the kernels never see the difference.
        self.me = r * U + u
        self.ulysses_low = bool(ulysses_low)
        self.group_rank_of = [d if ulysses_low else (d % U) * R + d // U for d in range(self.P)]
        self.device = device
        self.slab: Optional[_Slab] = None
        self.sig: Optional[_Slab] = None
        self.key = None
        self.epoch = 0
        self.o_total = 0
        self.dkv_total = 0
        # push CTAs: one SM sustains ~70 GB/s of bulk copies into a peer (latency-bound: ~128 KiB in flight), so an NVLink
        # direction (~750 GB/s) needs >= 12; with the dynamic scheduler they join the compute pool afterwards (measured
        # N=2: 8 -> 16 CTAs: forward 2064 -> 2098 TFLOPS, Ulysses S=32K forward 5.8 -> 5.1 ms)
        self.n_comm = int(os.environ.get("LCA_B200_COMM_CTAS", "16"))
        self._plan = {}                         # call shape -> kv heads per launch (0 = refused), decided collectively, once
        self._segs = {}                         # (builder, args) -> marshalled segment lists of the current slab layout
        self.slab_bytes = 0
        self.with_bwd = os.environ.get("LCA_B200_FUSED_BWD", "1") == "1"
        # the signal pad lives in its own small slab so that growing the data slab never resets counters
        self.sig = self._logical(_make_slab(SIG_BYTES, sp_group, device))

    def _logical(self, slab):
        """Re-index a freshly exchanged slab's peer pointers by logical sp index (identity for ulysses-low meshes)."""
        slab.peer_ptrs = [slab.peer_ptrs[g] for g in self.group_rank_of]
        if slab.peer_ptrs[self.me] != slab.ptr:
            raise RuntimeError("fused engine: mesh coordinates do not match the rank inside the sp group")
        return slab

    def close(self) -> None:
        """Release the slabs (collective: peers must have stopped writing)."""
        torch.cuda.synchronize(self.device)
        try:
            dist.barrier(group=self.group)
        except Exception:   # noqa: BLE001 - the group may already be gone at interpreter exit
            pass
        for sl in (self.slab, self.sig):
            if sl is not None:
                sl.close()
        self.slab = self.sig = None
        self.key = None

    def supports_shapes(self, q, k) -> bool:
        """Shapes the push CTAs / kernels can handle; anything else takes the collective path."""
        rows, H, Hkv = q.shape[1], q.shape[2], k.shape[2]
        return (rows % 8 == 0 and k.shape[1] == rows and H % self.U == 0
                and (Hkv % self.U == 0 or self.U % Hkv == 0) and H % Hkv == 0)

    # ------------------------------------------------------------------------------ workspace
    def _layout(self, B, rows, H, Hkv, D, esz, bwd: bool):
        """Byte offsets of the staging tensors inside the slab -> (offsets dict, total bytes).  Forward-only calls
        (``torch.no_grad`` / inference) stage K/V of the whole sequence and Q of my ring block only; the
        owner-computes backward additionally stages Q, dO and the row statistics of every rank."""
        U = self.U
        Hl, Hkvl = H // U, (Hkv // U if Hkv >= U else 1)
        S = self.P * rows
        sq = B * (S if bwd else U * rows) * Hl * D * esz          # fwd: my ring block; bwd: every token
        skv = B * S * Hkvl * D * esz
        so = B * rows * H * D * esz
        sstat = B * Hl * S * 4
        sdkv = B * rows * Hkv * D * 4
        o = {}
        o["q"] = 0
        o["k"] = _align(o["q"] + sq)
        o["v"] = _align(o["k"] + skv)
        o["o"] = _align(o["v"] + skv)                    # out (fwd) / dq (bwd)
        o["lse_own"] = _align(o["o"] + so)               # (B, H, rows) fp32 LSE of my tokens (written by compute ranks)
        o["do"] = _align(o["lse_own"] + B * H * rows * 4)
        o["delta"] = _align(o["do"] + (sq if bwd else 0))
        o["lse2"] = _align(o["delta"] + (sstat if bwd else 0))
        o["dk"] = _align(o["lse2"] + (sstat if bwd else 0))
        o["dv"] = _align(o["dk"] + (sdkv if bwd else 0))
        total = _align(o["dv"] + (sdkv if bwd else 0)) + _ALIGN
        return o, total

    def staging_bytes(self, B, rows, H, Hkv, D, esz, bwd: bool) -> int:
        return self._layout(B, rows, H, Hkv, D, esz, bwd)[1]

    def _agree(self, ok: bool) -> bool:
        """Logical AND of ``ok`` over the sp group (every rank must take the same backend)."""
        t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
        return bool(int(t.item()))

    def _cap_bytes(self) -> int:
        """Upper bound for the data slab: ``LCA_B200_SLAB_MAX_GB``, else 60 % of the memory that is free right now
        (counting the slab we already hold)."""
        have = self.slab.nbytes if self.slab is not None else 0
        cap_gb = float(os.environ.get("LCA_B200_SLAB_MAX_GB", "0") or 0)
        if cap_gb > 0:
            return int(cap_gb * 2**30)
        return int(0.6 * (torch.cuda.mem_get_info(self.device)[0] + have))

    def reserve(self, q, k, need_bwd: bool) -> int:
        """Make sure the slab can hold this call (collective only for a call shape seen for the first time or when the
        slab has to grow).  Returns the number of KV HEADS PER FUSED LAUNCH: ``Hkv`` = the whole call in one launch;
        a smaller divisor = the call is executed as ``Hkv / c`` launches over head groups, so the staging is
        O(S * c * D) instead of O(S * Hkv * D) (bounded staging: the all-gather data model of the fused kernels keeps
        the K/V (and, for the backward, Q/dO) of every rank for the heads of ONE launch; looping over head groups is
        what keeps that under the cap at long sequence lengths); 0 -- on EVERY rank -- when not even the smallest head
        group fits or the slab cannot be allocated: the caller then uses the collective backend (O(S/P) memory).
        ``LCA_B200_HEAD_CHUNK=<kv heads>`` forces a group size (tests, memory-constrained jobs)."""
        B, rows, H, D = q.shape
        Hkv = k.shape[2]
        bwd = bool(need_bwd and self.with_bwd)
        key = (B, rows, H, Hkv, D, q.element_size(), bwd)
        if key in self._plan:
            return self._plan[key]
        c = self._plan_heads(*key)
        if c:
            try:
                self._ensure(B, rows, H // Hkv * c, c, D, q.element_size(), bwd)
            except _SlabDoesNotFit:
                c = 0
        self._plan[key] = c
        if 0 < c < Hkv:
            _LOG.info("fused engine %dx%d: q %s runs as %d launches of %d kv heads (slab %.2f GiB)", self.U, self.R,
                      tuple(q.shape), Hkv // c, c, self.slab_bytes / 2**30)
        return c

    def _plan_heads(self, B, rows, H, Hkv, D, esz, bwd: bool) -> int:
        cands = head_chunk_candidates(Hkv, self.U)
        forced = int(os.environ.get("LCA_B200_HEAD_CHUNK", "0") or 0)
        if forced in cands:
            cands = [forced]
        g = H // Hkv
        cap = self._cap_bytes()
        mine = next((c for c in cands if self._layout(B, rows, g * c, c, D, esz, bwd)[1] <= cap), 0)
        t = torch.tensor([mine], dtype=torch.int32, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)      # every rank must take the same plan
        return int(t.item())

    def _ensure(self, B, rows, H, Hkv, D, esz, bwd: bool = True):
        key = (B, rows, H, Hkv, D, esz, bwd)
        if key == self.key:
            return
        offs, total = self._layout(B, rows, H, Hkv, D, esz, bwd)
        if self.slab is None or self.slab.nbytes < total:
            cap = self._cap_bytes()
            if not self._agree(total <= cap):
                raise _SlabDoesNotFit(f"slab of {total / 2**30:.2f} GiB exceeds the cap ({cap / 2**30:.2f} GiB)")
            if self.slab is not None:
                torch.cuda.synchronize(self.device)
                dist.barrier(group=self.group)      # nobody may still be writing into the old slab
                self.slab.close()
                self.slab = None
            new = None
            try:
                new = _make_slab_local(total, self.device)
            except RuntimeError:
                torch.cuda.empty_cache()            # the caching allocator may be sitting on the memory we need
                try:
                    new = _make_slab_local(total, self.device)
                except RuntimeError:
                    new = None
            if not self._agree(new is not None):
                if new is not None:
                    new.free_local()
                raise _SlabDoesNotFit(f"could not allocate a {total / 2**30:.2f} GiB symmetric slab on every rank")
            self.slab = self._logical(new.exchange(self.group))
            _LOG.info("fused engine %dx%d: slab %.2f GiB (%s)", self.U, self.R, total / 2**30,
                      "fwd+bwd staging" if bwd else "fwd staging")
        self.off_q, self.off_k, self.off_v, self.off_o = offs["q"], offs["k"], offs["v"], offs["o"]
        self.off_lse_own, self.off_do, self.off_delta = offs["lse_own"], offs["do"], offs["delta"]
        self.off_lse2, self.off_dk, self.off_dv = offs["lse2"], offs["dk"], offs["dv"]
        self.key = key
        self._segs.clear()                      # segment lists embed peer pointers + offsets of the layout
        self.slab_bytes = self.slab.nbytes

    def dropout_seed(self) -> int:
        """One seed per module call, identical on every sp rank whatever their ``torch.manual_seed``: the first rank of
        the group draws it from torch's CPU generator (so a seeded run is reproducible and matches a single-device run
        with the same seed) and broadcasts it.  The owner-computes backward regenerates masks of queries whose forward
        ran on another rank -- with per-rank seeds dK/dV would silently be wrong."""
        seed = torch.randint(1, 2**31 - 1, (1,), dtype=torch.int64)
        t = seed.to(self.device)
        g = self.group if self.group is not None else dist.group.WORLD
        dist.broadcast(t, src=dist.get_global_rank(g, 0), group=g)
        return int(t.item())

    # ------------------------------------------------------------------------------ forward
    def forward(self, q, k, v, variant: str, p: AttnParams, need_bwd: bool = True, cu=None):
        """q (B, S/P, H, D), k/v (B, S/P, Hkv, D) local shards -> out (B, S/P, H, D), lse (B, H/U, S/R)."""
        C = native.ext()
        U, R, u, r, P = self.U, self.R, self.u, self.r, self.P
        B, rows, H, D = q.shape
        Hkv = k.shape[2]
        esz = q.element_size()
        if H % U:
            raise ValueError(f"query heads ({H}) must be divisible by the Ulysses degree ({U})")
        if not (Hkv % U == 0 or U % Hkv == 0):
            raise ValueError(f"kv heads ({Hkv}) must divide or be divisible by the Ulysses degree ({U})")
        Hl, Hkvl = H // U, (Hkv // U if Hkv >= U else 1)
        q, k, v = (_dense_heads(t) for t in (q, k, v))
        self._ensure(B, rows, H, Hkv, D, esz, bool(need_bwd and self.with_bwd))
        slab = self.slab
        Sr = U * rows                            # tokens per ring rank after the Ulysses gather
        self.epoch += 1
        push_q = 1 if U > 1 else 0
        kst = slab.tensor(self.off_k, (B, P * rows, Hkvl, D), q.dtype)
        vst = slab.tensor(self.off_v, (B, P * rows, Hkvl, D), q.dtype)
        if push_q:
            qst = slab.tensor(self.off_q, (B, Sr, Hl, D), q.dtype)
            out_local = slab.tensor(self.off_o, (B, rows, H, D), q.dtype)
        else:
            qst = q
            out_local = torch.empty((B, rows, H, D), dtype=q.dtype, device=q.device)
        lse = torch.empty((B, Hl, Sr), dtype=torch.float32, device=q.device)

        qsegs, n_my_tiles = self._q_segments(variant, rows, push_q, self.off_o, self.off_lse_own if push_q else None, cu)
        ksegs = self._k_segments(variant, rows, cu)
        qstride = R if canonical_variant(variant) == "stripe" else 1
        wl, wr = native.window_bounds(p)
        alibi = self._alibi(p, Hl)
        if push_q:
            self.o_total += U * B * Hl * n_my_tiles
            o_target = self.o_total & 0xFFFFFFFF
        else:
            o_target = 0
        self._arm_dropout(p, Hl)
        kvm, qm = self._push_masks(variant, rows, p, cu, False)
        C.usp_fwd(qst, kst, vst, q, k, v, qsegs, ksegs, qstride, qstride, out_local, u * Hl, lse,
                  float(p.softmax_scale), wl, wr, float(p.softcap), alibi,
                  [P, U, R, u, r, rows, self.n_comm, kvm, qm],
                  [self.off_q, self.off_k, self.off_v, Sr, P * rows],
                  slab.peer_ptrs, self.sig.peer_ptrs, self.sig.ptr, self.epoch, o_target)
        if push_q:   # the symmetric buffers are reused by the next call
            out = out_local.clone()
            lse_own = slab.tensor(self.off_lse_own, (B, H, rows), torch.float32).clone()
        else:
            out, lse_own = out_local, lse            # U == 1: (B, Hl, Sr) is already (B, H, rows)
        return out, lse, lse_own

    # ------------------------------------------------------------------------------ segment builders
    def _alibi(self, p, Hl):
        alibi = p.alibi_slopes
        if alibi is None:
            return None
        return alibi.to(device=self.device, dtype=torch.float32)[..., self.u * Hl:(self.u + 1) * Hl].contiguous()

    def _pos_of(self, variant, rows, cu=None):
        """ring rank -> position segments of that rank's gathered block (dense layout, or packed varlen sequences:
        ``cu`` = cumulative LOCAL sequence lengths, one attention group per sequence; ring-only meshes)."""
        Sr = self.U * rows
        if cu is None:
            return lambda rr: ring_positions(variant, rr, self.R, Sr)
        if self.U != 1:
            raise ValueError("packed variable-length batches are supported on ring-only meshes (U == 1)")
        if int(cu[-1]) != rows:
            raise ValueError(f"cu_seqlens ends at {int(cu[-1])} but the local shard has {rows} tokens")
        return lambda rr: varlen_positions(variant, rr, self.R, cu)

    def _push_masks(self, variant, rows, p: AttnParams, cu, backward: bool):
        """-> (kv_mask, q_mask): bit d set iff sp-rank d needs the DATA of my K/V rows / my Q-like rows.  Under a causal
        or sliding-window mask most (source, destination) pairs never meet (basic causal ring: rank d never reads keys
        of later ranks; window 8K at S=128K: only neighbours): the push CTAs still bump those destinations' arrival
        counters (the epochs stay aligned) but move no bytes.  Sound because the consumers' tile iterators visit a
        segment only if some tile of it is visible under the very same bounds (``native.window_bounds``), tested here
        on whole segments.  Destinations in my own ring block always get everything (their stationary operands)."""
        P, U, u, r = self.P, self.U, self.u, self.r
        every = (1 << P) - 1
        if not p.causal and tuple(p.window_size) == (-1, -1):
            return every, every
        key = ("m", variant, rows, cu, bool(p.causal), tuple(p.window_size), backward)
        hit = self._segs.get(key)
        if hit is not None:
            return hit
        pos_of = self._pos_of(variant, rows, cu)
        mine = tuple(sg for sg, _ in _slices_with_rows(pos_of(r), u * rows, (u + 1) * rows))
        wl, wr = native.window_bounds(p)           # the kernels' own bounds: visible iff -wl <= kpos - qpos <= wr

        def sees(qsegs, ksegs) -> bool:            # segment-pair granularity (a zigzag shard is two far-apart chunks)
            for a in qsegs:
                for b in ksegs:
                    if a.group != b.group or a.count == 0 or b.count == 0:
                        continue
                    qlo, qhi = a.start, a.start + (a.count - 1) * a.stride
                    klo, khi = b.start, b.start + (b.count - 1) * b.stride
                    if (wr >= 0 and klo - qhi > wr) or (wl >= 0 and qlo - khi > wl):
                        continue
                    return True
            return False

        kv_mask = q_mask = 0
        for d in range(P):
            dr = d // U
            blk = pos_of(dr)                       # queries (dQ pass / forward) and keys (dK/dV pass) of d's ring block
            if dr == r or sees(blk, mine):
                kv_mask |= 1 << d
            if not backward or dr == r or sees(mine, blk):
                q_mask |= 1 << d
        self._segs[key] = (kv_mask, q_mask)
        return kv_mask, q_mask

    def _cached(self, name, fn, *args):
        """Segment lists are pure functions of (mesh, layout, variant, rows, cu): built once per call shape, not per
        call (the marshalling was a visible part of the step at the 4K-32K sequence lengths of BASELINE config 2)."""
        key = (name,) + tuple(tuple(a) if isinstance(a, list) else a for a in args)
        hit = self._segs.get(key)
        if hit is None:
            hit = self._segs[key] = fn(*args)
        return hit

    def _q_segments(self, variant, rows, pushed: bool, off_out: int, off_lse=None, cu=None):
        return self._cached("q", self._q_segments_build, variant, rows, pushed, off_out, off_lse, cu)

    def _k_segments(self, variant, rows, cu=None):
        return self._cached("k", self._k_segments_build, variant, rows, cu)

    def _bwd_segments(self, variant: str, rows: int, cu=None):
        return self._cached("b", self._bwd_segments_build, variant, rows, cu)

    def _q_segments_build(self, variant, rows, pushed: bool, off_out: int, off_lse=None, cu=None):
        """Rows of my gathered Q (ring rank r) split by source shard -> kernel q segments
        [row0, nrows, pos0, flag, o_row0, o_base, o_sig, group]; also the number of 128-row tiles over MY rows."""
        U, R, u, r = self.U, self.R, self.u, self.r
        segs, n_my_tiles = [], 0
        pos = self._pos_of(variant, rows, cu)(r)
        for su in range(U):
            owner = r * U + su
            for s, row0 in _slices_with_rows(pos, su * rows, (su + 1) * rows):
                if pushed:
                    o_base = self.slab.peer_ptrs[owner] + off_out
                    o_sig = self.sig.peer_ptrs[owner] + 4 * SIG_ODONE
                    flag = SIG_Q + su
                else:
                    o_base, o_sig, flag = 0, 0, -1
                seg = [row0, s.count, s.start, flag, row0 - su * rows, o_base, o_sig, s.group]
                if off_lse is not None:
                    seg.append(self.slab.peer_ptrs[owner] + off_lse)
                segs.append(seg)
                if su == u:
                    n_my_tiles += (s.count + 127) // 128
        segs.sort(key=lambda x: -x[2])          # heaviest (latest positions) first
        return segs, n_my_tiles

    def _k_segments_build(self, variant, rows, cu=None):
        """All K/V rows in my staging (every source shard) -> [row0, nrows, pos0, flag, group], own block first."""
        U, R, u, r = self.U, self.R, self.u, self.r
        Sr = U * rows
        pos_of = self._pos_of(variant, rows, cu)
        segs = []
        for sr in [(r - i) % R for i in range(R)]:        # own ring block first, then "ring step" order
            pos = pos_of(sr)
            us = [u] + [x for x in range(U) if x != u] if sr == r else list(range(U))
            for su in us:
                for s, row0 in _slices_with_rows(pos, su * rows, (su + 1) * rows):
                    segs.append([sr * Sr + row0, s.count, s.start, SIG_KV + sr * U + su, s.group])
        return segs

    def _bwd_segments_build(self, variant: str, rows: int, cu=None):
        """Every token shard of the mesh as a segment list over the (B, S, ...) all-rank staging layout of the
        owner-computes backward -> (all_segs, mine, tiles_of_me); a segment is (src sp-rank, staging row0, nrows,
        global position of its first token, attention group).  ``mine`` = segments of my ring block (the stationary
        rows of both passes); ``tiles_of_me`` = 128-row tiles the whole mesh produces for tokens I own (completion
        count)."""
        U, R, u, r = self.U, self.R, self.u, self.r
        Sr = U * rows
        pos_of = self._pos_of(variant, rows, cu)
        all_segs = []
        for sr in [(r - i) % R for i in range(R)]:        # own ring block first, then "ring step" order
            pos = pos_of(sr)
            us = [u] + [x for x in range(U) if x != u] if sr == r else list(range(U))
            for su in us:
                for sg, row0 in _slices_with_rows(pos, su * rows, (su + 1) * rows):
                    all_segs.append((sr * U + su, sr * Sr + row0, sg.count, sg.start, sg.group))
        mine = [t for t in all_segs if t[0] // U == r]
        tiles_of_me = sum((n + 127) // 128 for (src, _, n, _, _) in all_segs if src == self.me)
        return all_segs, mine, tiles_of_me

    def too_many_segments(self, variant, rows, cu) -> bool:
        return len(self._k_segments(variant, rows, cu)) > native.MAX_SEG

    # ------------------------------------------------------------------------------ backward
    def backward(self, dout, q, k, v, out, lse, lse_own, variant: str, p: AttnParams, cu=None):
        """Owner-computes backward (default whenever ``Hkv % U == 0``).

        dQ pass kernel: push CTAs send q, dO (+ delta, lse2) head-slices to EVERY sp-rank and k, v likewise; compute
        CTAs produce dQ for my ring block's queries against all K/V and scatter dQ tiles to the token owners.
        dK/dV pass kernel: stationary = the K/V rows of MY ring block, streamed = every rank's Q/dO as it arrived, so
        dK/dV come out complete -- no cross-rank reduction, no fp32 traffic (the reference circulates fp32 dK/dV
        partials around the ring for R hops, ``ring_flash_attn.py:141-145``) -- and are scattered to the token owners
        in 16-bit like dQ."""
        U = self.U
        if k.shape[2] % U:
            return self.backward_reduce(dout, q, k, v, out, lse, variant, p, cu)
        C = native.ext()
        R, u, r, P = self.R, self.u, self.r, self.P
        B, rows, H, D = q.shape
        Hkv = k.shape[2]
        esz = q.element_size()
        Hl, Hkvl = H // U, Hkv // U
        q, k, v, dout = (_dense_heads(t) for t in (q, k, v, dout))
        self._ensure(B, rows, H, Hkv, D, esz, True)
        slab = self.slab
        Sr, S = U * rows, P * rows
        delta_local = native.attn_delta(out, dout)                                   # (B, H, rows)
        lse2_local = torch.where(torch.isinf(lse_own), torch.full_like(lse_own, float("inf")),
                                 lse_own * 1.4426950408889634).contiguous()
        q_all = slab.tensor(self.off_q, (B, S, Hl, D), q.dtype)
        do_all = slab.tensor(self.off_do, (B, S, Hl, D), q.dtype)
        kst = slab.tensor(self.off_k, (B, S, Hkvl, D), q.dtype)
        vst = slab.tensor(self.off_v, (B, S, Hkvl, D), q.dtype)
        delta_all = slab.tensor(self.off_delta, (B, Hl, S), torch.float32)
        lse2_all = slab.tensor(self.off_lse2, (B, Hl, S), torch.float32)
        dq_own = slab.tensor(self.off_o, (B, rows, H, D), q.dtype)
        dk_own = slab.tensor(self.off_dk, (B, rows, Hkv, D), q.dtype)
        dv_own = slab.tensor(self.off_dv, (B, rows, Hkv, D), q.dtype)
        self.epoch += 1
        fe = self.epoch * self.n_comm
        stride = R if canonical_variant(variant) == "stripe" else 1
        wl, wr = native.window_bounds(p)
        alibi = self._alibi(p, Hl)
        all_segs, mine, tiles_of_me = self._bwd_segments(variant, rows, cu)
        ksegs = [[row0, n, pos0, SIG_KV + src, grp] for (src, row0, n, pos0, grp) in all_segs]
        # ---- pass 1: dQ of my ring block's queries (stationary) against all K/V (streamed)
        xq = [[row0, n, pos0, grp, row0 - src * rows, SIG_QA + src, slab.peer_ptrs[src] + self.off_o, 0,
               self.sig.peer_ptrs[src] + 4 * SIG_ODONE] for (src, row0, n, pos0, grp) in sorted(mine, key=lambda t: -t[3])]
        self.o_total += U * B * Hl * tiles_of_me * 2
        self._arm_dropout(p, Hl)
        kvm, qm = self._push_masks(variant, rows, p, cu, True)
        C.usp_bwd_pass(False, q_all, do_all, kst, vst, xq, ksegs, stride, stride, lse2_all, delta_all, dq_own, None, 0,
                       u * Hl, float(p.softmax_scale), wl, wr, float(p.softcap), alibi, self.sig.ptr, fe,
                       [P, U, R, u, r, rows, self.n_comm, kvm, qm], [q, dout], [self.off_q, self.off_do], [k, v],
                       [self.off_k, self.off_v], [delta_local, lse2_local], [self.off_delta, self.off_lse2], True, Sr, S,
                       slab.peer_ptrs, self.sig.peer_ptrs, self.sig.ptr, self.epoch, self.o_total & 0xFFFFFFFF, H, Hkv)
        # ---- pass 2: dK/dV of my ring block's keys (stationary) against EVERY rank's queries (streamed)
        xk = [[row0, n, pos0, grp, row0 - src * rows, SIG_KV + src, slab.peer_ptrs[src] + self.off_dk,
               slab.peer_ptrs[src] + self.off_dv, self.sig.peer_ptrs[src] + 4 * SIG_DKV] for (src, row0, n, pos0, grp) in mine]
        yq = [[row0, n, pos0, SIG_QA + src, grp] for (src, row0, n, pos0, grp) in all_segs]
        self.dkv_total += U * B * Hkvl * tiles_of_me * 2
        self._arm_dropout(p, Hl)
        C.usp_bwd_pass(True, kst, vst, q_all, do_all, xk, yq, stride, stride, lse2_all, delta_all, dk_own, dv_own, 0,
                       u * Hkvl, float(p.softmax_scale), wr, wl, float(p.softcap), alibi, self.sig.ptr, fe, [], [], [], [],
                       [], [], [], False, Sr, S, slab.peer_ptrs, self.sig.peer_ptrs, self.sig.ptr, self.epoch, 0, H, Hkv)
        C.symm_wait(self.sig.ptr + 4 * SIG_DKV, self.dkv_total & 0xFFFFFFFF)
        return dq_own.clone(), dk_own.clone(), dv_own.clone()

    def backward_reduce(self, dout, q, k, v, out, lse, variant: str, p: AttnParams, cu=None):
        """Reduction backward (used when kv heads are replicated across Ulysses ranks, ``Hkv < U``): every compute rank
        forms partial dK/dV for all K/V rows it holds and reduces them into the owners' fp32 accumulators with
        ``red.global.add.v4.f32`` over NVLink."""
        C = native.ext()
        U, R, u, r, P = self.U, self.R, self.u, self.r, self.P
        B, rows, H, D = q.shape
        Hkv = k.shape[2]
        esz = q.element_size()
        Hl, Hkvl = H // U, (Hkv // U if Hkv >= U else 1)
        q, k, v, dout = (_dense_heads(t) for t in (q, k, v, dout))
        self._ensure(B, rows, H, Hkv, D, esz, True)
        slab = self.slab
        Sr = U * rows
        pushed = U > 1
        delta_local = native.attn_delta(out, dout)                                  # (B, H, rows) fp32
        lse2 = torch.where(torch.isinf(lse), torch.full_like(lse, float("inf")), lse * 1.4426950408889634)
        kst = slab.tensor(self.off_k, (B, P * rows, Hkvl, D), q.dtype)
        vst = slab.tensor(self.off_v, (B, P * rows, Hkvl, D), q.dtype)
        dk_acc = slab.tensor(self.off_dk, (B, rows, Hkv, D), torch.float32)
        dv_acc = slab.tensor(self.off_dv, (B, rows, Hkv, D), torch.float32)
        dk_acc.zero_()
        dv_acc.zero_()
        if pushed:
            qst = slab.tensor(self.off_q, (B, Sr, Hl, D), q.dtype)
            dost = slab.tensor(self.off_do, (B, Sr, Hl, D), q.dtype)
            delta_c = slab.tensor(self.off_delta, (B, Hl, Sr), torch.float32)
            dq_local = slab.tensor(self.off_o, (B, rows, H, D), q.dtype)
        else:
            qst, dost, delta_c = q, dout, delta_local
            dq_local = torch.empty((B, rows, H, D), dtype=q.dtype, device=q.device)
        self.epoch += 1
        fe = self.epoch * self.n_comm
        qsegs, n_my_tiles = self._q_segments(variant, rows, pushed, self.off_o, None, cu)
        ksegs = self._k_segments(variant, rows, cu)
        stride = R if canonical_variant(variant) == "stripe" else 1
        wl, wr = native.window_bounds(p)
        alibi = self._alibi(p, Hl)
        # ---- pass 1: dQ (+ all pushes)
        xq = [[s[0], s[1], s[2], s[7], s[4], s[3], s[5], 0, s[6]] for s in qsegs]
        if pushed:
            self.o_total += U * B * Hl * n_my_tiles * 2          # two warpgroups publish each tile
            o_target = self.o_total & 0xFFFFFFFF
        else:
            o_target = 0
        kvm, _ = self._push_masks(variant, rows, p, cu, False)      # Q-like tensors only go to my own ring block here
        mesh = [P, U, R, u, r, rows, self.n_comm, kvm, (1 << P) - 1]
        ql, qo = ([q, dout], [self.off_q, self.off_do]) if pushed else ([], [])
        self._arm_dropout(p, Hl)
        C.usp_bwd_pass(False, qst, dost, kst, vst, xq, ksegs, stride, stride, lse2, delta_c, dq_local, None, 0, u * Hl,
                       float(p.softmax_scale), wl, wr, float(p.softcap), alibi, self.sig.ptr, fe, mesh, ql, qo, [k, v],
                       [self.off_k, self.off_v], [delta_local] if pushed else [], [self.off_delta] if pushed else [],
                       False, Sr, P * rows, slab.peer_ptrs, self.sig.peer_ptrs, self.sig.ptr, self.epoch, o_target, H, Hkv)
        # ---- pass 2: dK/dV for every K/V row I hold, reduced into the owners' accumulators
        h0 = u * Hkvl if Hkv >= U else (u * Hkv) // U
        xk, n_my_kv_tiles = [], 0
        for s in ksegs:
            src = s[3] - SIG_KV
            o_row0 = s[0] - src * rows                          # row inside the owner's local shard
            xk.append([s[0], s[1], s[2], s[4], o_row0, s[3], slab.peer_ptrs[src] + self.off_dk,
                       slab.peer_ptrs[src] + self.off_dv, self.sig.peer_ptrs[src] + 4 * SIG_DKV])
            if src == self.me:
                n_my_kv_tiles += (s[1] + 127) // 128
        yq = [[s[0], s[1], s[2], s[3], s[7]] for s in qsegs]
        self.dkv_total += P * B * Hkvl * n_my_kv_tiles * 2
        self._arm_dropout(p, Hl)
        C.usp_bwd_pass(True, kst, vst, qst, dost, xk, yq, stride, stride, lse2, delta_c, dk_acc, dv_acc, 3, h0,
                       float(p.softmax_scale), wr, wl, float(p.softcap), alibi, self.sig.ptr, fe, [], [], [], [], [],
                       [], [], False, Sr, P * rows, slab.peer_ptrs, self.sig.peer_ptrs, self.sig.ptr, self.epoch, 0, H, Hkv)
        C.symm_wait(self.sig.ptr + 4 * SIG_DKV, self.dkv_total & 0xFFFFFFFF)
        dq = dq_local.clone() if pushed else dq_local
        return dq, dk_acc.to(k.dtype), dv_acc.to(v.dtype)

    # ------------------------------------------------------------------------------ autograd entry
    def attention(self, q, k, v, variant, softmax_scale, causal, window_size, softcap, alibi_slopes, deterministic,
                  dropout_p: float = 0.0, dropout_seed: int = 0, cu_seqlens=None, return_lse: bool = False,
                  kv_heads_per_launch: int = 0):
        from dataclasses import replace
        cu = None if cu_seqlens is None else tuple(int(x) for x in cu_seqlens)
        H, Hkv = q.shape[2], k.shape[2]
        c = kv_heads_per_launch if 0 < kv_heads_per_launch < Hkv else Hkv
        g = H // Hkv
        outs, lses = [], []
        for h0 in range(0, Hkv, c):          # one launch per head group (a single iteration unless the slab is capped)
            whole = c == Hkv
            qs = q if whole else q[:, :, h0 * g:(h0 + c) * g]
            ks, vs = (k, v) if whole else (k[:, :, h0:h0 + c], v[:, :, h0:h0 + c])
            al = alibi_slopes if (whole or alibi_slopes is None) else alibi_slopes[..., h0 * g:(h0 + c) * g]
            p = AttnParams.make(qs, softmax_scale, causal, window_size, softcap, al, dropout_p, deterministic)
            if p.dropout_p > 0.0:    # the mask needs no communication: it is a function of global coordinates
                p = replace(p, dropout_seed=int(dropout_seed), head_offset=h0 * g)
            o, l = _FusedAttnFunc.apply(qs, ks, vs, self, canonical_variant(variant), p, cu)
            outs.append(o)
            lses.append(l)
        out = outs[0] if len(outs) == 1 else torch.cat(outs, dim=2)
        if not return_lse:
            return out
        return out, (lses[0] if len(lses) == 1 else torch.cat(lses, dim=1))

    def _arm_dropout(self, p: AttnParams, Hl: int) -> None:
        """Hand the dropout key to the NEXT fused launch: local query head h of this rank is global head u*Hl + h."""
        if p.dropout_p > 0.0:
            from ..ops import dropout as _d
            native.ext().set_next_dropout([_d.p8_of(p.dropout_p), int(p.dropout_seed) & 0xFFFFFFFF,
                                           self.u * Hl + int(p.head_offset)])


def head_chunk_candidates(Hkv: int, U: int):
    """KV-head group sizes a call may be split into, largest first: divisors of ``Hkv`` that keep every group
    divisible by the Ulysses degree (replicated kv heads, ``Hkv < U``, are never split)."""
    if Hkv % U:
        return [Hkv]
    return [c for c in range(Hkv, 0, -1) if Hkv % c == 0 and c % U == 0]


def _dense_heads(t: torch.Tensor) -> torch.Tensor:
    ok = (t.stride(3) == 1 and t.stride(2) == t.shape[3] and t.stride(1) % 8 == 0 and t.stride(0) % 8 == 0
          and t.data_ptr() % 16 == 0)
    return t if ok else t.contiguous()


def _slices_with_rows(spec, begin: int, end: int):
    """Position segments of local rows [begin, end) together with their first row index."""
    out, off = [], 0
    for s in spec:
        lo, hi = max(begin, off), min(end, off + s.count)
        if hi > lo:
            out.append((Seg(s.start + (lo - off) * s.stride, hi - lo, s.stride, s.group), lo))
        off += s.count
    return out


class _SelfGroupType:
    """Sentinel process group of size 1 (see globals.group_size / group_rank)."""


_SelfGroup = _SelfGroupType()


class _FusedAttnFunc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, eng: FusedUSPEngine, variant: str, p: AttnParams, cu=None):
        out, lse, lse_own = eng.forward(q, k, v, variant, p, need_bwd=any(ctx.needs_input_grad[:3]), cu=cu)
        ctx.save_for_backward(q, k, v, out, lse, lse_own)
        ctx.eng, ctx.variant, ctx.p, ctx.cu = eng, variant, p, cu
        ctx.mark_non_differentiable(lse_own)
        return out, lse_own           # lse_own: (B, H, rows) fp32 LSE of MY tokens (all heads)

    @staticmethod
    def backward(ctx, dout, _dlse=None):
        """Owner-computes fused backward; ``LCA_B200_FUSED_BWD=0``: collective backward (NCCL a2a + ring P2P around
        the tcgen05 backward kernels)."""
        from .all_to_all import all_to_all_4D
        from .ring_attention import ring_attn_backward
        q, k, v, out, lse, lse_own = ctx.saved_tensors
        eng, p = ctx.eng, ctx.p
        if eng.with_bwd:
            dq, dk, dv = eng.backward(dout, q, k, v, out, lse, lse_own, ctx.variant, p, ctx.cu)
            return dq, dk, dv, None, None, None, None
        ug, rg = eng.ulysses_pg, eng.ring_pg
        if eng.U > 1 and k.shape[2] % eng.U:
            raise NotImplementedError("backward with kv_heads < ulysses degree needs the fused backward (round 2)")
        a2a = (lambda t: all_to_all_4D(t.contiguous(), 2, 1, group=ug)) if eng.U > 1 else (lambda t: t)
        alibi = p.alibi_slopes
        if alibi is not None and eng.U > 1:
            hl = q.shape[2] // eng.U
            alibi = alibi[..., eng.u * hl:(eng.u + 1) * hl].contiguous()
        from dataclasses import replace
        pl = replace(p, alibi_slopes=alibi, head_offset=eng.u * (q.shape[2] // eng.U) if eng.U > 1 else p.head_offset)
        if eng.R == 1:      # no ring dimension: `None` would mean the WORLD group to the ring loop
            rg = _SelfGroup
        dq, dk, dv = ring_attn_backward(rg, a2a(dout), a2a(q), a2a(k), a2a(v), a2a(out), lse, ctx.variant, pl, None, 0,
                                        ctx.cu, ctx.cu)
        back = (lambda t: all_to_all_4D(t.contiguous(), 1, 2, group=ug)) if eng.U > 1 else (lambda t: t)
        return back(dq), back(dk), back(dv), None, None, None, None


# ---------------------------------------------------------------------------------- factories
_ENGINES = {}


def invalidate_engines() -> None:
    """Drop every cached engine (called by ``set_seq_parallel_pg``: the process groups the engines were built on are
    being replaced, and ``id(group)`` of a new group may collide with a dead one).  Collective over each engine's group."""
    for key, eng in list(_ENGINES.items()):
        if eng is not None:
            eng.close()
    _ENGINES.clear()


def _same_node_p2p(group, device) -> bool:
    world = dist.get_world_size(group)
    info = [None] * world
    dist.all_gather_object(info, (socket.gethostname(), device.index), group=group)
    if len({h for h, _ in info}) != 1:
        return False
    C = native.ext()
    return all(d == device.index or C.can_access_peer(device.index, d) for _, d in info)


def _supported_input(q) -> bool:
    return (q.is_cuda and native.available() and q.dtype in (torch.bfloat16, torch.float16)
            and q.shape[-1] in native.SUPPORTED_HEAD_DIMS)   # padded head dims use the collective path


def engine_for_mesh(pgs, q, strict: bool = False):
    """Engine for the global U x R mesh set by set_seq_parallel_pg (collective on first call)."""
    mesh = pgs.mesh
    if mesh is None or mesh.sp_degree == 1 or pgs.SP_PG is None:
        return None
    if not _supported_input(q) or mesh.sp_degree > MAX_PEERS:
        if strict:
            raise RuntimeError("fused backend unavailable for this input/mesh")
        return None
    key = ("mesh", id(pgs.SP_PG), q.device.index)
    if key not in _ENGINES:
        if not _same_node_p2p(pgs.SP_PG, q.device):
            if strict:
                raise RuntimeError("fused backend needs all SP ranks on one node with P2P access")
            _ENGINES[key] = None
        else:
            eng = FusedUSPEngine(pgs.SP_PG, mesh.ulysses_degree, mesh.ring_degree, mesh.ulysses_rank, mesh.ring_rank,
                                 q.device, ulysses_low=mesh.use_ulysses_low)
            eng.ulysses_pg, eng.ring_pg = pgs.ULYSSES_PG, pgs.RING_PG
            _ENGINES[key] = eng
    return _ENGINES[key]


def engine_for_ring_group(group, q, strict: bool = False):
    """Engine for a bare ring group (``ring_flash_attn_*_func(group=...)``, the varlen entry points): U = 1."""
    if not (dist.is_available() and dist.is_initialized()):
        return None
    if type(group).__name__ == "_SelfGroupType":
        return None
    world = dist.get_world_size(group)
    if world == 1:
        return None
    if not _supported_input(q) or world > MAX_PEERS:
        if strict:
            raise RuntimeError("fused backend unavailable for this input/group")
        return None
    key = ("ring", id(group), q.device.index)
    if key not in _ENGINES:
        g = group if group is not None else dist.group.WORLD
        if not _same_node_p2p(g, q.device):
            if strict:
                raise RuntimeError("fused backend needs all ranks on one node with P2P access")
            _ENGINES[key] = None
        else:
            eng = FusedUSPEngine(g, 1, world, 0, dist.get_rank(g), q.device)
            eng.ulysses_pg, eng.ring_pg = None, g
            _ENGINES[key] = eng
    return _ENGINES[key]


def engine_for_ulysses_group(group, q, strict: bool = False):
    """Engine for a bare Ulysses group (UlyssesAttention(sequence_process_group=...))."""
    if not (dist.is_available() and dist.is_initialized()):
        return None
    world = dist.get_world_size(group)
    if world == 1:
        return None
    if not _supported_input(q) or world > MAX_PEERS:
        if strict:
            raise RuntimeError("fused backend unavailable for this input/group")
        return None
    key = ("ulysses", id(group), q.device.index)
    if key not in _ENGINES:
        g = group if group is not None else dist.group.WORLD
        if not _same_node_p2p(g, q.device):
            if strict:
                raise RuntimeError("fused backend needs all ranks on one node with P2P access")
            _ENGINES[key] = None
        else:
            eng = FusedUSPEngine(g, world, 1, dist.get_rank(g), 0, q.device)
            eng.ulysses_pg, eng.ring_pg = g, None
            _ENGINES[key] = eng
    return _ENGINES[key]
