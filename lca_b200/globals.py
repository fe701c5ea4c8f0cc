"""Process-global sequence-parallel state: the U x R mesh and capability flags.

Parity target: ``yunchang/globals.py:5-135`` (reference).  Differences, by design:

* the mesh arithmetic lives in :mod:`lca_b200.parallel.mesh` as pure functions that are
  unit-testable without ``torch.distributed``;
* divisibility is validated *before* any group is built and with an accurate message
  (reference defect: ``globals.py:29-34``);
* capability probes never touch the CUDA driver at import time (reference defect:
  ``globals.py:99-110`` dies with ``RuntimeError`` on a CUDA-less host);
* a process that never initialised ``torch.distributed`` gets a degenerate 1x1 mesh so the
  whole API is usable single-GPU / single-process.
"""
from __future__ import annotations

import importlib.util
from typing import Optional

import torch.distributed as dist

from .parallel.mesh import MeshSpec, build_mesh_spec


class _ProcessGroupState:
    """Singleton container, attribute-compatible with the reference's ``PROCESS_GROUP``."""

    _instance: Optional["_ProcessGroupState"] = None

    def __new__(cls):
        if cls._instance is None:
            cls._instance = super().__new__(cls)
            cls._instance._reset()
        return cls._instance

    def _reset(self):
        self.ULYSSES_PG = None
        self.RING_PG = None
        self.DP_PG = None  # replicas of the SP mesh (bookkeeping; the reference has none)
        self.SP_PG = None  # all U*R ranks of this replica (used by the fused NVLink path)
        self.mesh: Optional[MeshSpec] = None
        self.initialized = False

    # convenience -----------------------------------------------------------------------
    @property
    def ulysses_degree(self) -> int:
        return self.mesh.ulysses_degree if self.mesh is not None else 1

    @property
    def ring_degree(self) -> int:
        return self.mesh.ring_degree if self.mesh is not None else 1

    @property
    def ulysses_rank(self) -> int:
        return self.mesh.ulysses_rank if self.mesh is not None else 0

    @property
    def ring_rank(self) -> int:
        return self.mesh.ring_rank if self.mesh is not None else 0


PROCESS_GROUP = _ProcessGroupState()

# names the reference exposes from ``yunchang.globals`` (``globals.py:5-24``); code that does
# ``ProcessGroupSingleton()`` gets the one shared state object, exactly like there.
ProcessGroupSingleton = _ProcessGroupState


class Singleton:
    """Base for process-wide singletons (``yunchang/globals.py:5-11``)."""

    _instance = None

    def __new__(cls, *args, **kwargs):
        if cls._instance is None:
            cls._instance = super().__new__(cls)
        return cls._instance


def set_seq_parallel_pg(
    sp_ulysses_degree: int,
    sp_ring_degree: int,
    rank: int,
    world_size: int,
    use_ulysses_low: bool = True,
) -> None:
    """Build the ``world/(U*R)`` replicas of the ``U x R`` sequence-parallel mesh.

    Signature-compatible with ``yunchang.set_seq_parallel_pg`` (``globals.py:22-81``).
    Every rank calls every ``new_group`` (NCCL/gloo requirement); the group lists come from
    :func:`lca_b200.parallel.mesh.build_mesh_spec`.
    """
    spec = build_mesh_spec(
        sp_ulysses_degree, sp_ring_degree, rank, world_size, use_ulysses_low
    )
    _invalidate_fused_engines()
    PROCESS_GROUP._reset()
    PROCESS_GROUP.mesh = spec

    if world_size == 1 or not (dist.is_available() and dist.is_initialized()):
        if world_size != 1:
            raise RuntimeError(
                "set_seq_parallel_pg(world_size>1) needs torch.distributed to be initialised"
            )
        PROCESS_GROUP.initialized = True
        return

    for ranks in spec.all_ulysses_groups:
        g = dist.new_group(list(ranks))
        if rank in ranks:
            PROCESS_GROUP.ULYSSES_PG = g
    for ranks in spec.all_ring_groups:
        g = dist.new_group(list(ranks))
        if rank in ranks:
            PROCESS_GROUP.RING_PG = g
    for ranks in spec.all_sp_groups:
        g = dist.new_group(list(ranks))
        if rank in ranks:
            PROCESS_GROUP.SP_PG = g
    if spec.dp_degree > 1:
        for ranks in spec.all_dp_groups:
            g = dist.new_group(list(ranks))
            if rank in ranks:
                PROCESS_GROUP.DP_PG = g
    PROCESS_GROUP.initialized = True


def _invalidate_fused_engines() -> None:
    """Engines of the fused NVLink backend are cached per process group; rebuilding the groups retires them."""
    import sys
    mod = sys.modules.get("lca_b200.parallel.fused_engine")
    if mod is not None:
        mod.invalidate_engines()


def _is_self_group(group) -> bool:
    return type(group).__name__ == "_SelfGroupType"


def group_size(group) -> int:
    if _is_self_group(group):
        return 1
    if group is None:
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size()
        return 1
    return dist.get_world_size(group)


def group_rank(group) -> int:
    if _is_self_group(group):
        return 0
    if group is None:
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank()
        return 0
    return dist.get_rank(group)


# ------------------------------------------------------------------------------------------
# Capability flags.  Same names as the reference (``globals.py:83-135``) so downstream
# ``if HAS_FLASH_ATTN`` code keeps working; probes are find_spec-only (no import side effects).
# ------------------------------------------------------------------------------------------
def _has(mod: str) -> bool:
    try:
        return importlib.util.find_spec(mod) is not None
    except (ImportError, ValueError):
        return False


HAS_FLASH_ATTN = _has("flash_attn")              # library FA2 (only used by the reference arm)
HAS_FLASH_ATTN_HOPPER = False                    # FA3 is sm_90a-only; never usable on B200
HAS_FLASHINFER = _has("flashinfer")
HAS_AITER = False                                # ROCm only
HAS_SAGE_ATTENTION = _has("sageattention")
HAS_SPARSE_SAGE_ATTENTION = _has("spas_sage_attn")
HAS_NPU = False                                  # Ascend only


def get_cuda_arch() -> str:
    """``"major.minor"`` of the current device (``yunchang/globals.py:102-104``); ``"10.0"`` (the only build target)
    when no GPU is visible.  Unlike the reference this never touches ``TORCH_CUDA_ARCH_LIST``."""
    import torch
    if torch.cuda.is_available():
        major, minor = torch.cuda.get_device_capability()
        return f"{major}.{minor}"
    return "10.0"


def has_native_kernels() -> bool:
    """True when the in-tree sm_100a extension is built and a Blackwell GPU is visible."""
    from .ops import native

    return native.available()
