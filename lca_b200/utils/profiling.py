"""Tracing / timing helpers (the reference only has ``--use_profiler`` in its benchmark,
``benchmark/benchmark_longctx.py:82-100``; NVTX ranges and device timers are new)."""
from __future__ import annotations

import contextlib
import time
from typing import Optional

import torch


@contextlib.contextmanager
def nvtx_range(name: str):
    """NVTX range visible in Nsight Systems/Compute timelines; no-op without CUDA."""
    on = torch.cuda.is_available()
    if on:
        torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        if on:
            torch.cuda.nvtx.range_pop()


class CudaTimer:
    """Device-side timer (CUDA events on the current stream); falls back to wall clock on CPU.

        with CudaTimer() as t: step()
        t.ms  # after exit; t.max_over_ranks(group) for the multi-GPU figure
    """

    def __init__(self, device: Optional[torch.device] = None):
        self.cuda = torch.cuda.is_available() and (device is None or device.type == "cuda")
        self.ms = float("nan")

    def __enter__(self):
        if self.cuda:
            self._e0, self._e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            self._e0.record()
        else:
            self._t0 = time.perf_counter()
        return self

    def __exit__(self, *exc):
        if self.cuda:
            self._e1.record()
            torch.cuda.synchronize()
            self.ms = self._e0.elapsed_time(self._e1)
        else:
            self.ms = (time.perf_counter() - self._t0) * 1e3
        return False

    def max_over_ranks(self, group=None) -> float:
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            return self.ms
        t = torch.tensor([self.ms], dtype=torch.float64, device="cuda" if self.cuda else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        return float(t)


@contextlib.contextmanager
def profile_to_tensorboard(logdir: str = "./profile/", wait: int = 0, warmup: int = 2, active: int = 4):
    """torch.profiler with the reference's schedule (wait=0, warmup=2, active=4, tensorboard handler)."""
    acts = [torch.profiler.ProfilerActivity.CPU]
    if torch.cuda.is_available():
        acts.append(torch.profiler.ProfilerActivity.CUDA)
    with torch.profiler.profile(activities=acts, schedule=torch.profiler.schedule(wait=wait, warmup=warmup, active=active, repeat=1),
                                on_trace_ready=torch.profiler.tensorboard_trace_handler(logdir), record_shapes=True,
                                with_stack=True) as prof:
        yield prof
