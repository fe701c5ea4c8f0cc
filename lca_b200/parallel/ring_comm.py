"""Ring point-to-point communicator (collective path).

Parity: ``yunchang/ring/utils.py:118-161`` (``RingComm``: send_recv / commit / wait).  Same
contract -- ``commit`` posts one batched isend/irecv group (deadlock-free because every rank
posts its send and its receive in the same NCCL group), ``wait`` blocks the current stream.
Works on NCCL and gloo.  With a degenerate group (size 1) it is a no-op passthrough.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist

from ..globals import group_rank, group_size


class RingComm:
    def __init__(self, process_group=None):
        self._pg = process_group
        self._ops: List[dist.P2POp] = []
        self._reqs = None
        self.rank = group_rank(process_group)
        self.world_size = group_size(process_group)
        self.send_rank = (self.rank + 1) % self.world_size
        self.recv_rank = (self.rank - 1) % self.world_size
        if process_group is not None and self.world_size > 1:
            self.send_rank = dist.get_global_rank(process_group, self.send_rank)
            self.recv_rank = dist.get_global_rank(process_group, self.recv_rank)

    def send_recv(self, to_send: torch.Tensor, recv_tensor: Optional[torch.Tensor] = None) -> torch.Tensor:
        res = torch.empty_like(to_send) if recv_tensor is None else recv_tensor
        if self.world_size == 1:
            res.copy_(to_send)
            return res
        self._ops.append(dist.P2POp(dist.isend, to_send, self.send_rank, group=self._pg))
        self._ops.append(dist.P2POp(dist.irecv, res, self.recv_rank, group=self._pg))
        return res

    def commit(self):
        if self._reqs is not None:
            raise RuntimeError("commit called twice")
        self._reqs = dist.batch_isend_irecv(self._ops) if self._ops else []

    def wait(self):
        if self._reqs is None:
            raise RuntimeError("wait called before commit")
        for r in self._reqs:
            r.wait()
        self._reqs = None
        self._ops = []
