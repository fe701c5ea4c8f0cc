# Excerpt-shaped fixture of megatron/core/parallel_state.py (anchor lines only).
import torch
from typing import Optional

from .utils import GlobalMemoryBuffer

# Intra-layer model parallel group that the current rank belongs to.
_TENSOR_MODEL_PARALLEL_GROUP = None
_SEQUENCE_PARALLEL_GROUP = None
_SEQUENCE_DATA_PARALLEL_GROUP = None
_SEQUENCE_PARALLEL_WORLD_SIZE = None


def initialize_model_parallel(
    tensor_model_parallel_size: int = 1,
    pipeline_model_parallel_size: int = 1,
    sequence_parallel_size: int = 1,
    virtual_pipeline_model_parallel_size: Optional[int] = None,
    pipeline_model_parallel_split_rank: Optional[int] = None,
    use_fp8: bool = False,
    use_distributed_optimizer: bool = False,
) -> None:
    """Initialize model data parallel groups."""
    world_size: int = torch.distributed.get_world_size()
    rank = torch.distributed.get_rank()
    global _SEQUENCE_PARALLEL_GROUP
    for i in range(world_size // sequence_parallel_size):
        ranks = range(i * sequence_parallel_size, (i + 1) * sequence_parallel_size)
        group = torch.distributed.new_group(ranks)
        if rank in ranks:
            _SEQUENCE_PARALLEL_GROUP = group

    # Build the sequence data parallel groups.
    global _SEQUENCE_DATA_PARALLEL_GROUP
    assert _SEQUENCE_DATA_PARALLEL_GROUP is None, \
        'sequence data parallel group is already initialized'


def get_sequence_parallel_world_size():
    """Return world size for the sequence parallel group."""
    global _SEQUENCE_PARALLEL_WORLD_SIZE
    return torch.distributed.get_world_size(group=_SEQUENCE_PARALLEL_GROUP)
