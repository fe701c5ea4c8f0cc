from .megatron import DistributedAttention, initialize_sequence_parallel

__all__ = ["DistributedAttention", "initialize_sequence_parallel"]
