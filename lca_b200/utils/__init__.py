from .profiling import CudaTimer, nvtx_range, profile_to_tensorboard
from .logging import get_logger, log_rank0

__all__ = ["CudaTimer", "nvtx_range", "profile_to_tensorboard", "get_logger", "log_rank0"]
