"""``yunchang.ring.triton_utils`` module path (reference ``ring/triton_utils.py``: Triton LSE flatten/unflatten).
No Triton here: on a GPU the CUDA kernels in ``csrc/util_kernels.cu`` do the work, on CPU plain indexing."""
from .utils import flatten_varlen_lse, unflatten_varlen_lse  # noqa: F401


def _not_a_kernel(*args, **kwargs):
    raise RuntimeError("flatten_kernel / unflatten_kernel are Triton kernels in the reference; call "
                       "flatten_varlen_lse / unflatten_varlen_lse instead")


flatten_kernel = unflatten_kernel = _not_a_kernel
