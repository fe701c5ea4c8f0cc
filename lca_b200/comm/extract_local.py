"""``yunchang.comm.extract_local`` module path (reference ``comm/extract_local.py``)."""
from ..parallel.layout import (EXTRACT_FUNC_DICT, basic_extract_local, stripe_extract_local,  # noqa: F401
                               zigzag_extract_local)
