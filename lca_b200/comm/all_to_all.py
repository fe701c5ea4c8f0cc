"""``yunchang.comm.all_to_all`` module path (reference ``comm/all_to_all.py``)."""
from ..parallel.all_to_all import SeqAllToAll4D, SeqAllToAll5D, all_to_all_4D, all_to_all_5D  # noqa: F401
