"""Structured logging (the reference uses bare ``print``; ``test/test_hybrid_attn.py:94-117`` serialises
rank output with barriers)."""
from __future__ import annotations

import logging
import os

_FMT = "[%(asctime)s lca_b200 r%(rank)s %(levelname)s] %(message)s"


class _RankFilter(logging.Filter):
    def filter(self, record):
        record.rank = os.environ.get("RANK", "0")
        return True


def get_logger(name: str = "lca_b200") -> logging.Logger:
    log = logging.getLogger(name)
    if not log.handlers:
        h = logging.StreamHandler()
        h.setFormatter(logging.Formatter(_FMT, "%H:%M:%S"))
        h.addFilter(_RankFilter())
        log.addHandler(h)
        log.setLevel(os.environ.get("LCA_B200_LOGLEVEL", "WARNING").upper())
        log.propagate = False
    return log


def log_rank0(msg: str, level: int = logging.INFO) -> None:
    if os.environ.get("RANK", "0") == "0":
        get_logger().log(level, msg)
