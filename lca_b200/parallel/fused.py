"""Fused NVLink backend entry points (engine lives in fused_engine.py once symmetric memory is up)."""
from __future__ import annotations


def get_engine_if_supported(process_group_state, q, strict: bool = False):
    try:
        from .fused_engine import engine_for_mesh
    except ImportError:
        if strict:
            raise
        return None
    return engine_for_mesh(process_group_state, q, strict)


def get_ulysses_engine_if_supported(group, q, strict: bool = False):
    try:
        from .fused_engine import engine_for_ulysses_group
    except ImportError:
        if strict:
            raise
        return None
    return engine_for_ulysses_group(group, q, strict)
