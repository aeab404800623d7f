"""elementary_b200 — B200-native implementation of Elementary's block-rate audio graph evaluator.

Scope: ``elem::Runtime<float>::process()`` and the builtin node kernels it walks, re-cast so
that thousands of independent voices render in one fused sm_100a kernel per block, behind the reference's own
Runtime API (:class:`Runtime`) and unchanged instruction stream (:mod:`elementary_b200.el` emits it for tests).
"""
from . import el, graphs
from .runtime import Runtime, device_count, describe_return_code, load_library, RETURN_CODES

__all__ = ["Runtime", "el", "graphs", "device_count", "describe_return_code", "load_library", "RETURN_CODES"]
