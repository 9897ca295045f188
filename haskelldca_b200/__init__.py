"""haskelldca_b200 -- B200 (sm_100a) implementation of haskelldca's per-event hot path.

The product is libdca_b200.so (haskelldca_b200/csrc, C ABI in include/dca_b200.h).
This package is the thin Python host layer used by tests and bench.py; it mirrors
the reference's module API (Opt / Simulator / SimRunner / Gridfuncs / Gridneighs).
"""
from . import _ffi  # noqa: F401
from .opt import Opt, get_opts  # noqa: F401
from .simulator import Replicas, allreduce_stats, run_many  # noqa: F401
from . import agent, gridfuncs, gridneighs, sim_runner, stats  # noqa: F401

__all__ = ["Opt", "get_opts", "Replicas", "allreduce_stats", "run_many", "agent", "gridfuncs", "gridneighs", "sim_runner", "stats"]
