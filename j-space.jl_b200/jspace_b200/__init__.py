"""jspace_b200 — host-side mirror of J-SPACE's spatial clonal-dynamics API over libjspace_b200.so.

Importing this package loads the CUDA library (no CPU fallback).
"""
from . import _lib
from ._lib import JSpaceError
from .api import PhiloxSeed, SpatialGraph, simulate_evolution, spatial_graph
from .distributed import shard_range
from .ensemble import EnsembleResult, Graph, Point, run_ensemble
from .start import Start_J_Space

_lib.lib()  # fail loudly at import time when the library has not been built

__all__ = ["Graph", "Point", "run_ensemble", "EnsembleResult", "JSpaceError", "PhiloxSeed", "SpatialGraph",
           "simulate_evolution", "spatial_graph", "Start_J_Space", "shard_range"]
