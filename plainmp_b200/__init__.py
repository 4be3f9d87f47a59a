"""plainmp_b200 — B200-native batched validity checking behind plainmp's object surface.

Submodules mirror the reference's Python modules for the hot path only:
``psdf``, ``kdtree``, ``kinematics``, ``constraint``, ``ompl`` (motion validators), ``robot_spec``;
``parallel`` shards item batches over the GPUs of one node.  All arithmetic runs in
``libplainmp_b200.so`` (hand-written sm_100a CUDA behind ``include/plainmp_b200.h``).
"""
from . import _capi  # noqa: F401

__version__ = "0.1.0"
