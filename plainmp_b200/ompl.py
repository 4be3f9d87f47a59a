"""Host mirror of the motion-validator part of ``plainmp.ompl`` (reference:
cpp/plainmp/ompl/motion_validator.{hpp,cpp}, cpp/plainmp/ompl/ompl_thin_wrap.hpp:50-59,80-91,
binding cpp/plainmp/bindings/ompl_wrapper.cpp:26-40).

The reference installs a ``BoxMotionValidator`` / ``EuclideanMotionValidator`` inside OMPL's
SpaceInformation and calls ``checkMotion(s1, s2)`` one edge at a time.  ``MotionValidator`` here
keeps that call and adds ``check_motion_batch``; waypoints are expanded on the device with
per-edge early termination.  The planners themselves are out of scope.
"""
from __future__ import annotations

import enum
from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np

from . import _capi, tables


class ValidatorType(enum.IntEnum):
    BOX = _capi.VALIDATOR_BOX
    EUCLIDEAN = _capi.VALIDATOR_EUCLIDEAN


@dataclass
class ValidatorConfig:
    """ompl_thin_wrap.hpp:50-59."""

    type: ValidatorType = ValidatorType.BOX
    resolution: float = 0.0
    box_width: List[float] = field(default_factory=list)


class MotionValidator:
    """``checkMotion`` for a collision constraint (motion_validator.cpp:17-79)."""

    def __init__(self, ineq_cst, vconfig: ValidatorConfig):
        if ineq_cst is None:
            raise RuntimeError("ineq_cst is nullptr. You should provide a constraint")  # ompl_thin_wrap.hpp:70-73
        self._cst = ineq_cst
        self._dim = ineq_cst.q_dim()
        self._type = ValidatorType(vconfig.type)
        if self._type == ValidatorType.BOX:
            if len(vconfig.box_width) != self._dim:
                raise RuntimeError("box dimension and space dimension mismatch")  # :84-86
            self._v, self._keep = tables.make_validator(int(self._type), self._dim, vconfig.box_width, 0.0)
        else:
            self._v, self._keep = tables.make_validator(int(self._type), self._dim, None, vconfig.resolution)

    def check_motion(self, s1: Sequence[float], s2: Sequence[float]) -> bool:
        v, _ = self.check_motion_batch(np.asarray(s1).reshape(1, -1), np.asarray(s2).reshape(1, -1))
        return bool(v[0])

    def check_motion_batch(self, S1, S2):
        """(valid bool[n], number of waypoint validity checks executed)."""
        return tables.check_edges(self._cst._robot(), self._cst._scene(), self._v, S1, S2)


def step_ratio(vconfig: ValidatorConfig, s1: np.ndarray, s2: np.ndarray) -> float:
    """determine_step_ratio on the host, float64 (motion_validator.cpp:89-126); used by callers
    that need the waypoint count (n_test = floor(1/ratio) + 2) without launching anything."""
    s1 = np.asarray(s1, dtype=np.float64)
    s2 = np.asarray(s2, dtype=np.float64)
    if ValidatorType(vconfig.type) == ValidatorType.BOX:
        w = np.asarray(vconfig.box_width, dtype=np.float64)
        diff = s2 - s1
        k = int(np.argmax(np.abs(diff) * (1.0 / w)))  # first maximum wins, like the strict > loop
        if abs(diff[k]) < 1e-6:
            return float("inf")
        return float(w[k] / abs(diff[k]))
    dist = float(np.sqrt(((s1 - s2) ** 2).sum()))
    return float("inf") if dist == 0.0 else vconfig.resolution / dist
