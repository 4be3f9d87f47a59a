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
        v, _ = self.check_motion_batch(np.asarray(s1, dtype=np.float64).reshape(1, -1), np.asarray(s2, dtype=np.float64).reshape(1, -1))
        return bool(v[0])

    def check_motion_batch(self, S1, S2, with_calls: bool = False):
        """(valid bool[n], number of waypoint validity checks executed[, is_valid calls per edge]).
        float64 inputs are taken as they are (the step ratio is computed from the doubles)."""
        return tables.check_edges(self._cst._robot(), self._cst._scene(), self._v, S1, S2, with_calls)


class ValidatorSession:
    """What ``CollisionAwareSpaceInformation`` is to the reference's planners (ompl_thin_wrap.hpp:60-125):
    ``is_valid(state)`` and ``check_motion(s1, s2)`` on FP64 states, backed by a ``pmp_session`` (own
    stream, pinned + device staging, one H2D / kernel / D2H per call — no torch tensors, no
    allocation).  Accepts one item or a batch; counts calls like ``is_valid_call_count_`` does.

    The session snapshots the constraint's flattened robot and scene tables at construction; call
    ``refresh()`` after mutating the scene or the non-controlled joints."""

    def __init__(self, ineq_cst, vconfig: ValidatorConfig | None = None, capacity: int = 4096, max_is_valid_call: int | None = None):
        if ineq_cst is None:
            raise RuntimeError("ineq_cst is nullptr. You should provide a constraint")  # ompl_thin_wrap.hpp:70-73
        import ctypes as C

        self._C = C
        self._cst = ineq_cst
        self._dim = ineq_cst.q_dim()
        self._lib = _capi.load()
        _capi.require_cuda()
        v = None
        self._keep = None
        if vconfig is not None:
            t = ValidatorType(vconfig.type)
            if t == ValidatorType.BOX:
                if len(vconfig.box_width) != self._dim:
                    raise RuntimeError("box dimension and space dimension mismatch")  # :84-86
                v, self._keep = tables.make_validator(int(t), self._dim, vconfig.box_width, 0.0)
            else:
                v, self._keep = tables.make_validator(int(t), self._dim, None, vconfig.resolution)
        self._robot = ineq_cst._robot()
        self._scene = ineq_cst._scene()
        h = C.c_void_p()
        _capi.check(
            self._lib.pmp_session_create(
                self._robot.handle, self._scene.handle if self._scene is not None else None,
                C.byref(v) if v is not None else None, int(capacity), C.byref(h),
            )
        )
        self._h = h
        self.is_valid_call_count = 0
        self.max_is_valid_call = max_is_valid_call

    def refresh(self) -> None:
        """Re-read the constraint's scene tables (after set_sdf / translate / rotate_z)."""
        self._scene = self._cst._scene()
        _capi.check(self._lib.pmp_session_set_scene(self._h, self._scene.handle if self._scene is not None else None))

    def reset_count(self) -> None:
        self.is_valid_call_count = 0

    def is_terminatable(self) -> bool:  # ompl_thin_wrap.hpp:110-112
        return self.max_is_valid_call is not None and self.is_valid_call_count > self.max_is_valid_call

    def is_valid(self, states):
        """bool for one state [dim], bool[n] for a batch [n, dim]."""
        q = np.ascontiguousarray(states, dtype=np.float64)
        single = q.ndim == 1
        q = q.reshape(-1, self._dim) if self._dim else q.reshape(1 if single else q.shape[0], 0)
        out = np.empty(q.shape[0], dtype=np.uint8)
        _capi.check(self._lib.pmp_session_is_valid(self._h, q.ctypes.data, q.shape[0], out.ctypes.data))
        self.is_valid_call_count += q.shape[0]
        return bool(out[0]) if single else out.astype(bool)

    def check_motion(self, s1, s2):
        """bool for one edge, bool[n] for a batch of edges."""
        a = np.ascontiguousarray(s1, dtype=np.float64)
        b = np.ascontiguousarray(s2, dtype=np.float64)
        if a.shape != b.shape:
            raise ValueError("s1 and s2 must have the same shape")
        single = a.ndim == 1
        a = a.reshape(-1, self._dim)
        b = b.reshape(-1, self._dim)
        out = np.empty(a.shape[0], dtype=np.uint8)
        n = self._C.c_uint64(0)
        _capi.check(self._lib.pmp_session_check_motion(self._h, a.ctypes.data, b.ctypes.data, a.shape[0], out.ctypes.data, self._C.byref(n)))
        self.is_valid_call_count += int(n.value)
        return bool(out[0]) if single else out.astype(bool)

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.pmp_session_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


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
