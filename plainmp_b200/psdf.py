"""Host mirror of ``plainmp.psdf`` (reference: cpp/plainmp/bindings/primitive.cpp:20-111,
cpp/plainmp/collision/primitive_sdf.hpp).

The objects keep the primitive parameters on the host in float64 exactly as the reference
classes do (pose, alignment flags, AABB) and are what ``SphereCollisionCst.set_sdf`` flattens
into the device scene table.  Every *evaluation* (``evaluate``, ``evaluate_batch``,
``is_outside``) runs in the CUDA library; without a GPU they raise.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import numpy as np

T_UNION, T_BOX, T_CYLINDER, T_SPHERE, T_GROUND, T_CLOUD = 0, 1, 2, 3, 4, 5


def _inverse3(a: np.ndarray) -> np.ndarray:
    """Eigen's fixed-size 3x3 inverse: cofactors / determinant (Eigen/src/LU/InverseImpl.h)."""

    def cof(i, j):
        i1, i2, j1, j2 = (i + 1) % 3, (i + 2) % 3, (j + 1) % 3, (j + 2) % 3
        return a[i1, j1] * a[i2, j2] - a[i1, j2] * a[i2, j1]

    c00, c10, c20 = cof(0, 0), cof(1, 0), cof(2, 0)
    invdet = 1.0 / (c00 * a[0, 0] + c10 * a[1, 0] + c20 * a[2, 0])
    r = np.empty((3, 3))
    for i in range(3):
        for j in range(3):
            r[j, i] = cof(i, j) * invdet
    return r


import itertools

_UIDS = itertools.count(1)  # see PrimitiveSDFBase.__init__

class Pose:
    """primitive_sdf.hpp:30-75; binding primitive.cpp:20-32."""

    def __init__(self, position: Optional[Sequence[float]] = None, rotation: Optional[np.ndarray] = None):
        self._position = np.zeros(3) if position is None else np.array(position, dtype=np.float64).reshape(3)
        self._rot = np.eye(3) if rotation is None else np.array(rotation, dtype=np.float64).reshape(3, 3)
        self._rot_inv = _inverse3(self._rot)
        self._update_alignment_flags()

    def _update_alignment_flags(self) -> None:  # :57-68
        diff = self._rot - np.eye(3)
        # Eigen isApprox(Identity), default precision 1e-12
        self._axis_aligned = bool(
            float((diff * diff).sum()) <= 1e-24 * min(float((self._rot * self._rot).sum()), 3.0)
        )
        if self._axis_aligned:
            self._z_axis_aligned = True
        else:
            tol = 1e-6
            r = self._rot
            self._z_axis_aligned = bool(
                abs(r[2, 2] - 1.0) < tol
                and abs(r[0, 2]) < tol
                and abs(r[1, 2]) < tol
                and abs(r[2, 0]) < tol
                and abs(r[2, 1]) < tol
            )

    def translate(self, translation: Sequence[float]) -> None:  # :41-43
        self._position = self._position + np.asarray(translation, dtype=np.float64).reshape(3)

    def rotate_z(self, angle: float) -> None:  # :45-53
        c, s = math.cos(angle), math.sin(angle)
        rot_z = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
        self._rot = rot_z @ self._rot
        self._rot_inv = _inverse3(self._rot)
        self._update_alignment_flags()

    def transform_points(self, p: np.ndarray) -> np.ndarray:  # :36-38, p is 3xN
        return self._rot @ p + self._position[:, None]

    @property
    def axis_aligned(self) -> bool:
        return self._axis_aligned

    @property
    def z_axis_aligned(self) -> bool:
        return self._z_axis_aligned

    @property
    def position(self) -> np.ndarray:
        return self._position.copy()

    @property
    def rotation(self) -> np.ndarray:
        return self._rot.copy()

    def _copy(self) -> "Pose":
        return Pose(self._position, self._rot)

    def __copy__(self) -> "Pose":
        return self._copy()

    def __deepcopy__(self, memo) -> "Pose":
        return self._copy()


class SDFBase:
    """primitive_sdf.hpp:79-97.  Evaluation goes through a cached device scene."""

    def __init__(self):
        self._scene_cache = None

    def get_type(self) -> int:
        raise NotImplementedError

    def _leaves(self) -> List["PrimitiveSDFBase"]:
        raise NotImplementedError

    def _state_key(self):
        raise NotImplementedError

    def _scene(self):
        from . import tables

        key = self._state_key()
        if self._scene_cache is None or self._scene_cache[0] != key:
            self._scene_cache = (key, tables.SceneTable(self._leaves()))
        return self._scene_cache[1]

    # --- binding surface (primitive.cpp:41-111) ---
    def evaluate_batch(self, points: np.ndarray) -> np.ndarray:
        """``points`` is 3xN (the binding takes an Eigen::Matrix3Xd); returns N values."""
        p = np.asarray(points, dtype=np.float64)
        if p.ndim != 2 or p.shape[0] != 3:
            raise ValueError("points must be a 3xN array")
        return self._scene().evaluate(np.ascontiguousarray(p.T))

    def evaluate(self, p: Sequence[float]) -> float:
        return float(self._scene().evaluate(np.asarray(p, dtype=np.float64).reshape(1, 3))[0])

    def is_outside(self, p: Sequence[float], radius: float) -> bool:
        return bool(
            self._scene().is_outside(np.asarray(p, dtype=np.float64).reshape(1, 3), np.array([radius]))[0]
        )

    # additive batch form of is_outside
    def is_outside_batch(self, points: np.ndarray, radii: np.ndarray) -> np.ndarray:
        p = np.asarray(points, dtype=np.float64)
        if p.ndim != 2 or p.shape[0] != 3:
            raise ValueError("points must be a 3xN array")
        return self._scene().is_outside(np.ascontiguousarray(p.T), np.asarray(radii, dtype=np.float64))


class UnionSDF(SDFBase):
    """primitive_sdf.hpp:99-173."""

    def __init__(self, sdfs: Sequence[SDFBase]):
        super().__init__()
        self.sdfs: List[SDFBase] = list(sdfs)

    def get_type(self) -> int:
        return T_UNION

    def clone(self) -> "UnionSDF":
        return UnionSDF([s.clone() for s in self.sdfs])

    def merge(self, other: "UnionSDF", clone: bool = False) -> None:
        for s in other.sdfs:
            self.sdfs.append(s.clone() if clone else s)

    def add(self, sdf: SDFBase, clone: bool = False) -> None:
        self.sdfs.append(sdf.clone() if clone else sdf)

    def rotate_z(self, angle: float) -> None:
        for s in self.sdfs:
            s.rotate_z(angle)

    def translate(self, translation: Sequence[float]) -> None:
        for s in self.sdfs:
            s.translate(translation)

    def _leaves(self) -> List["PrimitiveSDFBase"]:
        # primitive_sphere_collision.cpp:436-448 (recursive expansion, in order)
        out: List[PrimitiveSDFBase] = []
        for s in self.sdfs:
            out.extend(s._leaves())
        return out

    def _state_key(self):
        return tuple(s._state_key() for s in self.sdfs)


class PrimitiveSDFBase(SDFBase):
    """primitive_sdf.hpp:175-223."""

    def __init__(self):
        super().__init__()
        self._lb = np.zeros(3)
        self._ub = np.zeros(3)
        self._version = 0
        # a process-unique id: CPython reuses addresses of freed objects, so id(self) cannot key a cache
        # (a new SDF could inherit the id of a dead one and be mistaken for it, unmoved)
        self._uid = next(_UIDS)

    @property
    def lb(self) -> np.ndarray:
        return self._lb.copy()

    @property
    def ub(self) -> np.ndarray:
        return self._ub.copy()

    def _leaves(self) -> List["PrimitiveSDFBase"]:
        return [self]

    def _state_key(self):
        return (self._uid, self._version)


class GroundSDF(PrimitiveSDFBase):
    """primitive_sdf.hpp:225-259."""

    def __init__(self, height: float):
        super().__init__()
        self._height = float(height)
        # :250-255 — the AABB is pinned at z = 0 whatever the height
        self._lb = np.array([-np.inf, -np.inf, 0.0])
        self._ub = np.array([np.inf, np.inf, 0.0])

    def get_type(self) -> int:
        return T_GROUND

    @property
    def height(self) -> float:
        return self._height

    def clone(self) -> "GroundSDF":
        return GroundSDF(self._height)

    def rotate_z(self, angle: float) -> None:
        raise RuntimeError("GroundSDF does not support rotation")

    def translate(self, translation) -> None:
        raise RuntimeError("GroundSDF does not support translation")


class TransformableSDFBase(PrimitiveSDFBase):
    """primitive_sdf.hpp:261-276."""

    def __init__(self, pose: Pose):
        super().__init__()
        self._pose = pose._copy()

    @property
    def pose(self) -> Pose:
        return self._pose._copy()

    def rotate_z(self, angle: float) -> None:
        self._pose.rotate_z(angle)
        self._update_aabb()
        self._version += 1

    def translate(self, translation) -> None:
        self._pose.translate(translation)
        self._update_aabb()
        self._version += 1

    def _aabb_from_local_box(self, hx: float, hy: float, hz: float) -> None:  # :400-421, :500-521
        v = np.array(
            [[sx * hx, sy * hy, sz * hz] for sz in (-1, 1) for sy in (-1, 1) for sx in (-1, 1)], dtype=np.float64
        ).T
        w = self._pose.transform_points(v)
        self._lb = w.min(axis=1)
        self._ub = w.max(axis=1)

    def _update_aabb(self) -> None:
        raise NotImplementedError


class BoxSDF(TransformableSDFBase):
    """primitive_sdf.hpp:278-426."""

    def __init__(self, width: Sequence[float], pose: Optional[Pose] = None):
        super().__init__(pose if pose is not None else Pose())
        self._width = np.array(width, dtype=np.float64).reshape(3)
        self._update_aabb()

    def get_type(self) -> int:
        return T_BOX

    @property
    def width(self) -> np.ndarray:
        return self._width.copy()

    def clone(self) -> "BoxSDF":
        return BoxSDF(self._width, self._pose)

    def _update_aabb(self) -> None:
        self._aabb_from_local_box(self._width[0] * 0.5, self._width[1] * 0.5, self._width[2] * 0.5)


class CylinderSDF(TransformableSDFBase):
    """primitive_sdf.hpp:428-528."""

    def __init__(self, radius: float, height: float, pose: Optional[Pose] = None):
        super().__init__(pose if pose is not None else Pose())
        self._radius = float(radius)
        self._height = float(height)
        self._update_aabb()

    def get_type(self) -> int:
        return T_CYLINDER

    @property
    def radius(self) -> float:
        return self._radius

    @property
    def height(self) -> float:
        return self._height

    def clone(self) -> "CylinderSDF":
        return CylinderSDF(self._radius, self._height, self._pose)

    def _update_aabb(self) -> None:
        self._aabb_from_local_box(self._radius, self._radius, 0.5 * self._height)


class SphereSDF(TransformableSDFBase):
    """primitive_sdf.hpp:530-566."""

    def __init__(self, radius: float, pose: Optional[Pose] = None):
        super().__init__(pose if pose is not None else Pose())
        self._radius = float(radius)
        self._update_aabb()

    def get_type(self) -> int:
        return T_SPHERE

    @property
    def radius(self) -> float:
        return self._radius

    def clone(self) -> "SphereSDF":
        return SphereSDF(self._radius, self._pose)

    def _update_aabb(self) -> None:  # :558-561
        self._lb = self._pose._position - self._radius
        self._ub = self._pose._position + self._radius


class CloudSDF(PrimitiveSDFBase):
    """primitive_sdf.hpp:568-612.  ``points`` is N x 3."""

    def __init__(self, points: np.ndarray, radius: float):
        super().__init__()
        pts = np.asarray(points, dtype=np.float64)
        if pts.ndim != 2 or pts.shape[1] != 3 or pts.shape[0] == 0:
            raise ValueError("points must be a non-empty N x 3 array")
        self._points = np.ascontiguousarray(pts)
        self._radius = float(radius)
        self._lb = pts.min(axis=0) - self._radius  # :573-582
        self._ub = pts.max(axis=0) + self._radius

    def get_type(self) -> int:
        return T_CLOUD

    @property
    def points(self) -> np.ndarray:
        return self._points

    @property
    def radius(self) -> float:
        return self._radius

    def clone(self):
        raise RuntimeError("TODO: Not implemented yet")  # :593-595

    def rotate_z(self, angle: float) -> None:
        raise RuntimeError("TODO: Not implemented yet")  # :585-587

    def translate(self, translation) -> None:
        raise RuntimeError("TODO: Not implemented yet")  # :589-591
