"""Host mirror of ``plainmp.kinematics`` (reference: cpp/plainmp/bindings/kinematics.cpp:30-50).

``KinematicModel`` holds the link/joint tables of a URDF exactly as the reference's
``KinematicModel<double>`` does (cpp/plainmp/kinematics/kinematics.cpp:29-223) and is the
source the flattener (``plainmp_b200.tables``) bakes the GPU link program from.  It does *no*
forward kinematics on the host: ``debug_get_link_pose`` goes through the CUDA library.
"""
from __future__ import annotations

import enum
import math
import xml.etree.ElementTree as ET
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


class BaseType(enum.IntEnum):
    """bindings/kinematics.cpp:46-49 (integer values are this repo's C-ABI codes)."""

    FIXED = 0
    PLANAR = 1
    FLOATING = 2


# joint kinds as stored per link (kind of the link's parent joint)
J_NONE, J_FIXED, J_REVOLUTE, J_CONTINUOUS, J_PRISMATIC = -1, 0, 1, 2, 3
_JTYPE = {"fixed": J_FIXED, "revolute": J_REVOLUTE, "continuous": J_CONTINUOUS, "prismatic": J_PRISMATIC}


def quat_from_rpy(roll: float, pitch: float, yaw: float, normalize: bool) -> np.ndarray:
    """(x, y, z, w).  urdf Rotation::setFromRPY normalises (third/urdf_parser/include/
    urdf_model/pose.h:202-216); QuatTrans::setQuaternionFromRPY does not
    (cpp/plainmp/kinematics/transform.hpp:122-130)."""
    phi, the, psi = roll / 2.0, pitch / 2.0, yaw / 2.0
    x = math.sin(phi) * math.cos(the) * math.cos(psi) - math.cos(phi) * math.sin(the) * math.sin(psi)
    y = math.cos(phi) * math.sin(the) * math.cos(psi) + math.sin(phi) * math.cos(the) * math.sin(psi)
    z = math.cos(phi) * math.cos(the) * math.sin(psi) - math.sin(phi) * math.sin(the) * math.cos(psi)
    w = math.cos(phi) * math.cos(the) * math.cos(psi) + math.sin(phi) * math.sin(the) * math.sin(psi)
    q = np.array([x, y, z, w], dtype=np.float64)
    if normalize:
        s = math.sqrt(float(q @ q))
        q = np.array([0.0, 0.0, 0.0, 1.0]) if s == 0.0 else q / s
    return q


def quat_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Eigen quaternion product, (x, y, z, w) storage."""
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array(
        [
            aw * bx + ax * bw + ay * bz - az * by,
            aw * by + ay * bw + az * bx - ax * bz,
            aw * bz + az * bw + ax * by - ay * bx,
            aw * bw - ax * bx - ay * by - az * bz,
        ]
    )


def quat_rotate(q: np.ndarray, v: np.ndarray) -> np.ndarray:
    """Eigen ``q * v`` (no normalisation assumed)."""
    qv = q[:3]
    uv = 2.0 * np.cross(qv, v)
    return v + q[3] * uv + np.cross(qv, uv)


@dataclass
class _Link:
    name: str
    parent: int  # link id, -1 for the root
    kind: int  # kind of the parent joint (J_*)
    origin_xyz: np.ndarray
    origin_quat: np.ndarray  # x, y, z, w
    axis: np.ndarray
    joint_id: int  # id of the movable parent joint, -1 otherwise
    consider_rotation: bool = True
    quat_identity_flag: bool = False  # tf_plink_to_hlink_cache_[.].is_quat_identity_ at init


class KinematicModel:
    """``KinematicModel(urdf_xml_string)`` — same constructor and methods as the binding."""

    def __init__(self, urdf_xml_string: str):
        if not urdf_xml_string:
            raise RuntimeError("xml string is empty")  # kinematics.cpp:186-188
        self._urdf_xml = urdf_xml_string  # kept so that the same text can be handed to another parser
        root = ET.fromstring(urdf_xml_string)
        link_names = [el.get("name") for el in root.findall("link")]
        joints = []
        for el in root.findall("joint"):
            jt = el.get("type")
            if jt not in _JTYPE:
                if jt in ("floating", "planar"):
                    raise RuntimeError(f"joint type {jt} is not supported")
                raise RuntimeError(f"unknown joint type {jt}")
            origin = el.find("origin")
            xyz = np.zeros(3)
            rpy = np.zeros(3)
            if origin is not None:
                if origin.get("xyz"):
                    xyz = np.array([float(t) for t in origin.get("xyz").split()])
                if origin.get("rpy"):
                    rpy = np.array([float(t) for t in origin.get("rpy").split()])
            axis_el = el.find("axis")
            axis = np.array([1.0, 0.0, 0.0])  # urdf default axis
            if axis_el is not None and axis_el.get("xyz"):
                axis = np.array([float(t) for t in axis_el.get("xyz").split()])
            lim = el.find("limit")
            lower = float(lim.get("lower", 0.0)) if lim is not None else 0.0
            upper = float(lim.get("upper", 0.0)) if lim is not None else 0.0
            joints.append(
                dict(
                    name=el.get("name"),
                    kind=_JTYPE[jt],
                    parent=el.find("parent").get("link"),
                    child=el.find("child").get("link"),
                    xyz=xyz,
                    quat=quat_from_rpy(*rpy, normalize=True),
                    axis=axis,
                    lower=lower,
                    upper=upper,
                )
            )
        children: Dict[str, List[dict]] = {n: [] for n in link_names}
        has_parent = set()
        for j in joints:
            if j["parent"] not in children or j["child"] not in children:
                raise RuntimeError(f"joint {j['name']} refers to an unknown link")
            children[j["parent"]].append(j)
            has_parent.add(j["child"])
        roots = [n for n in link_names if n not in has_parent]
        if len(roots) != 1:
            raise RuntimeError(f"URDF must have exactly one root link, found {roots}")

        # link / joint ids by depth-first traversal with an explicit stack
        # (kinematics.cpp:191-206, :66-135); the ids are not observable through results.
        self._links: List[_Link] = []
        self._link_ids: Dict[str, int] = {}
        self._joint_ids: Dict[str, int] = {}
        self._joint_child_link: List[int] = []
        self._joint_kind: List[int] = []
        self._joint_limits: List[Tuple[float, float]] = []
        stack = [(roots[0], None)]
        while stack:
            name, pj = stack.pop()
            lid = len(self._links)
            self._link_ids[name] = lid
            if pj is None:
                link = _Link(name, -1, J_NONE, np.zeros(3), np.array([0.0, 0.0, 0.0, 1.0]), np.zeros(3), -1)
            else:
                jid = -1
                if pj["kind"] != J_FIXED:
                    jid = len(self._joint_child_link)
                    self._joint_ids[pj["name"]] = jid
                    self._joint_child_link.append(lid)
                    self._joint_kind.append(pj["kind"])
                    if pj["kind"] == J_CONTINUOUS:
                        self._joint_limits.append((-math.inf, math.inf))
                    else:
                        self._joint_limits.append((pj["lower"], pj["upper"]))
                link = _Link(
                    name,
                    self._link_ids[pj["parent"]],
                    pj["kind"],
                    pj["xyz"],
                    pj["quat"],
                    pj["axis"],
                    jid,
                    True,
                    abs(pj["quat"][3] - 1.0) < 1e-6,  # kinematics.cpp:176-178
                )
            self._links.append(link)
            for cj in children[name]:
                stack.append((cj["child"], cj))
        n_j = len(self._joint_child_link)
        self._joint_angles = np.zeros(n_j)
        # 0 = never set (local transform is the URDF origin), 1 = exact sin/cos, 2 = approximate
        self._joint_set_mode = np.zeros(n_j, dtype=np.int8)
        self._base_pose = np.array([0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0])
        self._version = 0  # bumped on every state change; flattened tables key on it
        self._n_urdf_links = len(self._links)
        self._added_links: List[tuple] = []  # (parent_id, xyz, rpy, consider_rotation), in order

    # ---- table accessors used by the flattener and by the test-side oracle bridge ----
    @property
    def links(self) -> List[_Link]:
        return self._links

    @property
    def n_joints(self) -> int:
        return len(self._joint_child_link)

    @property
    def joint_child_links(self) -> List[int]:
        return self._joint_child_link

    @property
    def joint_kinds(self) -> List[int]:
        return self._joint_kind

    @property
    def all_joint_orientation_identity(self) -> bool:
        # kinematics.cpp:139-144
        return all(abs(self._links[c].origin_quat[3] - 1.0) < 1e-6 for c in self._joint_child_link)

    @property
    def version(self) -> int:
        return self._version

    def state_key(self, exclude_joint_ids: Sequence[int] = ()):
        """Hashable snapshot of everything a flattened table depends on, except the angles of
        ``exclude_joint_ids`` (the controlled joints, which are kernel inputs)."""
        mask = np.ones(self.n_joints, dtype=bool)
        mask[list(exclude_joint_ids)] = False
        return (
            len(self._links),
            self._base_pose.tobytes(),
            self._joint_angles[mask].tobytes(),
            self._joint_set_mode[mask].tobytes(),
        )

    # ---- binding surface ----
    def get_link_ids(self, link_names: Sequence[str]) -> List[int]:
        out = []
        for n in link_names:
            if n not in self._link_ids:
                raise ValueError("no link named " + n)  # std::invalid_argument, kinematics.cpp:270-272
            out.append(self._link_ids[n])
        return out

    def get_joint_ids(self, joint_names: Sequence[str]) -> List[int]:
        out = []
        for n in joint_names:
            if n not in self._joint_ids:
                raise ValueError("no joint named " + n)  # kinematics.cpp:243-245
            out.append(self._joint_ids[n])
        return out

    def get_joint_position_limits(self, joint_ids: Sequence[int]) -> List[Tuple[float, float]]:
        return [self._joint_limits[i] for i in joint_ids]

    def set_joint_positions(self, joint_ids: Sequence[int], positions: Sequence[float], accurate: bool = True) -> None:
        for i, p in zip(joint_ids, positions):
            self._joint_angles[i] = float(p)
            self._joint_set_mode[i] = 1 if accurate else 2
        self._version += 1

    def get_joint_positions(self, joint_ids: Sequence[int]) -> List[float]:
        return [float(self._joint_angles[i]) for i in joint_ids]

    def set_base_pose(self, pose_vec: Sequence[float]) -> None:
        v = np.asarray(pose_vec, dtype=np.float64)
        if v.shape != (7,):
            raise ValueError("base pose must be [x, y, z, qx, qy, qz, qw]")
        self._base_pose = v.copy()
        self._version += 1

    def get_base_pose(self) -> np.ndarray:
        return self._base_pose.copy()

    def add_new_link(
        self,
        link_name: str,
        parent_name: str,
        position: Sequence[float],
        rpy: Sequence[float],
        consider_rotation: bool,
    ) -> int:
        """kinematic_model_wrapper.cpp:31-39 -> kinematics.cpp:283-337."""
        parent_id = self.get_link_ids([parent_name])[0]
        if link_name in self._link_ids:
            raise RuntimeError("link name : " + link_name + " already exists")
        lid = len(self._links)
        self._link_ids[link_name] = lid
        self._links.append(
            _Link(
                link_name,
                parent_id,
                J_FIXED,
                np.asarray(position, dtype=np.float64).copy(),
                quat_from_rpy(*[float(a) for a in rpy], normalize=False),
                np.zeros(3),
                -1,
                bool(consider_rotation),
                False,  # a pushed-back Transform carries is_quat_identity_ = false
            )
        )
        self._added_links.append(
            (parent_id, np.asarray(position, dtype=np.float64).copy(), np.asarray(rpy, dtype=np.float64).copy(), bool(consider_rotation))
        )
        self._version += 1
        return lid

    def debug_get_link_pose(self, link_name: str) -> np.ndarray:
        """[x, y, z, qx, qy, qz, qw] of ``link_name`` for the model's current joint state
        (kinematic_model_wrapper.cpp:50-54).  Computed on the GPU (FP32)."""
        from . import tables  # late import: tables imports this module

        return tables.link_pose_via_cuda(self, link_name)

    # ---- local transforms as the reference caches them ----
    def local_transform(self, link_id: int, angle: Optional[float] = None, accurate: bool = True):
        """(quat xyzw, trans, is_quat_identity) of tf_plink_to_hlink_cache_[link_id] when the
        link's parent joint sits at ``angle`` (default: the stored joint angle), float64.
        algorithm.cpp:109-158 for movable joints, kinematics.cpp:147-181 for fixed ones."""
        L = self._links[link_id]
        if L.kind in (J_NONE, J_FIXED):
            return L.origin_quat.copy(), L.origin_xyz.copy(), L.quat_identity_flag
        mode = self._joint_set_mode[L.joint_id]
        if angle is None:
            if mode == 0:
                return L.origin_quat.copy(), L.origin_xyz.copy(), L.quat_identity_flag
            angle = float(self._joint_angles[L.joint_id])
            accurate = mode == 1
        if L.kind == J_PRISMATIC:
            return np.array([0.0, 0.0, 0.0, 1.0]), L.origin_xyz + L.axis * angle, True
        x = angle * 0.5
        if accurate:
            s, c = math.sin(x), math.cos(x)
        else:
            s, c = approx_sin_cos(x)
        jq = np.array([s * L.axis[0], s * L.axis[1], s * L.axis[2], c])
        if self.all_joint_orientation_identity or abs(L.origin_quat[3] - 1.0) < 1e-6:
            q = jq
        else:
            q = quat_mul(L.origin_quat, jq)
        return q, L.origin_xyz.copy(), False


def approx_sin_cos(x: float) -> Tuple[float, float]:
    """algorithm.cpp:49-86 in float64 (used only to bake non-controlled joints that were set
    with ``accurate=False``)."""
    half_pi = math.pi * 0.5
    sign = 1.0
    if x > half_pi or x < -half_pi:
        if x > math.pi or x < -math.pi:
            x = x - 2 * math.pi * math.floor(x * (1.0 / (2 * math.pi)) + 0.5)
        if x < -half_pi:
            x = -x - math.pi
            sign = -1.0
        elif x > half_pi:
            x = -x + math.pi
            sign = -1.0
    xx = x * x
    x4 = xx * xx
    x6 = x4 * xx
    x8 = x4 * x4
    s = x * (1 - xx / 6.0 + x4 / 120.0 - x6 / 5040.0 + x8 / 362880.0)
    c = sign * (1 - xx / 2.0 + x4 / 24.0 - x6 / 720.0 + x8 / 40320.0)
    return s, c
