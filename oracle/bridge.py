"""ctypes bridge to the FP64 CPU oracle (oracle/plainmp_oracle.cpp).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under plainmp_b200/ imports this module.

The oracle objects are built from the *host mirror* objects of plainmp_b200 (KinematicModel,
psdf.*, SphereCollisionCst) but from their un-flattened content: full link tree, un-reordered
sphere specs, the SDF tree.  The flattening done for the GPU (tables.py) is therefore checked by
the parity tests too.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path
from typing import List, Optional, Sequence, Tuple

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "libplainmp_oracle.so"


def build(force: bool = False) -> Path:
    src = HERE / "plainmp_oracle.cpp"
    if force or not LIB.exists() or LIB.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "-B" if force else "-s"], check=True)
    return LIB


_lib = None
_dp = C.POINTER(C.c_double)


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB.exists():
            build()
        L = C.CDLL(str(LIB))
        vp, i, d, i64 = C.c_void_p, C.c_int, C.c_double, C.c_int64
        sig = {
            "orc_last_error": (C.c_char_p, []),
            "orc_kin_create": (vp, [i, i, vp, vp, vp, i, vp, vp, vp]),
            "orc_kin_destroy": (None, [vp]),
            "orc_kin_set_joint_angles": (None, [vp, i, vp, vp, i]),
            "orc_kin_set_base_pose": (None, [vp, vp]),
            "orc_kin_get_link_pose": (None, [vp, i, vp]),
            "orc_kin_add_new_link": (i, [vp, i, vp, vp, i]),
            "orc_approx_sin_cos": (None, [d, _dp, _dp]),
            "orc_sdf_box": (vp, [vp, vp, vp]),
            "orc_sdf_cylinder": (vp, [d, d, vp, vp]),
            "orc_sdf_sphere": (vp, [d, vp, vp]),
            "orc_sdf_ground": (vp, [d]),
            "orc_sdf_cloud": (vp, [vp, i64, d]),
            "orc_sdf_union": (vp, [vp, i]),
            "orc_sdf_destroy": (None, [vp]),
            "orc_sdf_clone": (vp, [vp]),
            "orc_sdf_translate": (i, [vp, vp]),
            "orc_sdf_rotate_z": (i, [vp, d]),
            "orc_sdf_evaluate": (d, [vp, vp]),
            "orc_sdf_is_outside": (i, [vp, vp, d]),
            "orc_sdf_evaluate_batch": (None, [vp, vp, i64, vp]),
            "orc_sdf_is_outside_batch": (None, [vp, vp, vp, i64, vp]),
            "orc_sdf_aabb": (i, [vp, vp, vp]),
            "orc_sdf_pose": (i, [vp, vp, vp, vp]),
            "orc_kdtree_create": (vp, [vp, i64]),
            "orc_kdtree_destroy": (None, [vp]),
            "orc_kdtree_query": (None, [vp, vp, vp]),
            "orc_kdtree_sqdist": (d, [vp, vp]),
            "orc_kdtree_sqdist_batch": (None, [vp, vp, i64, vp]),
            "orc_cst_create": (vp, [vp, i, vp, i, i, vp, vp, vp, vp, i, vp, vp, i]),
            "orc_cst_destroy": (None, [vp]),
            "orc_cst_set_sdf": (None, [vp, vp]),
            "orc_cst_q_dim": (i, [vp]),
            "orc_cst_n_spheres": (i, [vp]),
            "orc_cst_n_groups": (i, [vp]),
            "orc_cst_n_prims": (i, [vp]),
            "orc_cst_is_valid": (i, [vp, vp]),
            "orc_cst_evaluate": (i, [vp, vp, vp]),
            "orc_cst_update_kintree": (None, [vp, vp, i]),
            "orc_cst_get_all_spheres": (None, [vp, vp]),
            "orc_cst_get_group_spheres": (None, [vp, vp]),
            "orc_cst_get_sphere_order": (None, [vp, vp]),
            "orc_cst_is_valid_batch": (None, [vp, vp, i64, i, vp]),
            "orc_cst_evaluate_batch": (None, [vp, vp, i64, i, vp]),
            "orc_cst_evaluate_grad_batch": (None, [vp, vp, i64, i, vp, vp]),
            "orc_cst_margin_batch": (None, [vp, vp, i64, i, vp]),
            "orc_cst_spheres_batch": (None, [vp, vp, i64, i, vp]),
            "orc_cst_work_counts": (i64, [vp, vp, i64, i, vp]),
            "orc_validator_create": (vp, [i, i, vp, d]),
            "orc_validator_destroy": (None, [vp]),
            "orc_validator_step_ratio": (d, [vp, vp, vp]),
            "orc_check_motion": (i, [vp, vp, vp, vp, C.POINTER(C.c_uint64)]),
            "orc_check_motion_batch": (None, [vp, vp, vp, vp, i64, i, vp, C.POINTER(C.c_uint64)]),
            "orc_check_motion_batch_ex": (None, [vp, vp, vp, vp, i64, i, vp, vp, vp]),
            "orc_composite_check": (i, [vp, i]),
            "orc_composite_is_valid_batch": (None, [vp, i, vp, i64, i, vp]),
            "orc_composite_margin_batch": (None, [vp, i, vp, i64, i, vp]),
            "orc_composite_evaluate_batch": (None, [vp, i, vp, i64, i, vp]),
            "orc_sequence_row": (None, [i, vp]),
            "orc_farthest_sequence": (None, [i, vp]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


# ---------------------------------------------------------------------------------------------
class OracleKin:
    """KinModel built from a plainmp_b200.kinematics.KinematicModel, joint state replayed."""

    def __init__(self, kin):
        links = kin.links
        n = len(links)
        parent = np.array([L.parent for L in links], dtype=np.int32)
        xyz = _f64([L.origin_xyz for L in links])
        quat = _f64([L.origin_quat for L in links])
        root = int(np.where(parent < 0)[0][0])
        nj = kin.n_joints
        jtypes = np.array(kin.joint_kinds, dtype=np.int32)  # 1 revolute, 2 continuous, 3 prismatic
        jchild = np.array(kin.joint_child_links, dtype=np.int32)
        jaxes = _f64([links[c].axis for c in kin.joint_child_links]).reshape(-1, 3)
        # links appended by add_new_link are replayed through the oracle's own add_new_link
        n_urdf = kin._n_urdf_links
        self.h = lib().orc_kin_create(
            n_urdf, root, _p(parent[:n_urdf].copy()), _p(xyz[:n_urdf].copy()), _p(quat[:n_urdf].copy()), nj,
            _p(jtypes), _p(jchild), _p(jaxes)
        )
        for parent_id, pxyz, prpy, consider in kin._added_links:
            lib().orc_kin_add_new_link(self.h, int(parent_id), _p(_f64(pxyz)), _p(_f64(prpy)), int(consider))
        # replay the joint state
        for mode, acc in ((1, 1), (2, 0)):
            ids = np.where(kin._joint_set_mode == mode)[0].astype(np.int32)
            if len(ids):
                ang = _f64(kin._joint_angles[ids])
                lib().orc_kin_set_joint_angles(self.h, len(ids), _p(ids), _p(ang), acc)
        lib().orc_kin_set_base_pose(self.h, _p(_f64(kin.get_base_pose())))
        self.kin = kin

    def set_joint_angles(self, joint_ids, angles, accurate=True):
        ids = np.ascontiguousarray(joint_ids, dtype=np.int32)
        lib().orc_kin_set_joint_angles(self.h, len(ids), _p(ids), _p(_f64(angles)), 1 if accurate else 0)

    def link_pose(self, link_id: int) -> np.ndarray:
        out = np.zeros(7)
        lib().orc_kin_get_link_pose(self.h, int(link_id), _p(out))
        return out

    def add_new_link(self, parent_id: int, xyz, rpy, consider_rotation=True) -> int:
        return lib().orc_kin_add_new_link(self.h, int(parent_id), _p(_f64(xyz)), _p(_f64(rpy)), int(consider_rotation))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_kin_destroy(self.h)
            self.h = None


class OracleSDF:
    def __init__(self, h, members=()):
        self.h = h
        self._members = list(members)  # keep children alive (shared_ptr copies are held inside anyway)

    @staticmethod
    def from_psdf(sdf) -> "OracleSDF":
        from plainmp_b200 import psdf

        t = sdf.get_type()
        L = lib()
        if t == psdf.T_UNION:
            ms = [OracleSDF.from_psdf(s) for s in sdf.sdfs]
            arr = (C.c_void_p * len(ms))(*[m.h for m in ms])
            return OracleSDF(L.orc_sdf_union(arr, len(ms)), ms)
        if t == psdf.T_GROUND:
            return OracleSDF(L.orc_sdf_ground(sdf.height))
        if t == psdf.T_CLOUD:
            pts = _f64(sdf.points)
            return OracleSDF(L.orc_sdf_cloud(_p(pts), pts.shape[0], sdf.radius))
        pos, rot = _f64(sdf._pose._position), _f64(sdf._pose._rot)
        if t == psdf.T_BOX:
            return OracleSDF(L.orc_sdf_box(_p(_f64(sdf._width)), _p(pos), _p(rot)))
        if t == psdf.T_CYLINDER:
            return OracleSDF(L.orc_sdf_cylinder(sdf.radius, sdf.height, _p(pos), _p(rot)))
        if t == psdf.T_SPHERE:
            return OracleSDF(L.orc_sdf_sphere(sdf.radius, _p(pos), _p(rot)))
        raise ValueError(t)

    def evaluate_batch(self, pts_nx3) -> np.ndarray:
        p = _f64(pts_nx3)
        out = np.zeros(p.shape[0])
        lib().orc_sdf_evaluate_batch(self.h, _p(p), p.shape[0], _p(out))
        return out

    def is_outside_batch(self, pts_nx3, radii) -> np.ndarray:
        p = _f64(pts_nx3)
        r = _f64(np.broadcast_to(radii, (p.shape[0],)))
        out = np.zeros(p.shape[0], dtype=np.uint8)
        lib().orc_sdf_is_outside_batch(self.h, _p(p), _p(r), p.shape[0], _p(out))
        return out.astype(bool)

    def evaluate(self, p) -> float:
        return lib().orc_sdf_evaluate(self.h, _p(_f64(p)))

    def is_outside(self, p, r) -> bool:
        return bool(lib().orc_sdf_is_outside(self.h, _p(_f64(p)), float(r)))

    def translate(self, t) -> None:
        if lib().orc_sdf_translate(self.h, _p(_f64(t))) != 0:
            raise RuntimeError(lib().orc_last_error().decode())

    def rotate_z(self, a) -> None:
        if lib().orc_sdf_rotate_z(self.h, float(a)) != 0:
            raise RuntimeError(lib().orc_last_error().decode())

    def clone(self) -> "OracleSDF":
        h = lib().orc_sdf_clone(self.h)
        if not h:
            raise RuntimeError(lib().orc_last_error().decode())
        return OracleSDF(h)

    def aabb(self) -> Tuple[np.ndarray, np.ndarray]:
        lb, ub = np.zeros(3), np.zeros(3)
        if lib().orc_sdf_aabb(self.h, _p(lb), _p(ub)) != 0:
            raise RuntimeError("not a primitive")
        return lb, ub

    def pose(self):
        pos, rot, fl = np.zeros(3), np.zeros(9), np.zeros(2, dtype=np.int32)
        if lib().orc_sdf_pose(self.h, _p(pos), _p(rot), _p(fl)) != 0:
            raise RuntimeError("not a transformable primitive")
        return pos, rot.reshape(3, 3), bool(fl[0]), bool(fl[1])

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_sdf_destroy(self.h)
            self.h = None


class OracleKDTree:
    def __init__(self, pts):
        p = _f64(pts)
        self.h = lib().orc_kdtree_create(_p(p), p.shape[0])

    def query(self, p) -> np.ndarray:
        out = np.zeros(3)
        lib().orc_kdtree_query(self.h, _p(_f64(p)), _p(out))
        return out

    def sqdist(self, p) -> float:
        return lib().orc_kdtree_sqdist(self.h, _p(_f64(p)))

    def sqdist_batch(self, pts) -> np.ndarray:
        p = _f64(pts)
        out = np.zeros(p.shape[0])
        lib().orc_kdtree_sqdist_batch(self.h, _p(p), p.shape[0], _p(out))
        return out

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_kdtree_destroy(self.h)
            self.h = None


class OracleCst:
    """SphereCollisionCst of the oracle, mirrored from a plainmp_b200 SphereCollisionCst."""

    def __init__(self, cst, n_threads: int = 1, kin: "Optional[OracleKin]" = None):
        self.kin = kin if kin is not None else OracleKin(cst.get_kin())  # `kin`: shared by the members of a composite
        self.n_threads = n_threads
        offs = [0]
        flat = []
        for sp in cst._group_spheres_unordered:
            flat.append(sp)
            offs.append(offs[-1] + sp.shape[0])
        spheres = _f64(np.concatenate(flat, axis=0)) if flat else np.zeros((0, 4))
        glinks = np.array(cst._group_links, dtype=np.int32)
        goffs = np.array(offs, dtype=np.int32)
        gonly = np.array(cst._group_only_self, dtype=np.uint8)
        pair_links = np.array(
            [[cst._group_links[a], cst._group_links[b]] for a, b in cst._pairs], dtype=np.int32
        ).reshape(-1, 2)
        ctrl = np.array(cst._control_joint_ids, dtype=np.int32)
        self._fixed = OracleSDF.from_psdf(cst._fixed_sdf) if cst._fixed_sdf is not None else None
        self.h = lib().orc_cst_create(
            self.kin.h, len(ctrl), _p(ctrl), int(cst._base_type), len(glinks), _p(glinks), _p(goffs), _p(spheres),
            _p(gonly), pair_links.shape[0], _p(pair_links), self._fixed.h if self._fixed else None,
            1 if cst._reorder_spheres else 0,
        )
        if not self.h:
            raise RuntimeError(lib().orc_last_error().decode())
        self._sdf = None
        if cst.get_sdf() is not None:
            self.set_sdf(OracleSDF.from_psdf(cst.get_sdf()))
        self.q_dim = lib().orc_cst_q_dim(self.h)
        self.n_spheres = lib().orc_cst_n_spheres(self.h)

    def set_sdf(self, sdf: Optional[OracleSDF]) -> None:
        self._sdf = sdf
        lib().orc_cst_set_sdf(self.h, sdf.h if sdf else None)

    def is_valid(self, q) -> bool:
        return bool(lib().orc_cst_is_valid(self.h, _p(_f64(q))))

    def evaluate(self, q) -> np.ndarray:
        v = np.zeros(2)
        n = lib().orc_cst_evaluate(self.h, _p(_f64(q)), _p(v))
        return v[:n]

    def is_valid_batch(self, Q, n_threads: Optional[int] = None) -> np.ndarray:
        Q = _f64(Q)
        out = np.zeros(Q.shape[0], dtype=np.uint8)
        lib().orc_cst_is_valid_batch(self.h, _p(Q), Q.shape[0], n_threads or self.n_threads, _p(out))
        return out.astype(bool)

    def evaluate_batch(self, Q, n_threads: Optional[int] = None) -> np.ndarray:
        Q = _f64(Q)
        out = np.zeros((Q.shape[0], 2))
        lib().orc_cst_evaluate_batch(self.h, _p(Q), Q.shape[0], n_threads or self.n_threads, _p(out))
        return out

    def evaluate_grad_batch(self, Q, n_threads: Optional[int] = None):
        """(vals [n, 2], jac [n, 2, q_dim]) — evaluate() with its Jacobian rows, NaN where disabled."""
        Q = _f64(Q)
        vals = np.zeros((Q.shape[0], 2))
        jac = np.zeros((Q.shape[0], 2, Q.shape[1]))
        lib().orc_cst_evaluate_grad_batch(self.h, _p(Q), Q.shape[0], n_threads or self.n_threads, _p(vals), _p(jac))
        return vals, jac

    def margin_batch(self, Q, n_threads: Optional[int] = None) -> np.ndarray:
        Q = _f64(Q)
        out = np.zeros(Q.shape[0])
        lib().orc_cst_margin_batch(self.h, _p(Q), Q.shape[0], n_threads or self.n_threads, _p(out))
        return out

    WORK_FIELDS = ("joints_rot", "joints_pri", "links_rot", "links_ident", "groups_placed", "spheres_placed", "aabb_tests",
                   "box_tests", "cylinder_tests", "sphere_tests", "ground_tests", "cloud_tests", "pair_tests",
                   "sphere_pair_tests", "valid")
    # SURVEY 8d flop convention: FMA = 2, add/mul/abs/cmp/min/max/sqrt = 1
    WORK_FLOPS = {"joints_rot": 53, "joints_pri": 11, "links_rot": 61, "links_ident": 33, "groups_placed": 24 + 18,
                  "spheres_placed": 18, "aabb_tests": 12, "box_tests": 22, "cylinder_tests": 25, "sphere_tests": 11,
                  "ground_tests": 4, "cloud_tests": 0, "pair_tests": 11, "sphere_pair_tests": 11}

    def work_counts(self, Q, dense: bool = False) -> dict:
        """Instrumented counters of SURVEY 8d summed over Q: the work the reference executes for
        is_valid (lazy FK, broad phases, early exit), or with dense=True the FK + placement part of
        the dense count.  'flops' applies the SURVEY's per-unit convention."""
        Q = _f64(Q)
        out = np.zeros(15, dtype=np.uint64)
        diff = lib().orc_cst_work_counts(self.h, _p(Q), Q.shape[0], 1 if dense else 0, _p(out))
        if diff:
            raise AssertionError(f"instrumented is_valid disagrees with is_valid on {diff} items")
        d = {k: int(v) for k, v in zip(self.WORK_FIELDS, out)}
        d["flops"] = float(sum(self.WORK_FLOPS[k] * d[k] for k in self.WORK_FLOPS))
        d["n"] = int(Q.shape[0])
        return d

    def spheres_batch(self, Q, accurate: bool) -> np.ndarray:
        Q = _f64(Q)
        out = np.zeros((Q.shape[0], self.n_spheres, 4))
        lib().orc_cst_spheres_batch(self.h, _p(Q), Q.shape[0], 1 if accurate else 0, _p(out))
        return out

    def group_spheres(self, q, accurate: bool) -> np.ndarray:
        lib().orc_cst_update_kintree(self.h, _p(_f64(q)), 1 if accurate else 0)
        out = np.zeros((lib().orc_cst_n_groups(self.h), 4))
        lib().orc_cst_get_group_spheres(self.h, _p(out))
        return out

    def sphere_order(self) -> np.ndarray:
        out = np.zeros(self.n_spheres, dtype=np.int32)
        lib().orc_cst_get_sphere_order(self.h, _p(out))
        return out

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_cst_destroy(self.h)
            self.h = None


class OracleValidator:
    def __init__(self, vtype: int, dim: int, box_width=None, resolution: float = 0.0):
        w = _f64(box_width) if box_width is not None else np.zeros(dim)
        self.dim = dim
        self.h = lib().orc_validator_create(int(vtype), dim, _p(w), float(resolution))

    def step_ratio(self, s1, s2) -> float:
        return lib().orc_validator_step_ratio(self.h, _p(_f64(s1)), _p(_f64(s2)))

    def check_motion(self, cst: OracleCst, s1, s2) -> Tuple[bool, int]:
        n = C.c_uint64(0)
        ok = lib().orc_check_motion(self.h, cst.h, _p(_f64(s1)), _p(_f64(s2)), C.byref(n))
        return bool(ok), int(n.value)

    def check_motion_batch(self, cst: OracleCst, S1, S2, n_threads: int = 1) -> Tuple[np.ndarray, int]:
        S1, S2 = _f64(S1), _f64(S2)
        out = np.zeros(S1.shape[0], dtype=np.uint8)
        n = C.c_uint64(0)
        lib().orc_check_motion_batch(self.h, cst.h, _p(S1), _p(S2), S1.shape[0], n_threads, _p(out), C.byref(n))
        return out.astype(bool), int(n.value)

    def check_motion_batch_ex(self, cst: OracleCst, S1, S2, n_threads: int = 1, with_margin: bool = True):
        """(valid [n], calls_per_edge [n], min_margin_per_edge [n] or None): the serial walk per edge, how
        many is_valid calls it makes, and the smallest contact-band margin over the waypoints it visits."""
        S1, S2 = _f64(S1), _f64(S2)
        n = S1.shape[0]
        out = np.zeros(n, dtype=np.uint8)
        calls = np.zeros(n, dtype=np.uint32)
        margin = np.zeros(n) if with_margin else None
        lib().orc_check_motion_batch_ex(self.h, cst.h, _p(S1), _p(S2), n, n_threads, _p(out), _p(calls), _p(margin) if with_margin else None)
        return out.astype(bool), calls, margin

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_validator_destroy(self.h)
            self.h = None


class OracleComposite:
    """IneqCompositeCst of the oracle (composite_constraint.hpp:27-44, 98-105) over OracleCst members
    built on ONE shared OracleKin."""

    def __init__(self, composite, n_threads: int = 1):
        members = list(composite.constraints)
        self.kin = OracleKin(members[0].get_kin())
        self.members = [OracleCst(c, n_threads, kin=self.kin) for c in members]
        self.n_threads = n_threads
        self._h = (C.c_void_p * len(self.members))(*[m.h for m in self.members])
        if lib().orc_composite_check(self._h, len(self.members)) != 0:
            raise RuntimeError("All constraints must have the same kinematic chain")
        self.q_dim = self.members[0].q_dim

    def is_valid_batch(self, Q) -> np.ndarray:
        Q = _f64(Q)
        out = np.zeros(Q.shape[0], dtype=np.uint8)
        lib().orc_composite_is_valid_batch(self._h, len(self.members), _p(Q), Q.shape[0], self.n_threads, _p(out))
        return out.astype(bool)

    def margin_batch(self, Q) -> np.ndarray:
        Q = _f64(Q)
        out = np.zeros(Q.shape[0])
        lib().orc_composite_margin_batch(self._h, len(self.members), _p(Q), Q.shape[0], self.n_threads, _p(out))
        return out

    def evaluate_batch(self, Q) -> np.ndarray:
        """[n, n_members, 2] (ext, self) per member, NaN where a term is disabled."""
        Q = _f64(Q)
        out = np.zeros((Q.shape[0], len(self.members), 2))
        lib().orc_composite_evaluate_batch(self._h, len(self.members), _p(Q), Q.shape[0], self.n_threads, _p(out))
        return out


def sequence_row(n: int) -> List[int]:
    out = np.zeros(n, dtype=np.int32)
    lib().orc_sequence_row(n, _p(out))
    return out.tolist()


def approx_sin_cos(x: float) -> Tuple[float, float]:
    s, c = C.c_double(), C.c_double()
    lib().orc_approx_sin_cos(float(x), C.byref(s), C.byref(c))
    return s.value, c.value
