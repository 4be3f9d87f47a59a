"""ctypes bridge to oracle/_ref/libplainmp_ref.so — the REFERENCE's own C++ classes for the hot path,
compiled from /root/reference (oracle/Makefile.ref) against the Eigen/OMPL API stand-ins of
oracle/ref_shim/.

TEST INFRASTRUCTURE ONLY (tests/, tools/make_golden.py, bench.py's reference arm / cpu_baseline).
Nothing under plainmp_b200/ imports this module.  The library exists only where it was built from
the reference tree (this container; it travels to the GPU box as a prebuilt git-ignored file).

Objects are built from the host mirror objects of plainmp_b200 exactly as a plainmp user would build
the upstream ones: the URDF *text* goes through the reference's own urdf_parser, sphere specs, pair
names and SDFs go through the reference constructors.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path
from typing import Optional, Tuple

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "_ref" / "libplainmp_ref.so"
REFERENCE_ROOT = Path("/root/reference")


def can_build() -> bool:
    return (REFERENCE_ROOT / "cpp" / "plainmp").is_dir()


def available() -> bool:
    return LIB.exists()


def build(force: bool = False) -> Optional[Path]:
    """Compile the reference sources where they lie.  No-op (returns None) without /root/reference."""
    if not can_build():
        return LIB if LIB.exists() else None
    (HERE / "_ref").mkdir(exist_ok=True)
    cmd = ["make", "-C", str(HERE), "-f", "Makefile.ref", "-s"] + (["-B"] if force else [])
    subprocess.run(cmd, check=True)
    return LIB


_lib = None


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _names(names) -> bytes:
    return "\n".join(names).encode()


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB.exists():
            build()
        L = C.CDLL(str(LIB))
        vp, i, d, i64, s = C.c_void_p, C.c_int, C.c_double, C.c_int64, C.c_char_p
        sig = {
            "ref_last_error": (s, []),
            "ref_last_error_is_invalid_argument": (i, []),
            "ref_kin_create": (vp, [s]),
            "ref_kin_destroy": (None, [vp]),
            "ref_kin_add_new_link": (i, [vp, s, s, vp, vp, i]),
            "ref_kin_set_joint_positions": (i, [vp, s, vp, i]),
            "ref_kin_set_base_pose": (None, [vp, vp]),
            "ref_kin_link_pose": (i, [vp, s, vp]),
            "ref_kin_joint_limits": (i, [vp, s, vp]),
            "ref_sdf_box": (vp, [vp, vp, vp]),
            "ref_sdf_cylinder": (vp, [d, d, vp, vp]),
            "ref_sdf_sphere": (vp, [d, vp, vp]),
            "ref_sdf_ground": (vp, [d]),
            "ref_sdf_cloud": (vp, [vp, i64, d]),
            "ref_sdf_union": (vp, [vp, i]),
            "ref_sdf_destroy": (None, [vp]),
            "ref_sdf_clone": (vp, [vp]),
            "ref_sdf_translate": (i, [vp, vp]),
            "ref_sdf_rotate_z": (i, [vp, d]),
            "ref_sdf_evaluate_batch": (None, [vp, vp, i64, vp]),
            "ref_sdf_evaluate_each": (None, [vp, vp, i64, vp]),
            "ref_sdf_is_outside_batch": (None, [vp, vp, vp, i64, vp]),
            "ref_sdf_aabb": (i, [vp, vp, vp]),
            "ref_sdf_pose": (i, [vp, vp, vp, vp]),
            "ref_kdtree_create": (vp, [vp, i64]),
            "ref_kdtree_destroy": (None, [vp]),
            "ref_kdtree_query_batch": (None, [vp, vp, i64, vp, vp]),
            "ref_cst_create": (vp, [vp, s, i, i, s, vp, vp, vp, s, vp, i]),
            "ref_cst_destroy": (None, [vp]),
            "ref_cst_set_sdf": (None, [vp, vp]),
            "ref_cst_q_dim": (i, [vp]),
            "ref_cst_dim": (i, [vp]),
            "ref_cst_n_spheres": (i, [vp]),
            "ref_cst_n_groups": (i, [vp]),
            "ref_cst_is_valid_batch": (None, [vp, vp, i64, vp]),
            "ref_cst_evaluate_batch": (None, [vp, vp, i64, vp, vp]),
            "ref_cst_spheres_batch": (None, [vp, vp, i64, i, i, vp]),
            "ref_cst_flatten": (vp, [vp]),
            "ref_flat_destroy": (None, [vp]),
            "ref_flat_sizes": (None, [vp, vp]),
            "ref_flat_copy": (None, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
            "ref_check_motion_batch": (None, [vp, i, i, vp, d, vp, vp, i64, vp, C.POINTER(C.c_uint64)]),
            "ref_check_motion_batch_ex": (None, [vp, i, i, vp, d, vp, vp, i64, vp, vp, C.POINTER(C.c_uint64)]),
            "ref_composite_create": (vp, [vp, i]),
            "ref_composite_destroy": (None, [vp]),
            "ref_composite_dim": (i, [vp]),
            "ref_composite_is_valid_batch": (None, [vp, vp, i64, vp]),
            "ref_composite_evaluate_batch": (None, [vp, vp, i64, vp, vp]),
            "ref_step_ratio": (d, [i, i, vp, d, vp, vp]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


def _raise():
    msg = lib().ref_last_error().decode()
    if lib().ref_last_error_is_invalid_argument():
        raise ValueError(msg)  # std::invalid_argument -> ValueError, as pybind11 maps it
    raise RuntimeError(msg)


class RefKin:
    """plainmp::kinematics::utils::_KinematicModel built from the same URDF text as the host mirror,
    added links and joint state replayed through the reference's own methods."""

    def __init__(self, kin=None, urdf_xml: Optional[str] = None):
        xml = urdf_xml if urdf_xml is not None else kin._urdf_xml
        self.h = lib().ref_kin_create(xml.encode())
        if not self.h:
            _raise()
        if kin is not None:
            links = kin.links
            for k, (parent_id, xyz, rpy, consider) in enumerate(kin._added_links):
                name = links[kin._n_urdf_links + k].name
                self.add_new_link(name, links[parent_id].name, xyz, rpy, consider)
            id2name = {v: k for k, v in kin._joint_ids.items()}
            for mode, acc in ((1, True), (2, False)):
                ids = [j for j in range(kin.n_joints) if kin._joint_set_mode[j] == mode]
                if ids:
                    self.set_joint_positions([id2name[j] for j in ids], kin._joint_angles[ids], acc)
            self.set_base_pose(kin.get_base_pose())

    def add_new_link(self, name, parent, xyz, rpy, consider_rotation=True) -> int:
        r = lib().ref_kin_add_new_link(self.h, name.encode(), parent.encode(), _p(_f64(xyz)), _p(_f64(rpy)), int(consider_rotation))
        if r < 0:
            _raise()
        return r

    def set_joint_positions(self, joint_names, q, accurate=True) -> None:
        if lib().ref_kin_set_joint_positions(self.h, _names(joint_names), _p(_f64(q)), int(accurate)) != 0:
            _raise()

    def set_base_pose(self, pose7) -> None:
        lib().ref_kin_set_base_pose(self.h, _p(_f64(pose7)))

    def link_pose(self, link_name: str) -> np.ndarray:
        out = np.zeros(7)
        if lib().ref_kin_link_pose(self.h, link_name.encode(), _p(out)) != 0:
            _raise()
        return out

    def joint_limits(self, joint_names) -> np.ndarray:
        out = np.zeros((len(joint_names), 2))
        if lib().ref_kin_joint_limits(self.h, _names(joint_names), _p(out)) < 0:
            _raise()
        return out

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_kin_destroy(self.h)
            self.h = None


class RefSDF:
    def __init__(self, h, members=()):
        if not h:
            _raise()
        self.h = h
        self._members = list(members)

    @staticmethod
    def from_psdf(sdf) -> "RefSDF":
        from plainmp_b200 import psdf

        t = sdf.get_type()
        L = lib()
        if t == psdf.T_UNION:
            ms = [RefSDF.from_psdf(s) for s in sdf.sdfs]
            arr = (C.c_void_p * len(ms))(*[m.h for m in ms])
            return RefSDF(L.ref_sdf_union(arr, len(ms)), ms)
        if t == psdf.T_GROUND:
            return RefSDF(L.ref_sdf_ground(sdf.height))
        if t == psdf.T_CLOUD:
            pts = _f64(sdf.points)
            return RefSDF(L.ref_sdf_cloud(_p(pts), pts.shape[0], sdf.radius))
        pos, rot = _f64(sdf._pose._position), _f64(sdf._pose._rot)
        if t == psdf.T_BOX:
            return RefSDF(L.ref_sdf_box(_p(_f64(sdf._width)), _p(pos), _p(rot)))
        if t == psdf.T_CYLINDER:
            return RefSDF(L.ref_sdf_cylinder(sdf.radius, sdf.height, _p(pos), _p(rot)))
        if t == psdf.T_SPHERE:
            return RefSDF(L.ref_sdf_sphere(sdf.radius, _p(pos), _p(rot)))
        raise ValueError(t)

    def evaluate_batch(self, pts_nx3, each: bool = False) -> np.ndarray:
        p = _f64(pts_nx3)
        out = np.zeros(p.shape[0])
        (lib().ref_sdf_evaluate_each if each else lib().ref_sdf_evaluate_batch)(self.h, _p(p), p.shape[0], _p(out))
        return out

    def is_outside_batch(self, pts_nx3, radii) -> np.ndarray:
        p = _f64(pts_nx3)
        r = _f64(np.broadcast_to(radii, (p.shape[0],)))
        out = np.zeros(p.shape[0], dtype=np.uint8)
        lib().ref_sdf_is_outside_batch(self.h, _p(p), _p(r), p.shape[0], _p(out))
        return out.astype(bool)

    def translate(self, t) -> None:
        if lib().ref_sdf_translate(self.h, _p(_f64(t))) != 0:
            _raise()

    def rotate_z(self, a) -> None:
        if lib().ref_sdf_rotate_z(self.h, float(a)) != 0:
            _raise()

    def clone(self) -> "RefSDF":
        return RefSDF(lib().ref_sdf_clone(self.h))

    def aabb(self) -> Tuple[np.ndarray, np.ndarray]:
        lb, ub = np.zeros(3), np.zeros(3)
        if lib().ref_sdf_aabb(self.h, _p(lb), _p(ub)) != 0:
            raise RuntimeError("not a primitive")
        return lb, ub

    def pose(self):
        pos, rot, fl = np.zeros(3), np.zeros(9), np.zeros(2, dtype=np.int32)
        if lib().ref_sdf_pose(self.h, _p(pos), _p(rot), _p(fl)) != 0:
            raise RuntimeError("not a transformable primitive")
        return pos, rot.reshape(3, 3), bool(fl[0]), bool(fl[1])

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_sdf_destroy(self.h)
            self.h = None


class RefKDTree:
    def __init__(self, pts):
        p = _f64(pts)
        self.h = lib().ref_kdtree_create(_p(p), p.shape[0])

    def sqdist_batch(self, pts) -> np.ndarray:
        p = _f64(pts)
        out = np.zeros(p.shape[0])
        lib().ref_kdtree_query_batch(self.h, _p(p), p.shape[0], _p(out), None)
        return out

    def query_batch(self, pts) -> np.ndarray:
        p = _f64(pts)
        out = np.zeros((p.shape[0], 3))
        lib().ref_kdtree_query_batch(self.h, _p(p), p.shape[0], None, _p(out))
        return out

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_kdtree_destroy(self.h)
            self.h = None


class RefCst:
    """plainmp::constraint::SphereCollisionCst mirrored from a plainmp_b200 SphereCollisionCst
    (un-reordered specs, link/pair NAMES, SDF tree) on its own private kinematic model."""

    def __init__(self, cst, kin: "Optional[RefKin]" = None):
        self.kin = kin if kin is not None else RefKin(cst.get_kin())  # `kin`: shared by the members of a composite
        offs, flat = [0], []
        for sp in cst._group_spheres_unordered:
            flat.append(sp)
            offs.append(offs[-1] + sp.shape[0])
        spheres = _f64(np.concatenate(flat, axis=0)) if flat else np.zeros((0, 4))
        goffs = np.array(offs, dtype=np.int32)
        gonly = np.array(cst._group_only_self, dtype=np.uint8)
        pair_names = []
        for a, b in cst._pairs:
            pair_names += [cst._group_link_names[a], cst._group_link_names[b]]
        self._fixed = RefSDF.from_psdf(cst._fixed_sdf) if cst._fixed_sdf is not None else None
        self.h = lib().ref_cst_create(
            self.kin.h, _names(cst._control_joint_names), int(cst._base_type), len(cst._group_link_names),
            _names(cst._group_link_names), _p(goffs), _p(spheres), _p(gonly), _names(pair_names),
            self._fixed.h if self._fixed else None, 1 if cst._reorder_spheres else 0,
        )
        if not self.h:
            _raise()
        self._sdf = None
        if cst.get_sdf() is not None:
            self.set_sdf(RefSDF.from_psdf(cst.get_sdf()))
        self.q_dim = lib().ref_cst_q_dim(self.h)
        self.n_spheres = lib().ref_cst_n_spheres(self.h)
        self.n_groups = lib().ref_cst_n_groups(self.h)

    def set_sdf(self, sdf: Optional[RefSDF]) -> None:
        self._sdf = sdf
        lib().ref_cst_set_sdf(self.h, sdf.h if sdf else None)

    def is_valid_batch(self, Q) -> np.ndarray:
        Q = _f64(Q)
        out = np.zeros(Q.shape[0], dtype=np.uint8)
        lib().ref_cst_is_valid_batch(self.h, _p(Q), Q.shape[0], _p(out))
        return out.astype(bool)

    def evaluate_batch(self, Q) -> np.ndarray:
        Q = _f64(Q)
        out = np.zeros((Q.shape[0], 2))
        lib().ref_cst_evaluate_batch(self.h, _p(Q), Q.shape[0], _p(out), None)
        return out

    def evaluate_grad_batch(self, Q):
        Q = _f64(Q)
        vals = np.zeros((Q.shape[0], 2))
        jac = np.zeros((Q.shape[0], 2, Q.shape[1]))
        lib().ref_cst_evaluate_batch(self.h, _p(Q), Q.shape[0], _p(vals), _p(jac))
        return vals, jac

    def spheres_batch(self, Q, accurate: bool) -> np.ndarray:
        Q = _f64(Q)
        out = np.zeros((Q.shape[0], self.n_spheres, 4))
        lib().ref_cst_spheres_batch(self.h, _p(Q), Q.shape[0], int(accurate), 0, _p(out))
        return out

    def group_spheres_batch(self, Q, accurate: bool) -> np.ndarray:
        Q = _f64(Q)
        out = np.zeros((Q.shape[0], self.n_groups, 4))
        lib().ref_cst_spheres_batch(self.h, _p(Q), Q.shape[0], int(accurate), 1, _p(out))
        return out

    def flatten(self) -> dict:
        """include/plainmp_b200_plainmp.hpp run on the reference's own objects: the C-ABI tables
        (numpy arrays in the dtypes of plainmp_b200._capi) a plainmp-side binding would hand to
        pmp_robot_create / pmp_scene_create."""
        from plainmp_b200 import _capi

        f = lib().ref_cst_flatten(self.h)
        if not f:
            _raise()
        try:
            sz = np.zeros(8, dtype=np.int64)
            lib().ref_flat_sizes(f, _p(sz))
            out = {
                "ops": np.zeros(sz[0], dtype=_capi.LINK_OP_DTYPE),
                "groups": np.zeros(sz[1], dtype=_capi.GROUP_DTYPE),
                "spheres": np.zeros((sz[2], 4), dtype=np.float32),
                "pairs": np.zeros((sz[3], 2), dtype=np.int32),
                "prims": np.zeros(sz[4], dtype=_capi.PRIM_DTYPE),
                "cloud_points": np.zeros((sz[5], 3), dtype=np.float32),
                "base_pose": np.zeros(7, dtype=np.float32),
                "dof": int(sz[6]),
                "base_type": int(sz[7]),
            }
            cr = np.zeros(1, dtype=np.float32)
            radii = np.zeros(max(int(sz[5]), 1), dtype=np.float32)
            has = np.zeros(1, dtype=np.int32)
            lib().ref_flat_copy(f, _p(out["ops"]), _p(out["groups"]), _p(out["spheres"]), _p(out["pairs"]), _p(out["prims"]),
                                _p(out["cloud_points"]), _p(out["base_pose"]), _p(cr), _p(radii), _p(has))
            out["cloud_radius"] = float(cr[0])
            out["cloud_point_radii"] = radii[: int(sz[5])] if has[0] else None  # CloudSDF leaves of different radii
            return out
        finally:
            lib().ref_flat_destroy(f)

    def check_motion_batch(self, vtype: int, box_width, resolution: float, S1, S2) -> Tuple[np.ndarray, int]:
        S1, S2 = _f64(S1), _f64(S2)
        dim = S1.shape[1]
        w = _f64(box_width) if box_width is not None else np.zeros(dim)
        out = np.zeros(S1.shape[0], dtype=np.uint8)
        n = C.c_uint64(0)
        lib().ref_check_motion_batch(self.h, int(vtype), dim, _p(w), float(resolution), _p(S1), _p(S2), S1.shape[0], _p(out), C.byref(n))
        return out.astype(bool), int(n.value)

    def check_motion_batch_ex(self, vtype: int, box_width, resolution: float, S1, S2):
        """(valid [n], calls_per_edge [n]): checkMotion per edge and how many times it called isValid."""
        S1, S2 = _f64(S1), _f64(S2)
        dim = S1.shape[1]
        w = _f64(box_width) if box_width is not None else np.zeros(dim)
        out = np.zeros(S1.shape[0], dtype=np.uint8)
        calls = np.zeros(S1.shape[0], dtype=np.uint32)
        n = C.c_uint64(0)
        lib().ref_check_motion_batch_ex(self.h, int(vtype), dim, _p(w), float(resolution), _p(S1), _p(S2), S1.shape[0], _p(out), _p(calls), C.byref(n))
        return out.astype(bool), calls

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_cst_destroy(self.h)
            self.h = None


class RefComposite:
    """plainmp::constraint::IneqCompositeCst (composite_constraint.hpp:88-105) over RefCst members that
    share ONE reference KinematicModel, as bindings/constraint.cpp:108-113 constructs it."""

    def __init__(self, composite):
        members = list(composite.constraints)
        self.kin = RefKin(members[0].get_kin())
        self.members = [RefCst(c, kin=self.kin) for c in members]
        hs = (C.c_void_p * len(self.members))(*[m.h for m in self.members])
        self.h = lib().ref_composite_create(hs, len(self.members))
        if not self.h:
            _raise()
        self.dim = lib().ref_composite_dim(self.h)
        self.q_dim = self.members[0].q_dim

    def is_valid_batch(self, Q) -> np.ndarray:
        Q = _f64(Q)
        out = np.zeros(Q.shape[0], dtype=np.uint8)
        lib().ref_composite_is_valid_batch(self.h, _p(Q), Q.shape[0], _p(out))
        return out.astype(bool)

    def evaluate_batch(self, Q, with_jac: bool = False):
        """vals [n, cst_dim] (+ jac [n, cst_dim, q_dim]): the enabled terms of the members stacked in order."""
        Q = _f64(Q)
        vals = np.zeros((Q.shape[0], self.dim))
        jac = np.zeros((Q.shape[0], self.dim, Q.shape[1])) if with_jac else None
        lib().ref_composite_evaluate_batch(self.h, _p(Q), Q.shape[0], _p(vals), _p(jac) if with_jac else None)
        return (vals, jac) if with_jac else vals

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_composite_destroy(self.h)
            self.h = None


def step_ratio(vtype: int, box_width, resolution: float, s1, s2) -> float:
    s1, s2 = _f64(s1), _f64(s2)
    w = _f64(box_width) if box_width is not None else np.zeros(len(s1))
    return lib().ref_step_ratio(int(vtype), len(s1), _p(w), float(resolution), _p(s1), _p(s2))


def is_valid_batch_threads(cst, Q, n_threads: int) -> np.ndarray:
    """The reference is single-threaded with a stateful kinematic model: use one private
    kin + cst per thread over disjoint slices (ctypes releases the GIL during the C++ loop)."""
    import threading

    Q = _f64(Q)
    n_threads = max(1, min(n_threads, len(Q)))
    insts = [RefCst(cst) for _ in range(n_threads)]
    out = np.zeros(len(Q), dtype=bool)
    bounds = np.linspace(0, len(Q), n_threads + 1).astype(int)

    def work(t):
        out[bounds[t] : bounds[t + 1]] = insts[t].is_valid_batch(Q[bounds[t] : bounds[t + 1]])

    ths = [threading.Thread(target=work, args=(t,)) for t in range(n_threads)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    return out


def check_motion_batch_threads(cst, vtype: int, box_width, resolution: float, S1, S2, n_threads: int):
    """check_motion_batch_ex over disjoint slices, one private reference kin + cst per thread."""
    import threading

    S1, S2 = _f64(S1), _f64(S2)
    n_threads = max(1, min(n_threads, len(S1)))
    insts = [RefCst(cst) for _ in range(n_threads)]
    valid = np.zeros(len(S1), dtype=bool)
    calls = np.zeros(len(S1), dtype=np.uint32)
    bounds = np.linspace(0, len(S1), n_threads + 1).astype(int)

    def work(t):
        lo, hi = bounds[t], bounds[t + 1]
        valid[lo:hi], calls[lo:hi] = insts[t].check_motion_batch_ex(vtype, box_width, resolution, S1[lo:hi], S2[lo:hi])

    ths = [threading.Thread(target=work, args=(t,)) for t in range(n_threads)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    return valid, calls
