"""ctypes binding of libplainmp_b200.so (include/plainmp_b200.h).

The library is the product: if it is missing, or there is no CUDA device, every compute call
raises.  There is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
REPO_ROOT = PKG_DIR.parent
# PMP_B200_LIB: another build of the same sources (kernel A/B runs on the GPU box, tools/gpu_ab.sh); never a fallback
LIB_PATH = Path(os.environ["PMP_B200_LIB"]) if os.environ.get("PMP_B200_LIB") else PKG_DIR / "libplainmp_b200.so"
SRC_DIR = PKG_DIR / "csrc"

NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-O3",
    "-std=c++17",
    "-shared",
    "-Xcompiler",
    "-fPIC",
]

PMP_OK, PMP_ERR_INVALID, PMP_ERR_CUDA, PMP_ERR_LIMIT = 0, -1, -2, -3
OP_FIXED, OP_REVOLUTE, OP_PRISMATIC = 0, 1, 2
PRIM_BOX, PRIM_CYLINDER, PRIM_SPHERE, PRIM_GROUND = 1, 2, 3, 4
PRIMFLAG_AXIS_ALIGNED, PRIMFLAG_Z_ALIGNED = 1, 2
OPFLAG_ORIGIN_IDENTITY = 1
VALIDATOR_BOX, VALIDATOR_EUCLIDEAN = 0, 1

LINK_OP_DTYPE = np.dtype(
    [
        ("origin_quat", "<f4", 4),
        ("origin_xyz", "<f4", 3),
        ("kind", "<i4"),
        ("axis", "<f4", 3),
        ("q_index", "<i4"),
        ("parent", "<i4"),
        ("flags", "<i4"),
        ("reserved", "<i4", 2),
    ]
)
GROUP_DTYPE = np.dtype(
    [
        ("center", "<f4", 3),
        ("radius", "<f4"),
        ("op", "<i4"),
        ("sphere_begin", "<i4"),
        ("sphere_count", "<i4"),
        ("only_self_collision", "<i4"),
    ]
)
PRIM_DTYPE = np.dtype(
    [
        ("position", "<f4", 3),
        ("type", "<i4"),
        ("rot_col0", "<f4", 3),
        ("param0", "<f4"),
        ("rot_col1", "<f4", 3),
        ("param1", "<f4"),
        ("rot_col2", "<f4", 3),
        ("param2", "<f4"),
        ("lb", "<f4", 3),
        ("flags", "<i4"),
        ("ub", "<f4", 3),
        ("reserved", "<i4"),
    ]
)
assert LINK_OP_DTYPE.itemsize == 64 and GROUP_DTYPE.itemsize == 32 and PRIM_DTYPE.itemsize == 96


class RobotDesc(C.Structure):
    _fields_ = [
        ("dof", C.c_int32),
        ("base_type", C.c_int32),
        ("base_pose", C.c_float * 7),
        ("n_ops", C.c_int32),
        ("ops", C.c_void_p),
        ("n_groups", C.c_int32),
        ("groups", C.c_void_p),
        ("n_spheres", C.c_int32),
        ("spheres", C.c_void_p),
        ("n_pairs", C.c_int32),
        ("pairs", C.c_void_p),
    ]


class SceneDesc(C.Structure):
    _fields_ = [
        ("n_prims", C.c_int32),
        ("prims", C.c_void_p),
        ("n_cloud_points", C.c_int64),
        ("cloud_points", C.c_void_p),
        ("cloud_radius", C.c_float),
        ("cloud_cell", C.c_float),
        ("cloud_point_radii", C.c_void_p),  # optional [n] float32: CloudSDF leaves of different radii in one scene
    ]


class Validator(C.Structure):
    _fields_ = [
        ("type", C.c_int32),
        ("dim", C.c_int32),
        ("box_width", C.c_void_p),
        ("resolution", C.c_double),
    ]


# every symbol include/plainmp_b200.h declares: name -> (restype, argtypes)
_p, _i64, _i32 = C.c_void_p, C.c_int64, C.c_int
SYMBOLS = {
    "pmp_last_error": (C.c_char_p, []),
    "pmp_abi_version": (_i32, []),
    "pmp_device_count": (_i32, []),
    "pmp_launch_count": (C.c_uint64, []),
    "pmp_host_alloc": (_i32, [C.c_size_t, _i32, C.POINTER(C.c_void_p)]),
    "pmp_host_free": (_i32, [_p]),
    "pmp_set_device": (_i32, [_i32]),
    "pmp_measure_fp32_peak": (_i32, [C.POINTER(C.c_double)]),
    "pmp_robot_create": (_i32, [C.POINTER(RobotDesc), C.POINTER(_p)]),
    "pmp_robot_create_ex": (_i32, [C.POINTER(RobotDesc), _p, C.POINTER(_p)]),
    "pmp_robot_destroy": (None, [_p]),
    "pmp_robot_q_dim": (_i32, [_p]),
    "pmp_robot_n_spheres": (_i32, [_p]),
    "pmp_scene_create": (_i32, [C.POINTER(SceneDesc), C.POINTER(_p)]),
    "pmp_scene_destroy": (None, [_p]),
    "pmp_scene_update_prims": (_i32, [_p, _p, _i32, _p]),
    "pmp_scene_update_cloud_ex": (_i32, [_p, _p, _i64, C.c_float, _p, C.c_float]),
    "pmp_scene_update_cloud": (_i32, [_p, _p, _i64, C.c_float, C.c_float]),
    "pmp_check_configs": (_i32, [_p, _p, _p, _i64, _p, _p, _p]),
    "pmp_eval_configs": (_i32, [_p, _p, _p, _i64, _p, _p]),
    "pmp_eval_grad_configs": (_i32, [_p, _p, _p, _i64, _p, _p, _p]),
    "pmp_fk_spheres": (_i32, [_p, _p, _i64, _i32, _p, _p]),
    "pmp_fk_frames": (_i32, [_p, _p, _i64, _i32, _p, _p]),
    "pmp_check_edges": (_i32, [_p, _p, C.POINTER(Validator), _p, _p, _i64, _p, _p, _p, _p]),
    "pmp_check_edges_ex": (_i32, [_p, _p, C.POINTER(Validator), _p, _p, _i32, _i64, _p, _p, _p, _p, _p]),
    "pmp_check_configs_host": (_i32, [_p, _p, _p, _i64, _p, C.POINTER(C.c_uint64)]),
    "pmp_check_configs_host_f64": (_i32, [_p, _p, _p, _i64, _p, C.POINTER(C.c_uint64)]),
    "pmp_eval_configs_host": (_i32, [_p, _p, _p, _i64, _p]),
    "pmp_check_edges_host": (
        _i32,
        [_p, _p, C.POINTER(Validator), _p, _p, _i64, _p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)],
    ),
    "pmp_check_edges_host_f64": (
        _i32,
        [_p, _p, C.POINTER(Validator), _p, _p, _i64, _p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)],
    ),
    "pmp_session_create": (_i32, [_p, _p, C.POINTER(Validator), _i64, C.POINTER(_p)]),
    "pmp_session_destroy": (None, [_p]),
    "pmp_session_set_scene": (_i32, [_p, _p]),
    "pmp_session_is_valid": (_i32, [_p, _p, _i64, _p]),
    "pmp_session_check_motion": (_i32, [_p, _p, _p, _i64, _p, C.POINTER(C.c_uint64)]),
    "pmp_sdf_evaluate": (_i32, [_p, _p, _i64, _p, _p]),
    "pmp_sdf_is_outside": (_i32, [_p, _p, _p, _i64, _p, _p]),
    "pmp_cloud_nearest": (_i32, [_p, _p, _i64, _p, _p, _p]),
}


class PmpError(RuntimeError):
    """A non-zero status from the C ABI (the reference throws std::runtime_error)."""


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile csrc/ into libplainmp_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = sorted(SRC_DIR.glob("*.cu")) + sorted(SRC_DIR.glob("*.cuh")) + [REPO_ROOT / "include" / "plainmp_b200.h"]
    if not force and LIB_PATH.exists() and all(LIB_PATH.stat().st_mtime >= s.stat().st_mtime for s in srcs):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc, *NVCC_FLAGS, "--threads", "2", "-o", str(LIB_PATH), str(SRC_DIR / "plainmp_b200.cu"), str(SRC_DIR / "pmp_cloud_build.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.run(cmd, check=True)
    return LIB_PATH


_lib = None


def load() -> C.CDLL:
    """Load the shared library (no compute happens here).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise PmpError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(plainmp_b200 has no CPU fallback)"
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.pmp_abi_version() != 1:
        raise PmpError("libplainmp_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != PMP_OK:
        msg = load().pmp_last_error().decode()
        if status == PMP_ERR_INVALID:
            raise ValueError(msg) if "no link" in msg or "no joint" in msg else PmpError(msg)
        raise PmpError(msg)


def require_cuda() -> None:
    if load().pmp_device_count() < 1:
        raise PmpError("plainmp_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
