"""The BASELINE.json workloads: robots, scenes and seeded inputs shared by the
tests, bench.py and smoke().  Scene definitions follow the reference's bench scripts:
example/bench/panda_plan.py:19-34, example/bench/fetch_plan.py:14-19, example/fetch_plan.py:24-30.
"""
from __future__ import annotations

import numpy as np

from . import psdf
from .kinematics import BaseType
from .robot_spec import FetchSpec, OpenArmV10DualSpec, PandaSpec, PR2DualarmSpec, PR2RarmSpec


def panda_dual_bars_sdf(ceiled: bool = False) -> psdf.UnionSDF:
    height = 1.0
    prims = [
        psdf.BoxSDF([2.0, 2.0, 0.05], psdf.Pose([0.0, 0.0, -0.05])),
        psdf.CylinderSDF(0.05, height, psdf.Pose([0.3, 0.3, 0.5 * height])),
        psdf.CylinderSDF(0.05, height, psdf.Pose([-0.3, -0.3, 0.5 * height])),
    ]
    if ceiled:
        prims.append(psdf.BoxSDF([2.0, 2.0, 0.05], psdf.Pose([0.0, 0.0, height])))
    return psdf.UnionSDF(prims)


def fetch_table_sdf() -> psdf.UnionSDF:
    table = psdf.BoxSDF([1.0, 2.0, 0.05], psdf.Pose([0.95, 0.0, 0.8]))
    return psdf.UnionSDF([table, psdf.GroundSDF(0.0)])


def table_cloud(n_points: int, seed: int = 4) -> np.ndarray:
    """Synthetic depth-camera-like cloud for config 4: half of the points sample the top slab of
    a 1 x 2 x 0.05 table centred at (0.8, 0, 0.8) (as example/fetch_plan.py:24-30 samples its
    table), the other half sample a dozen box-shaped objects standing on it.  (The SURVEY's
    'uniform clutter in the whole workspace' variant makes every configuration collide at these
    densities, which would turn the workload into pure early-exit.)"""
    rng = np.random.default_rng(seed)
    n_slab = n_points // 2
    slab = rng.uniform([0.3, -1.0, 0.775], [1.3, 1.0, 0.825], size=(n_slab, 3))
    n_obj = 12
    centres = rng.uniform([0.45, -0.85, 0.0], [1.15, 0.85, 0.0], size=(n_obj, 3))
    sizes = rng.uniform([0.06, 0.06, 0.08], [0.2, 0.2, 0.35], size=(n_obj, 3))
    which = rng.integers(0, n_obj, size=n_points - n_slab)
    u = rng.uniform(-0.5, 0.5, size=(n_points - n_slab, 3))
    obj = centres[which] + u * sizes[which]
    obj[:, 2] = 0.825 + (u[:, 2] + 0.5) * sizes[which, 2]
    return np.concatenate([slab, obj], axis=0).astype(np.float32).astype(np.float64)


def make_panda(ceiled: bool = False, self_collision: bool = True):
    spec = PandaSpec()
    cst = spec.create_collision_const(self_collision=self_collision, use_cache=False)
    cst.set_sdf(panda_dual_bars_sdf(ceiled))
    return spec, cst


def make_fetch():
    spec = FetchSpec()
    cst = spec.create_collision_const(use_cache=False)
    cst.set_sdf(fetch_table_sdf())
    return spec, cst


def make_pr2_cloud(n_points: int, seed: int = 4):
    spec = PR2DualarmSpec()
    spec.reflect_default_joint_positions()
    cst = spec.create_collision_const(use_cache=False)
    cst.set_sdf(psdf.CloudSDF(table_cloud(n_points, seed), 0.002))
    return spec, cst


def make_pr2_table():
    """PR2 dual arm against the table as primitives (the slab the cloud of config 4 samples)."""
    spec = PR2DualarmSpec()
    spec.reflect_default_joint_positions()
    cst = spec.create_collision_const(use_cache=False)
    cst.set_sdf(psdf.UnionSDF([psdf.BoxSDF([1.0, 2.0, 0.05], psdf.Pose([0.8, 0.0, 0.8])), psdf.GroundSDF(0.0)]))
    return spec, cst


def make_openarm():
    spec = OpenArmV10DualSpec()
    cst = spec.create_collision_const(use_cache=False)
    cst.set_sdf(psdf.UnionSDF([psdf.BoxSDF([0.6, 1.2, 0.05], psdf.Pose([0.45, 0.0, 0.55])), psdf.GroundSDF(0.0)]))
    return spec, cst


def random_configs(spec, n: int, seed: int) -> np.ndarray:
    """Uniform in the joint limits (angle_bounds: +-inf -> +-2 pi), rounded to float32 so that
    the FP64 oracle and the FP32 kernels consume bit-identical inputs."""
    lb, ub = spec.angle_bounds()
    rng = np.random.default_rng(seed)
    return rng.uniform(lb, ub, size=(n, lb.shape[0])).astype(np.float32)


def random_edges(spec, n: int, seed: int, planner_like: bool, valid_fn=None):
    """Distribution A: both endpoints uniform in the joint limits (long edges, almost all
    invalid: the early-exit regime).  Distribution B: q1 = q0 + U(-0.5, 0.5)^dof clipped
    (planner-like short edges); when ``valid_fn`` (q[n, dof] float32 -> bool[n]) is given, q0 is
    rejection-sampled to be collision-free as SURVEY §8d-3 asks (checkMotion assumes s1 valid)."""
    lb, ub = spec.angle_bounds()
    rng = np.random.default_rng(seed)
    dof = lb.shape[0]
    if valid_fn is None:
        q0 = rng.uniform(lb, ub, size=(n, dof)).astype(np.float32)
    else:
        chunks, have = [], 0
        while have < n:
            cand = rng.uniform(lb, ub, size=(max(1024, 2 * (n - have)), dof)).astype(np.float32)
            keep = cand[np.asarray(valid_fn(cand), dtype=bool)]
            chunks.append(keep)
            have += keep.shape[0]
        q0 = np.concatenate(chunks, axis=0)[:n]
    if planner_like:
        q1 = np.clip(q0.astype(np.float64) + rng.uniform(-0.5, 0.5, size=q0.shape), lb, ub).astype(np.float32)
    else:
        q1 = rng.uniform(lb, ub, size=q0.shape).astype(np.float32)
    return q0, q1


def make_jaxon(gripper_collision: bool = True):
    """JaxonSpec (FLOATING base, 31 joints + 6 base entries, the largest sphere table of conf/) standing
    next to a table and under a beam, on the ground."""
    from .robot_spec import JaxonSpec

    spec = JaxonSpec(gripper_collision=gripper_collision)
    cst = spec.create_collision_const(use_cache=False)
    cst.set_sdf(
        psdf.UnionSDF(
            [
                psdf.BoxSDF([0.8, 1.6, 0.05], psdf.Pose([0.9, 0.0, 0.8])),
                psdf.BoxSDF([0.2, 3.0, 0.2], psdf.Pose([0.3, 0.0, 2.3])),
                psdf.CylinderSDF(0.1, 1.2, psdf.Pose([0.6, -0.9, 0.6])),
                psdf.GroundSDF(0.0),
            ]
        )
    )
    return spec, cst


def make_pr2_base_only():
    """PR2BaseOnlySpec (no controlled joint, PLANAR base: q = x, y, yaw) between furniture-sized boxes."""
    from .robot_spec import PR2BaseOnlySpec

    spec = PR2BaseOnlySpec()
    spec.reflect_default_joint_positions()
    cst = spec.create_collision_const(use_cache=False)
    rot = np.array([[np.cos(0.4), -np.sin(0.4), 0], [np.sin(0.4), np.cos(0.4), 0], [0, 0, 1.0]])
    cst.set_sdf(
        psdf.UnionSDF(
            [
                psdf.BoxSDF([1.0, 2.0, 0.8], psdf.Pose([1.6, 0.0, 0.4])),
                psdf.BoxSDF([0.6, 0.6, 1.8], psdf.Pose([-0.5, 1.4, 0.9], rot)),
                psdf.CylinderSDF(0.25, 1.0, psdf.Pose([0.2, -1.5, 0.5])),
            ]
        )
    )
    return spec, cst


def random_configs_moving_base(spec, n: int, seed: int, base_lb, base_ub) -> np.ndarray:
    """Joints uniform in their limits, base entries uniform in [base_lb, base_ub] (3 entries for a PLANAR
    base, 6 for a FLOATING one), rounded to float32 like random_configs."""
    kin = spec.get_kin()
    limits = kin.get_joint_position_limits(kin.get_joint_ids(spec.control_joint_names))
    lb = np.array([l[0] for l in limits] + list(base_lb), dtype=np.float64)
    ub = np.array([l[1] for l in limits] + list(base_ub), dtype=np.float64)
    lb[lb == -np.inf] = -2 * np.pi
    ub[ub == np.inf] = 2 * np.pi
    rng = np.random.default_rng(seed)
    return rng.uniform(lb, ub, size=(n, lb.shape[0])).astype(np.float32)


def make_panda_position_only_link(offset_spheres: bool = False):
    """Panda with a link added with consider_rotation=False (kinematics.cpp:300-337: FK gives such a link
    its position only) that carries spheres at its origin and a self-collision pair with the forearm."""
    from .constraint import SphereAttachmentSpec, SphereCollisionCst
    from .kinematics import BaseType, KinematicModel

    spec = PandaSpec()
    kin = KinematicModel(spec.get_kin()._urdf_xml)
    kin.add_new_link("tool_tip", "panda_hand", np.array([0.02, -0.01, 0.22]), np.array([0.3, -0.2, 0.5]), False)
    pos = np.array([[0.01, 0, 0]] if offset_spheres else [[0.0, 0, 0], [0.0, 0, 0]]).T
    specs = [
        SphereAttachmentSpec("tool_tip", pos, np.array([0.06] if offset_spheres else [0.06, 0.0]), False),
        SphereAttachmentSpec("panda_link5", np.array([[0.0, 0.05, -0.1], [0.0, 0.0, 0.0]]).T, np.array([0.07, 0.05]), False),
    ]
    cst = SphereCollisionCst(kin, spec.control_joint_names, BaseType.FIXED, specs, [("tool_tip", "panda_link5")], None)
    cst.set_sdf(panda_dual_bars_sdf())
    return spec, cst
