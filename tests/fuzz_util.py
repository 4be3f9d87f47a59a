"""Random robots and scenes for differential testing (tests/test_fuzz_*.py): random URDF trees with
every joint type, identity and non-identity joint origins, branching, non-controlled (baked) joints,
added links, random sphere groups / pairs / self-only flags, FIXED / PLANAR / FLOATING bases, and
random unions of primitives in the three alignment modes (+ ground, + a small cloud)."""
from __future__ import annotations

import numpy as np

from plainmp_b200 import psdf
from plainmp_b200.constraint import SphereAttachmentSpec, SphereCollisionCst
from plainmp_b200.kinematics import BaseType, KinematicModel
from plainmp_b200.robot_spec import rpy_matrix


def random_urdf(rng: np.random.Generator, n_links: int) -> str:
    lines = ['<?xml version="1.0"?>', '<robot name="fuzz">']
    for i in range(n_links):
        lines.append(f'  <link name="L{i}" />')
    for i in range(1, n_links):
        # chains with occasional branching
        parent = i - 1 if rng.random() < 0.7 else int(rng.integers(0, i))
        jt = rng.choice(["revolute", "continuous", "prismatic", "fixed"], p=[0.45, 0.15, 0.15, 0.25])
        xyz = rng.uniform(-0.3, 0.3, 3) * (rng.random() > 0.15)  # sometimes exactly zero
        mode = rng.integers(0, 3)
        rpy = np.zeros(3) if mode == 0 else (np.array([0.0, 0.0, rng.uniform(-3, 3)]) if mode == 1 else rng.uniform(-3, 3, 3))
        if rng.random() < 0.5:
            axis = np.eye(3)[rng.integers(0, 3)] * rng.choice([-1.0, 1.0])
        else:
            axis = rng.standard_normal(3)
            axis /= np.linalg.norm(axis)
        lines.append(f'  <joint name="J{i}" type="{jt}">')
        lines.append(f'    <origin xyz="{xyz[0]:.9g} {xyz[1]:.9g} {xyz[2]:.9g}" rpy="{rpy[0]:.9g} {rpy[1]:.9g} {rpy[2]:.9g}" />')
        lines.append(f'    <parent link="L{parent}" />')
        lines.append(f'    <child link="L{i}" />')
        if jt != "fixed":
            lines.append(f'    <axis xyz="{axis[0]:.9g} {axis[1]:.9g} {axis[2]:.9g}" />')
        if jt == "revolute":
            lines.append('    <limit lower="-2.5" upper="2.5" effort="1" velocity="1" />')
        elif jt == "prismatic":
            lines.append('    <limit lower="-0.3" upper="0.4" effort="1" velocity="1" />')
        lines.append("  </joint>")
    lines.append("</robot>")
    return "\n".join(lines) + "\n"


def random_scene(rng: np.random.Generator, with_cloud: bool, big: bool = False):
    prims = []
    for _ in range(int(rng.integers(28, 45)) if big else int(rng.integers(1, 7))):
        pos = rng.uniform(-1.0, 1.0, 3) * (2.5 if big else 1.0)
        mode = rng.integers(0, 3)
        rot = np.eye(3) if mode == 0 else (rpy_matrix(0, 0, rng.uniform(-3, 3)) if mode == 1 else rpy_matrix(*rng.uniform(-3, 3, 3)))
        kind = rng.integers(0, 3)
        k = 0.4 if big else 1.0  # many small obstacles instead of a few large ones
        if kind == 0:
            prims.append(psdf.BoxSDF(rng.uniform(0.1, 0.9, 3) * k, psdf.Pose(pos, rot)))
        elif kind == 1:
            prims.append(psdf.CylinderSDF(rng.uniform(0.05, 0.4) * k, rng.uniform(0.1, 1.0) * k, psdf.Pose(pos, rot)))
        else:
            prims.append(psdf.SphereSDF(rng.uniform(0.05, 0.5) * k, psdf.Pose(pos, rot)))
    if rng.random() < (0.0 if big else 0.4):
        prims.append(psdf.GroundSDF(float(rng.uniform(0.0, 1.5))))
    if with_cloud:
        n = int(rng.integers(1, 4000))
        c = rng.uniform(-0.8, 0.8, 3)
        r0 = float(rng.uniform(0.0, 0.02))
        prims.append(psdf.CloudSDF(c + rng.uniform(-0.3, 0.3, (n, 3)) * rng.uniform(0.2, 1.0, 3), r0))
        # several CloudSDF leaves in one scene (round 2): of the same radius (one merged point set on the device)
        # or of different radii (radius classes)
        for _ in range(int(rng.integers(0, 3)) if rng.random() < 0.5 else 0):
            m = int(rng.integers(1, 1500))
            c2 = rng.uniform(-0.8, 0.8, 3)
            r = r0 if rng.random() < 0.4 else float(rng.uniform(0.0, 0.06))
            prims.insert(int(rng.integers(0, len(prims) + 1)), psdf.CloudSDF(c2 + rng.uniform(-0.25, 0.25, (m, 3)), r))
    return psdf.UnionSDF(prims) if len(prims) > 1 or rng.random() < 0.5 else prims[0]


def random_case(seed: int, big: bool = False):
    """(cst, lb, ub, description): a SphereCollisionCst on a random robot with a random scene, and
    sampling bounds for q.  big: 20-45 links, up to 35 groups of up to 20 spheres, dozens of pairs and
    28-44 primitives — the 64-thread CTAs, the queue-overflow path and the >30-primitive path."""
    rng = np.random.default_rng(seed)
    n_links = int(rng.integers(20, 46)) if big else int(rng.integers(2, 13))
    kin = KinematicModel(random_urdf(rng, n_links))
    movable = sorted(kin._joint_ids.keys(), key=lambda n: kin._joint_ids[n])
    # added links (RobotSpec custom_links), one of them a child of another
    n_added = int(rng.integers(0, 3))
    link_names = [f"L{i}" for i in range(n_links)]
    for k in range(n_added):
        parent = str(rng.choice(link_names))
        kin.add_new_link(f"A{k}", parent, rng.uniform(-0.2, 0.2, 3), rng.uniform(-1, 1, 3) * (rng.random() > 0.3), True)
        link_names.append(f"A{k}")
    control = [n for n in movable if rng.random() < 0.75]
    baked = [n for n in movable if n not in control]
    if baked:  # non-controlled joints sit at some angle, set exactly or approximately
        kin.set_joint_positions(kin.get_joint_ids(baked), rng.uniform(-1.0, 1.0, len(baked)), accurate=bool(rng.random() < 0.5))
    base_type = BaseType(int(rng.choice([0, 0, 1, 2])))
    if base_type == BaseType.FIXED and rng.random() < 0.5:
        qb = rng.standard_normal(4)
        qb /= np.linalg.norm(qb)
        kin.set_base_pose([*rng.uniform(-0.5, 0.5, 3), *qb])
    # sphere groups on a random subset of links
    n_groups = int(rng.integers(12, min(len(link_names), 35) + 1)) if big else int(rng.integers(1, min(len(link_names), 8) + 1))
    glinks = list(rng.choice(link_names, size=n_groups, replace=False))
    specs = []
    for ln in glinks:
        m = int(rng.integers(1, 21 if big else 13))
        rad = rng.uniform(0.02, 0.15, m) * (0.3 if big else 1.0)
        if rng.random() < 0.15:
            rad[:] = 0.0  # zero-radius attachments take the r < 1e-6 shortcuts
        specs.append(SphereAttachmentSpec(ln, rng.uniform(-0.15, 0.15, (3, m)) * (0.4 if big else 1.0), rad, bool(rng.random() < 0.2) and base_type == BaseType.FIXED))
    pairs = []
    if n_groups >= 2 and rng.random() < 0.8:
        for _ in range(int(rng.integers(1, (3 if big else 2) * n_groups))):
            a, b = rng.choice(n_groups, size=2, replace=False)
            if big and abs(int(glinks[a][1:]) - int(glinks[b][1:])) < 8:
                continue  # neighbouring links of a random chain overlap all the time
            if (glinks[a], glinks[b]) not in pairs and (glinks[b], glinks[a]) not in pairs:
                pairs.append((glinks[a], glinks[b]))
    fixed_sdf = psdf.UnionSDF([psdf.BoxSDF(rng.uniform(0.1, 0.4, 3), psdf.Pose(rng.uniform(-0.6, 0.6, 3)))]) if (
        rng.random() < 0.3 and base_type == BaseType.FIXED) else None
    cst = SphereCollisionCst(kin, control, base_type, specs, pairs, fixed_sdf, bool(rng.random() < 0.8))
    scene_kind = rng.integers(0, 4)  # 0: none, 1-2: primitives, 3: primitives + cloud
    if scene_kind > 0:
        cst.set_sdf(random_scene(rng, with_cloud=scene_kind == 3, big=big))
    limits = kin.get_joint_position_limits(kin.get_joint_ids(control))
    lb = np.array([max(l[0], -2 * np.pi) for l in limits] + ([-1.0, -1.0, -np.pi] if base_type == BaseType.PLANAR else []) +
                  ([-1.0, -1.0, -1.0, -np.pi, -1.5, -np.pi] if base_type == BaseType.FLOATING else []))
    ub = np.array([min(l[1], 2 * np.pi) for l in limits] + ([1.0, 1.0, np.pi] if base_type == BaseType.PLANAR else []) +
                  ([1.0, 1.0, 1.0, np.pi, 1.5, np.pi] if base_type == BaseType.FLOATING else []))
    desc = f"seed {seed}: {n_links} links (+{n_added}), {len(control)} controlled / {len(baked)} baked joints, base {base_type.name}, {n_groups} groups, {len(pairs)} pairs, scene kind {scene_kind}"
    return cst, lb, ub, desc


def sample_q(lb, ub, n, seed):
    rng = np.random.default_rng(seed)
    if lb.shape[0] == 0:
        return np.zeros((n, 0), dtype=np.float32)
    return rng.uniform(lb, ub, size=(n, lb.shape[0])).astype(np.float32)


def stress_scene(rng, with_cloud):
    """Many thin plates whose AABBs cover the whole workspace while their surfaces stay mostly clear of
    it: every sphere group passes the cull test of almost every primitive on almost every lane, so the
    128-entry environment queue overflows inside single groups and few lanes die early."""
    plates = []
    for _ in range(int(rng.integers(9, 28))):
        r, phi = rng.uniform(1.6, 2.6), rng.uniform(0, 2 * np.pi)
        if rng.random() < 0.7:
            p = psdf.BoxSDF([0.04, 6.0, 6.0], psdf.Pose([r * np.cos(phi), r * np.sin(phi), rng.uniform(-0.3, 0.3)]))
        else:
            p = psdf.CylinderSDF(4.0, 0.05, psdf.Pose([r * np.cos(phi), r * np.sin(phi), 0.0]))
        p.rotate_z(phi + rng.uniform(0.15, 0.35) * np.pi)
        plates.append(p)
    if with_cloud:
        pts = rng.uniform(-1.0, 1.0, (int(rng.integers(50, 400)), 3)) + np.array([2.2, 0.0, 0.0])
        plates.append(psdf.CloudSDF(pts, 0.01))
    return psdf.UnionSDF(plates)
