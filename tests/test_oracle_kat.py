"""Pin the FP64 oracle against every known-answer fact and property the REFERENCE's own tests and
comments hold for the hot path.  CPU only.

reference tests mirrored here (paths relative to /root/reference):
  tests/test_psdf.py:27-39,89-102,105-215   pose flags, AABB, SDF consistency, ground KAT, cloud
  tests/test_kdtree.py:6-20                 kd-tree == brute force
  tests/test_constraints.py:148-175         is_valid == all(evaluate > 0); no-SDF/no-pairs case
  cpp/plainmp/ompl/generate_sequence.py:37-39,66-73 and motion_validator.cpp:24-36
skrobot (the reference tests' external oracle) is not installed; the closed-form SDFs below are
written independently in numpy and play its role.
"""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import bridge
from plainmp_b200 import psdf, workloads
from plainmp_b200.constraint import SphereAttachmentSpec, SphereCollisionCst
from plainmp_b200.kinematics import BaseType
from plainmp_b200.robot_spec import PR2RarmSpec, rpy_matrix

GOLDEN = Path(__file__).parent / "golden"


# ---- independent closed-form SDFs (role of skrobot.sdf in tests/test_psdf.py) ----
def np_box(p, width, pos, rot):
    loc = (p - pos) @ rot  # rot^T (p - pos)
    d = np.abs(loc) - 0.5 * np.asarray(width)
    return np.linalg.norm(np.maximum(d, 0), axis=1) + np.minimum(d.max(axis=1), 0)


def np_cyl(p, radius, height, pos, rot):
    loc = (p - pos) @ rot
    d = np.stack([np.hypot(loc[:, 0], loc[:, 1]) - radius, np.abs(loc[:, 2]) - 0.5 * height], axis=1)
    return np.linalg.norm(np.maximum(d, 0), axis=1) + np.minimum(d.max(axis=1), 0)


def np_sphere(p, radius, pos):
    return np.linalg.norm(p - pos, axis=1) - radius


def test_pose_axis_align_flags():
    # tests/test_psdf.py:27-39
    for rot, aa, za in [
        (np.eye(3), True, True),
        (rpy_matrix(0.3, 0.0, 0.0), False, True),
        (rpy_matrix(0.3, 0.3, 0.3), False, False),
    ]:
        o = bridge.OracleSDF.from_psdf(psdf.BoxSDF([1, 1, 1], psdf.Pose(np.zeros(3), rot)))
        _, _, f_aa, f_za = o.pose()
        assert (f_aa, f_za) == (aa, za)


def test_aabb_after_transform():
    # tests/test_psdf.py:89-102
    o = bridge.OracleSDF.from_psdf(psdf.BoxSDF([1, 1, 1]))
    o.translate([1.0, 2.0, 3.0])
    lb, ub = o.aabb()
    np.testing.assert_allclose(lb, [0.5, 1.5, 2.5])
    np.testing.assert_allclose(ub, [1.5, 2.5, 3.5])
    o = bridge.OracleSDF.from_psdf(psdf.BoxSDF([1, 1, 1]))
    o.rotate_z(np.pi * 0.25)
    lb, ub = o.aabb()
    np.testing.assert_allclose(lb, [-0.5 * np.sqrt(2), -0.5 * np.sqrt(2), -0.5])
    np.testing.assert_allclose(ub, [0.5 * np.sqrt(2), 0.5 * np.sqrt(2), 0.5])


def test_ground_kat():
    # tests/test_psdf.py:201-204
    g = bridge.OracleSDF.from_psdf(psdf.GroundSDF(1.0))
    for x in (-2.0, 0.0, 0.7):
        assert g.evaluate([x, 0.0, 0.0]) == 1.0
    lb, ub = g.aabb()  # primitive_sdf.hpp:250-255: pinned at z = 0
    assert lb[2] == 0.0 and ub[2] == 0.0 and np.isinf(lb[0]) and np.isinf(ub[1])


@pytest.mark.parametrize("kind", ["box", "cylinder", "sphere"])
def test_sdf_consistency_with_closed_form(kind):
    # tests/test_psdf.py:105-156 (i=0 identity, i=1 yaw only, else random RPY; then
    # translate + rotate_z; then clone)
    rng = np.random.default_rng({"box": 10, "cylinder": 11, "sphere": 12}[kind])
    for i in range(40):
        trans = rng.standard_normal(3) * 0.5
        if i == 0:
            ypr = (0.0, 0.0, 0.0)
        elif i == 1:
            ypr = (rng.uniform(-np.pi, np.pi), 0.0, 0.0)
        else:
            ypr = tuple(rng.uniform(-np.pi, np.pi, 3))
        rot = rpy_matrix(*ypr)
        pose = psdf.Pose(trans, rot)
        if kind == "box":
            host = psdf.BoxSDF([0.5, 0.3, 0.6], pose)
            ref = lambda p, t, r: np_box(p, [0.5, 0.3, 0.6], t, r)  # noqa: E731
        elif kind == "cylinder":
            host = psdf.CylinderSDF(0.7, 0.2, pose)
            ref = lambda p, t, r: np_cyl(p, 0.7, 0.2, t, r)  # noqa: E731
        else:
            host = psdf.SphereSDF(0.5, pose)
            ref = lambda p, t, r: np_sphere(p, 0.5, t)  # noqa: E731
        o = bridge.OracleSDF.from_psdf(host)
        pts = rng.standard_normal((300, 3))
        radii = np.where(rng.random(300) < 0.5, 0.0, rng.uniform(0, 0.3, 300))
        want = ref(pts, trans, rot)
        np.testing.assert_allclose(o.evaluate_batch(pts), want, rtol=1e-9, atol=1e-12)
        clear = np.abs(want - radii) > 1e-9
        assert (o.is_outside_batch(pts, radii) == (want > radii))[clear].all()
        # translate + rotate_z, then clone
        dt = rng.standard_normal(3) * 0.2
        yaw = rng.uniform(-1, 1)
        o.translate(dt)
        o.rotate_z(yaw)
        rz = rpy_matrix(yaw, 0, 0)
        want2 = ref(pts, trans + dt, rz @ rot)
        np.testing.assert_allclose(o.evaluate_batch(pts), want2, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(o.clone().evaluate_batch(pts), want2, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("kind", ["box", "cylinder", "sphere", "ground", "cloud"])
def test_is_outside_equals_evaluate_gt_r(kind):
    # tests/test_psdf.py:165-169, :205-215
    rng = np.random.default_rng(3)
    pose = psdf.Pose(rng.standard_normal(3) * 0.3, rpy_matrix(*rng.uniform(-1, 1, 3)))
    host = {
        "box": lambda: psdf.BoxSDF([0.5, 0.3, 0.6], pose),
        "cylinder": lambda: psdf.CylinderSDF(0.4, 0.5, pose),
        "sphere": lambda: psdf.SphereSDF(0.5, pose),
        "ground": lambda: psdf.GroundSDF(0.3),
        "cloud": lambda: psdf.CloudSDF(rng.uniform(0, 1, (100, 3)), 0.1),
    }[kind]()
    o = bridge.OracleSDF.from_psdf(host)
    pts = rng.standard_normal((500, 3)) * 1.5
    vals = o.evaluate_batch(pts)
    for r in np.linspace(0, 2, 10):
        clear = np.abs(vals - r) > 1e-9
        assert (o.is_outside_batch(pts, r) == (vals > r))[clear].all()
    assert np.allclose([o.evaluate(p) for p in pts[:20]], vals[:20])


def test_union_is_min_of_members():
    # tests/test_psdf.py:218-254
    rng = np.random.default_rng(5)
    a = psdf.BoxSDF([0.4, 0.4, 0.4], psdf.Pose([0.3, 0, 0], rpy_matrix(0.2, 0.1, 0)))
    b = psdf.SphereSDF(0.3, psdf.Pose([-0.2, 0.1, 0.4]))
    c = psdf.CylinderSDF(0.2, 0.6, psdf.Pose([0, 0.5, 0]))
    u = bridge.OracleSDF.from_psdf(psdf.UnionSDF([a, psdf.UnionSDF([b, c])]))
    pts = rng.standard_normal((400, 3))
    members = [bridge.OracleSDF.from_psdf(x).evaluate_batch(pts) for x in (a, b, c)]
    np.testing.assert_allclose(u.evaluate_batch(pts), np.min(members, axis=0))
    assert (u.is_outside_batch(pts, 0.05) == (np.min(members, axis=0) > 0.05)).all()


def test_kdtree_matches_brute_force():
    # tests/test_kdtree.py:6-20
    rng = np.random.default_rng(0)
    pts = rng.standard_normal((1000, 3))
    tree = bridge.OracleKDTree(pts)
    for q in rng.standard_normal((300, 3)):
        d2 = ((pts - q) ** 2).sum(axis=1)
        k = int(np.argmin(d2))
        np.testing.assert_allclose(tree.query(q), pts[k])
        assert abs(tree.sqdist(q) - d2[k]) < 1e-12


def test_sequence_table_against_reference_generator_output():
    # tests/golden/sequence_table.json is the output of the reference's generate_sequence.py
    gold = json.loads((GOLDEN / "sequence_table.json").read_text())
    assert gold["sha256_of_str_rows"].startswith("1c32d53c") and gold["sha256_of_str_rows"].endswith("60a02d")
    for n in range(1, 129):
        row = bridge.sequence_row(n)
        assert row == gold["rows"][n - 1][:n]
        assert sorted(row) == list(range(n)) and row[0] == 0  # generate_sequence.py:66-73
        if n > 1:
            assert row[1] == n - 1
    assert bridge.sequence_row(20) == [0, 19, 10, 5, 15, 17, 13, 8, 3, 18, 16, 14, 12, 11, 9, 7, 6, 4, 2, 1]
    assert bridge.sequence_row(8) == [0, 7, 4, 2, 6, 5, 3, 1]  # motion_validator.cpp:26


def test_step_ratio_worked_example():
    # motion_validator.cpp:24-36: step_ratio 0.15 -> n_test 8 -> rates 1.0 .6 .3 .9 .75 .45 .15
    v = bridge.OracleValidator(0, 7, np.ones(7) * 0.05)
    s1 = np.zeros(7)
    s2 = np.zeros(7)
    s2[3] = -0.05 / 0.15
    ratio = v.step_ratio(s1, s2)
    assert abs(ratio - 0.15) < 1e-12
    n_test = int(np.floor(1 / ratio)) + 2
    assert n_test == 8
    seq = bridge.sequence_row(n_test)
    rates = [1.0] + [seq[i] * ratio for i in range(2, n_test)]
    np.testing.assert_allclose(rates, [1.0, 0.6, 0.3, 0.9, 0.75, 0.45, 0.15])
    # first maximum wins; tiny motion -> inf -> valid; euclidean = resolution / L2
    s2b = np.zeros(7)
    s2b[1] = s2b[5] = 0.2
    assert abs(v.step_ratio(s1, s2b) - 0.25) < 1e-12
    assert np.isinf(v.step_ratio(s1, s1 + 1e-7))
    ve = bridge.OracleValidator(1, 7, None, 0.1)
    assert abs(ve.step_ratio(s1, np.full(7, 0.3)) - 0.1 / (0.3 * np.sqrt(7))) < 1e-12


def test_approx_sin_cos_error_bounds():
    # SURVEY finding 3 / App. A.2 (measured bounds of the reference polynomial)
    x = np.linspace(-2 * np.pi, 2 * np.pi, 20001)
    sc = np.array([bridge.approx_sin_cos(v) for v in x])
    assert np.abs(sc[:, 0] - np.sin(x)).max() < 4e-6
    assert 1e-5 < np.abs(sc[:, 1] - np.cos(x)).max() < 3e-5


def test_panda_flange_known_pose():
    # public Franka kinematics: at q = 0 the flange (panda_link8) sits at (0.088, 0, 0.926)
    spec, _ = workloads.make_panda()
    kin = spec.get_kin()
    ok = bridge.OracleKin(kin)
    ok.set_joint_angles(kin.get_joint_ids(spec.control_joint_names), np.zeros(7), True)
    pose = ok.link_pose(kin.get_link_ids(["panda_link8"])[0])
    np.testing.assert_allclose(pose[:3], [0.088, 0.0, 0.926], atol=1e-12)
    np.testing.assert_allclose(np.abs(pose[3:]), [1, 0, 0, 0], atol=1e-12)


@pytest.mark.parametrize("base_type", [BaseType.FIXED, BaseType.PLANAR, BaseType.FLOATING])
@pytest.mark.parametrize("with_self", [False, True])
def test_is_valid_equals_all_evaluate_positive(base_type, with_self):
    # tests/test_constraints.py:47-53,148-162 (PR2 right arm + unit box at (0.5, 0.5, 0.5))
    spec = PR2RarmSpec(base_type=base_type)
    cst = spec.create_collision_const(self_collision=with_self, use_cache=False)
    cst.set_sdf(psdf.BoxSDF([1, 1, 1], psdf.Pose([0.5, 0.5, 0.5])))
    o = bridge.OracleCst(cst)
    rng = np.random.default_rng(7)
    n_agree = 0
    for _ in range(100):
        q = rng.standard_normal(o.q_dim)
        if base_type != BaseType.FIXED:
            q[7:] *= 0.1
        vals = o.evaluate(q)
        assert len(vals) == (2 if with_self else 1)
        if np.abs(vals).min() < 1e-4:
            continue  # approximate vs exact FK may disagree on the sign at contact
        assert o.is_valid(q) == bool((vals > 0).all())
        n_agree += 1
    assert n_agree > 90


@pytest.mark.parametrize("base_type", [BaseType.FIXED, BaseType.PLANAR, BaseType.FLOATING])
@pytest.mark.parametrize("with_self", [False, True])
def test_evaluate_jacobian_matches_numerical(base_type, with_self):
    # the reference's own Jacobian check, tests/test_constraints.py:22-44 (eps 1e-7, 4 decimals) on
    # the case of :148-160 (PR2 right arm + unit box, std 1.0 fixed base / 0.1 moving base):
    # pins evaluate_ext_collision / evaluate_self_collision gradients and get_attached_point_jacobian
    spec = PR2RarmSpec(base_type=base_type)
    cst = spec.create_collision_const(self_collision=with_self, use_cache=False)
    cst.set_sdf(psdf.BoxSDF([1, 1, 1], psdf.Pose([0.5, 0.5, 0.5])))
    o = bridge.OracleCst(cst)
    rng = np.random.default_rng(11)
    dim = o.q_dim
    std = 1.0 if base_type == BaseType.FIXED else 0.1
    Q = rng.standard_normal((100, dim)) * std
    vals, jac = o.evaluate_grad_batch(Q)
    eps = 1e-7
    n_rows = 2 if with_self else 1
    n_nonzero = 0
    for i in range(Q.shape[0]):
        Qp = np.repeat(Q[i : i + 1], dim, axis=0) + eps * np.eye(dim)
        vp = o.evaluate_batch(Qp)
        for t in range(2):
            if np.isnan(vals[i, t]):
                assert t == 1 and not with_self
                continue
            num = (vp[:, t] - vals[i, t]) / eps
            np.testing.assert_almost_equal(jac[i, t], num, decimal=4)
            n_nonzero += int(np.abs(jac[i, t]).max() > 0)
    assert n_nonzero > 20 * n_rows  # most samples are closer than the 0.1 cutoff, so the rows are not trivially zero


def test_no_sdf_no_pairs_is_always_valid():
    # tests/test_constraints.py:163-175, SURVEY App. C-7
    spec = PR2RarmSpec()
    cst = spec.create_collision_const(self_collision=False, use_cache=False)
    o = bridge.OracleCst(cst)
    rng = np.random.default_rng(8)
    for _ in range(20):
        q = rng.standard_normal(7)
        assert o.is_valid(q)
        assert len(o.evaluate(q)) == 0


def test_ground_aabb_quirk():
    # a sphere collides with Ground only if |c.z| <= r AND c.z + h <= r
    spec = PR2RarmSpec()
    kin = spec.get_kin()
    one = SphereAttachmentSpec("base_link", np.array([[0.0], [0.0], [0.0]]), np.array([0.1]), False)
    for height, base_z, expect_valid in [
        (0.0, 0.05 - 0.051, False),  # touching the plane through z = 0
        (0.0, -1.0 - 0.051, True),   # fully below the plane: reported collision free
        (0.5, -0.45 - 0.051, True),  # plane at z = -0.5, sphere there, but away from z = 0
        (0.5, 0.0 - 0.051, True),    # 0.5 above the plane z = -0.5
        (-0.05, 0.02 - 0.051, False),  # near z = 0 and below the plane z = 0.05
    ]:
        kin.set_base_pose([0, 0, base_z, 0, 0, 0, 1])
        cst = SphereCollisionCst(kin, [], BaseType.FIXED, [one], [], None, True)
        cst.set_sdf(psdf.GroundSDF(height))
        assert bridge.OracleCst(cst).is_valid(np.zeros(0)) == expect_valid
    kin.set_base_pose([0, 0, 0, 0, 0, 0, 1])


def test_oracle_reproduces_committed_vectors():
    # regression pin of the oracle itself (tools/make_golden.py)
    for name, maker in [
        ("panda_dualbars_seed1", workloads.make_panda),
        ("fetch_table_seed2", workloads.make_fetch),
        ("openarm_dual_seed6", workloads.make_openarm),
    ]:
        g = np.load(GOLDEN / f"{name}.npz")
        spec, cst = maker()
        n = min(int(g["n"]), 2048)
        Q = workloads.random_configs(spec, int(g["n"]), int(g["seed"]))[:n].astype(np.float64)
        o = bridge.OracleCst(cst)
        want = np.unpackbits(g["valid_bits"], bitorder="little")[:n].astype(bool)
        assert (o.is_valid_batch(Q) == want).all()
        np.testing.assert_allclose(o.evaluate_batch(Q), g["dist"][:n], atol=1e-6)
        assert (o.sphere_order() == g["sphere_order"]).all()
