"""Pin the restated FP64 oracle (oracle/plainmp_oracle.cpp) against the REFERENCE ITSELF: the
reference's own C++ sources for the hot path, compiled where they lie under /root/reference into
oracle/_ref/libplainmp_ref.so (oracle/Makefile.ref; Eigen and OMPL are replaced by the minimal API
stand-ins of oracle/ref_shim/).  CPU only.

Every comparison here is FP64 vs FP64 and is expected to agree to rounding (1e-12), validity bit
for bit.  What this pins that the KATs cannot: the URDF parser and link/joint numbering
(third/urdf_parser + kinematics.cpp:29-223), lazy FK (algorithm.cpp:88-210), the sphere-group tables
and reorder (primitive_sphere_collision.cpp:53-179), both broad phases and the early exits
(.cpp:181-254), evaluate + Jacobian rows (.cpp:256-399, algorithm.cpp:346-433), every primitive's
is_outside/evaluate/AABB (primitive_sdf.hpp), the kd-tree (kdtree.cpp) and both motion validators
(motion_validator.cpp).

The library is built in the authoring container (it needs /root/reference) and travels to the GPU
box as a prebuilt, git-ignored file; where neither the library nor the reference tree exists the
module is skipped, and the committed tests/golden/*.npz (reference-generated) carry the pin.
"""
import numpy as np
import pytest

from oracle import bridge, ref_bridge
from plainmp_b200 import psdf, workloads
from plainmp_b200.constraint import SphereAttachmentSpec, SphereCollisionCst
from plainmp_b200.kinematics import BaseType, KinematicModel
from plainmp_b200.robot_spec import FetchSpec, OpenArmV10DualSpec, PandaSpec, PR2DualarmSpec, PR2RarmSpec, rpy_matrix

pytestmark = pytest.mark.skipif(
    not (ref_bridge.available() or ref_bridge.can_build()),
    reason="oracle/_ref/libplainmp_ref.so is absent and /root/reference is not mounted",
)

TOL = 1e-12


@pytest.fixture(scope="module", autouse=True)
def _build_ref():
    ref_bridge.build()
    yield


MAKERS = {
    "panda_dualbars": (workloads.make_panda, 1, 4000),
    "panda_ceiled": (lambda: workloads.make_panda(ceiled=True), 3, 2000),
    "fetch_table": (workloads.make_fetch, 2, 4000),
    "pr2_cloud20k": (lambda: workloads.make_pr2_cloud(20000), 4, 1500),
    "pr2_table": (workloads.make_pr2_table, 9, 2000),
    "openarm_dual": (workloads.make_openarm, 6, 2000),
}


@pytest.mark.parametrize("name", list(MAKERS))
def test_is_valid_evaluate_spheres_match_reference(name):
    maker, seed, n = MAKERS[name]
    spec, cst = maker()
    Q = workloads.random_configs(spec, n, seed).astype(np.float64)
    o, r = bridge.OracleCst(cst), ref_bridge.RefCst(cst)
    assert (o.q_dim, o.n_spheres) == (r.q_dim, r.n_spheres)
    v_o, v_r = o.is_valid_batch(Q), r.is_valid_batch(Q)
    assert (v_o == v_r).all(), f"{(v_o != v_r).sum()} validity mismatches"
    assert 0.02 < v_r.mean() < 0.98  # both outcomes exercised
    m = min(n, 600)
    vals_o, jac_o = o.evaluate_grad_batch(Q[:m])
    vals_r, jac_r = r.evaluate_grad_batch(Q[:m])
    assert (np.isnan(vals_o) == np.isnan(vals_r)).all()
    np.testing.assert_allclose(vals_o, vals_r, atol=TOL, rtol=0)
    # the reference differentiates the SDF by forward differences with step 1e-6 (.cpp:305-312): its
    # own rounding noise is ~1e-16 / 1e-6 per unit of Jacobian, both sides execute the same formula
    np.testing.assert_allclose(jac_o, jac_r, atol=1e-8, rtol=0)
    for accurate in (False, True):
        np.testing.assert_allclose(o.spheres_batch(Q[:64], accurate), r.spheres_batch(Q[:64], accurate), atol=TOL, rtol=0)
        g_o = np.stack([o.group_spheres(q, accurate) for q in Q[:16]])
        np.testing.assert_allclose(g_o, r.group_spheres_batch(Q[:16], accurate), atol=TOL, rtol=0)


@pytest.mark.parametrize("base_type", [BaseType.FIXED, BaseType.PLANAR, BaseType.FLOATING])
@pytest.mark.parametrize("with_self", [False, True])
def test_moving_base_matches_reference(base_type, with_self):
    # the reference's own constraint-test case (tests/test_constraints.py:47-53,148-160)
    spec = PR2RarmSpec(base_type=base_type)
    cst = spec.create_collision_const(self_collision=with_self, use_cache=False)
    cst.set_sdf(psdf.BoxSDF([1, 1, 1], psdf.Pose([0.5, 0.5, 0.5])))
    o, r = bridge.OracleCst(cst), ref_bridge.RefCst(cst)
    rng = np.random.default_rng(21)
    Q = rng.standard_normal((400, o.q_dim)) * (1.0 if base_type == BaseType.FIXED else 0.3)
    assert (o.is_valid_batch(Q) == r.is_valid_batch(Q)).all()
    vals_o, jac_o = o.evaluate_grad_batch(Q)
    vals_r, jac_r = r.evaluate_grad_batch(Q)
    assert (np.isnan(vals_o) == np.isnan(vals_r)).all()
    np.testing.assert_allclose(vals_o, vals_r, atol=TOL, rtol=0)
    # base-rotation columns are forward differences with step 1e-7 (algorithm.cpp:392-431)
    np.testing.assert_allclose(jac_o, jac_r, atol=1e-7, rtol=0)
    np.testing.assert_allclose(o.spheres_batch(Q[:32], True), r.spheres_batch(Q[:32], True), atol=TOL, rtol=0)


@pytest.mark.parametrize("Spec", [PandaSpec, FetchSpec, PR2DualarmSpec, OpenArmV10DualSpec])
def test_urdf_parse_and_fk_match_reference(Spec):
    """Same URDF text through the reference's urdf_parser and through the host mirror's parser:
    every link pose, joint limits, approximate and exact FK."""
    spec = Spec()
    kin = spec.get_kin()
    names = spec.control_joint_names
    ids = kin.get_joint_ids(names)
    rk = ref_bridge.RefKin(kin)
    ok = bridge.OracleKin(kin)
    lim = rk.joint_limits(names)
    np.testing.assert_array_equal(lim, np.array(kin.get_joint_position_limits(ids)))
    rng = np.random.default_rng(5)
    link_names = [L.name for L in kin.links]
    for trial in range(6):
        q = rng.uniform(-2.5, 2.5, size=len(names))
        accurate = trial % 2 == 0
        rk.set_joint_positions(names, q, accurate)
        ok.set_joint_angles(ids, q, accurate)
        for lid, ln in enumerate(link_names):
            np.testing.assert_allclose(ok.link_pose(lid), rk.link_pose(ln), atol=TOL, rtol=0, err_msg=ln)


def test_added_links_and_base_pose_match_reference():
    spec = PandaSpec()
    xml = spec.get_kin()._urdf_xml
    kin = KinematicModel(xml)
    kin.add_new_link("tool", "panda_hand", [0.01, -0.02, 0.1], [0.3, -0.2, 0.5], True)
    kin.add_new_link("tool_tip", "tool", [0.0, 0.0, 0.05], [0.0, 0.0, 0.0], True)
    kin.set_base_pose([0.1, 0.2, 0.3, *(np.array([0.1, 0.2, 0.3, 0.9]) / np.linalg.norm([0.1, 0.2, 0.3, 0.9]))])
    names = spec.control_joint_names
    q = np.linspace(-1.0, 1.0, 7)
    kin.set_joint_positions(kin.get_joint_ids(names), q, True)
    rk, ok = ref_bridge.RefKin(kin), bridge.OracleKin(kin)
    for ln in ("tool", "tool_tip", "panda_link8"):
        np.testing.assert_allclose(ok.link_pose(kin.get_link_ids([ln])[0]), rk.link_pose(ln), atol=TOL, rtol=0)
    with pytest.raises(RuntimeError):  # kinematics.cpp:313-320
        rk.add_new_link("tool", "panda_hand", [0, 0, 0], [0, 0, 0], True)
    with pytest.raises(ValueError):  # std::invalid_argument, kinematics.cpp:270-272
        rk.link_pose("no_such_link")
    with pytest.raises(ValueError):  # kinematics.cpp:243-245
        rk.set_joint_positions(["no_such_joint"], [0.0])


def _random_prims(rng):
    out = []
    for k in range(24):
        pos = rng.uniform(-0.5, 0.5, 3)
        mode = k % 3
        rot = np.eye(3) if mode == 0 else (rpy_matrix(0, 0, rng.uniform(-3, 3)) if mode == 1 else rpy_matrix(*rng.uniform(-3, 3, 3)))
        kind = (k // 3) % 3
        if kind == 0:
            out.append(psdf.BoxSDF(rng.uniform(0.1, 0.8, 3), psdf.Pose(pos, rot)))
        elif kind == 1:
            out.append(psdf.CylinderSDF(rng.uniform(0.05, 0.4), rng.uniform(0.1, 0.9), psdf.Pose(pos, rot)))
        else:
            out.append(psdf.SphereSDF(rng.uniform(0.05, 0.5), psdf.Pose(pos, rot)))
    out.append(psdf.GroundSDF(0.0))
    out.append(psdf.GroundSDF(0.37))
    return out


def test_primitive_sdfs_match_reference():
    rng = np.random.default_rng(31)
    P = rng.uniform(-1.2, 1.2, size=(3000, 3))
    radii = np.where(rng.random(3000) < 0.2, 0.0, rng.uniform(0.0, 0.3, 3000))
    prims = _random_prims(rng)
    for s in prims:
        o, r = bridge.OracleSDF.from_psdf(s), ref_bridge.RefSDF.from_psdf(s)
        np.testing.assert_allclose(o.evaluate_batch(P), r.evaluate_batch(P), atol=TOL, rtol=0)
        np.testing.assert_allclose(r.evaluate_batch(P), r.evaluate_batch(P, each=True), atol=TOL, rtol=0)
        assert (o.is_outside_batch(P, radii) == r.is_outside_batch(P, radii)).all()
        for a, b in zip(o.aabb(), r.aabb()):
            np.testing.assert_allclose(a, b, atol=TOL, rtol=0)
        if s.get_type() != psdf.T_GROUND:
            po, pr = o.pose(), r.pose()
            assert po[2:] == pr[2:]
            np.testing.assert_allclose(po[1], pr[1], atol=TOL, rtol=0)
    u = psdf.UnionSDF(prims[:7])
    o, r = bridge.OracleSDF.from_psdf(u), ref_bridge.RefSDF.from_psdf(u)
    np.testing.assert_allclose(o.evaluate_batch(P), r.evaluate_batch(P), atol=TOL, rtol=0)
    assert (o.is_outside_batch(P, radii) == r.is_outside_batch(P, radii)).all()
    # scene mutation: translate / rotate_z on the union, clone independence
    for obj in (o, r):
        obj.translate([0.1, -0.2, 0.05])
        obj.rotate_z(0.7)
    np.testing.assert_allclose(o.evaluate_batch(P), r.evaluate_batch(P), atol=TOL, rtol=0)
    oc, rc = o.clone(), r.clone()
    o.translate([1.0, 0, 0])
    r.translate([1.0, 0, 0])
    np.testing.assert_allclose(oc.evaluate_batch(P), rc.evaluate_batch(P), atol=TOL, rtol=0)
    np.testing.assert_allclose(o.evaluate_batch(P), r.evaluate_batch(P), atol=TOL, rtol=0)


def test_reference_error_behaviour_matches_host_mirror():
    g = ref_bridge.RefSDF.from_psdf(psdf.GroundSDF(0.0))
    with pytest.raises(RuntimeError):  # primitive_sdf.hpp:234-240
        g.translate([0, 0, 1])
    with pytest.raises(RuntimeError):
        g.rotate_z(0.1)
    c = ref_bridge.RefSDF.from_psdf(psdf.CloudSDF(np.random.default_rng(0).random((10, 3)), 0.01))
    for call in (lambda: c.translate([0, 0, 1]), lambda: c.rotate_z(0.1), c.clone):  # :585-595
        with pytest.raises(RuntimeError):
            call()
    with pytest.raises(RuntimeError):
        psdf.GroundSDF(0.0).translate([0, 0, 1])
    with pytest.raises(RuntimeError):
        psdf.CloudSDF(np.random.default_rng(0).random((10, 3)), 0.01).clone()
    # bad pair name -> runtime_error (.cpp:171-174); unknown link in a spec -> invalid_argument
    spec = PandaSpec()
    kin = spec.get_kin()
    one = SphereAttachmentSpec("panda_link1", np.zeros((3, 1)), np.array([0.1]), False)
    good = SphereCollisionCst(kin, spec.control_joint_names, BaseType.FIXED, [one], [], None, True)
    ref_bridge.RefCst(good)
    good._group_link_names = ["nope"]  # the reference resolves the name itself: get_link_ids throws
    with pytest.raises(ValueError):
        ref_bridge.RefCst(good)
    with pytest.raises(RuntimeError):
        SphereCollisionCst(kin, spec.control_joint_names, BaseType.FIXED, [one], [("panda_link1", "nope")], None, True)
    with pytest.raises(ValueError):
        SphereCollisionCst(kin, spec.control_joint_names, BaseType.FIXED,
                           [SphereAttachmentSpec("nope", np.zeros((3, 1)), np.array([0.1]), False)], [], None, True)


def test_kdtree_and_cloud_match_reference():
    rng = np.random.default_rng(41)
    for n in (1, 2, 7, 1000, 20000):
        pts = rng.uniform(-1, 1, size=(n, 3))
        Pq = rng.uniform(-1.5, 1.5, size=(500, 3))
        o, r = bridge.OracleKDTree(pts), ref_bridge.RefKDTree(pts)
        sq_r = r.sqdist_batch(Pq)
        np.testing.assert_allclose(o.sqdist_batch(Pq), sq_r, atol=TOL, rtol=0)
        brute = ((Pq[:, None, :] - pts[None, :, :]) ** 2).sum(-1).min(1)
        np.testing.assert_allclose(sq_r, brute, atol=1e-12, rtol=0)  # tests/test_kdtree.py:6-20
        nn = r.query_batch(Pq)
        np.testing.assert_allclose(((nn - Pq) ** 2).sum(1), sq_r, atol=1e-12, rtol=0)
    cloud = psdf.CloudSDF(workloads.table_cloud(5000), 0.002)
    o, r = bridge.OracleSDF.from_psdf(cloud), ref_bridge.RefSDF.from_psdf(cloud)
    P = rng.uniform([0.2, -1.1, 0.6], [1.4, 1.1, 1.3], size=(2000, 3))
    np.testing.assert_allclose(o.evaluate_batch(P), r.evaluate_batch(P), atol=TOL, rtol=0)
    rad = rng.uniform(0, 0.1, 2000)
    assert (o.is_outside_batch(P, rad) == r.is_outside_batch(P, rad)).all()
    for a, b in zip(o.aabb(), r.aabb()):
        np.testing.assert_allclose(a, b, atol=TOL, rtol=0)


@pytest.mark.parametrize("vtype", [0, 1])
def test_motion_validators_match_reference(vtype):
    """checkMotion compiled from ompl/motion_validator.cpp (over the ompl::base stand-in): validity
    AND the number of is_valid calls, i.e. the waypoint order with its early exit."""
    spec, cst = workloads.make_panda(ceiled=True)
    o, r = bridge.OracleCst(cst), ref_bridge.RefCst(cst)
    widths = np.array([0.05, 0.07, 0.05, 0.1, 0.05, 0.2, 0.05])
    res = 0.1
    ov = bridge.OracleValidator(vtype, 7, widths, res)
    for planner_like, n in ((False, 600), (True, 600)):
        q0, q1 = workloads.random_edges(spec, n, 3, planner_like, (lambda q: o.is_valid_batch(q.astype(np.float64))) if planner_like else None)
        q0, q1 = q0.astype(np.float64), q1.astype(np.float64)
        want, n_ref = r.check_motion_batch(vtype, widths, res, q0, q1)
        got, n_orc = ov.check_motion_batch(o, q0, q1)
        assert (got == want).all()
        assert n_orc == n_ref
        for i in range(0, n, 37):
            assert ov.step_ratio(q0[i], q1[i]) == ref_bridge.step_ratio(vtype, widths, res, q0[i], q1[i])
    # degenerate edges: identical endpoints (ratio inf -> valid without any check), very long edges
    # (n_test >= 129 -> natural order), tiny box widths
    z = np.zeros((1, 7))
    for s1, s2, w in [(z, z, widths), (z, z + 1e-7, widths), (z - 2.0, z + 2.0, np.full(7, 0.01)), (z, z + 0.3, np.full(7, 0.3))]:
        ovv = bridge.OracleValidator(vtype, 7, w, 0.01)
        want, n_ref = r.check_motion_batch(vtype, w, 0.01, s1, s2)
        got, n_orc = ovv.check_motion_batch(o, s1, s2)
        assert (got == want).all() and n_orc == n_ref


def test_fetch_long_edges_natural_order_match_reference():
    # fetch has continuous joints clipped to +-2 pi: n_test up to 253 > 128 -> untabulated order
    spec, cst = workloads.make_fetch()
    o, r = bridge.OracleCst(cst), ref_bridge.RefCst(cst)
    q0, q1 = workloads.random_edges(spec, 300, 12, False)
    q0, q1 = q0.astype(np.float64), q1.astype(np.float64)
    w = np.full(8, 0.05)
    want, n_ref = r.check_motion_batch(0, w, 0.0, q0, q1)
    got, n_orc = bridge.OracleValidator(0, 8, w).check_motion_batch(o, q0, q1)
    assert (got == want).all() and n_orc == n_ref


def test_zero_radius_self_only_and_no_scene_match_reference():
    spec = PR2RarmSpec()
    kin = spec.get_kin()
    rng = np.random.default_rng(3)
    specs = [
        SphereAttachmentSpec("r_gripper_palm_link", rng.uniform(-0.1, 0.1, (3, 5)), np.zeros(5), False),
        SphereAttachmentSpec("r_forearm_link", rng.uniform(-0.1, 0.1, (3, 4)), rng.uniform(0.02, 0.08, 4), False),
        SphereAttachmentSpec("base_link", rng.uniform(-0.3, 0.3, (3, 6)), rng.uniform(0.05, 0.2, 6), True),
    ]
    pairs = [("r_gripper_palm_link", "base_link"), ("r_forearm_link", "base_link")]
    for reorder in (True, False):
        cst = SphereCollisionCst(kin, spec.control_joint_names, BaseType.FIXED, specs, pairs, None, reorder)
        Q = rng.standard_normal((500, 7))
        for sdf in (None, psdf.UnionSDF([psdf.BoxSDF([0.4, 0.4, 0.4], psdf.Pose([0.6, -0.2, 0.8], rpy_matrix(0.2, 0.3, 0.4))), psdf.SphereSDF(0.2, psdf.Pose([0.5, 0.3, 1.0]))])):
            cst.set_sdf(sdf)
            o, r = bridge.OracleCst(cst), ref_bridge.RefCst(cst)
            assert (o.is_valid_batch(Q) == r.is_valid_batch(Q)).all()
            vo, vr = o.evaluate_batch(Q), r.evaluate_batch(Q)
            assert (np.isnan(vo) == np.isnan(vr)).all()
            np.testing.assert_allclose(vo, vr, atol=TOL, rtol=0)
            np.testing.assert_allclose(o.spheres_batch(Q[:8], True), r.spheres_batch(Q[:8], True), atol=TOL, rtol=0)


# ---- reference objects -> C-ABI tables (include/plainmp_b200_plainmp.hpp) -----------------------------
def _canon_ops(ops, names=None):
    """Order-independent view of an op table: each op with its chain of ancestors' content."""
    rows = []
    for k in range(len(ops)):
        chain, i = [], k
        while i >= 0:
            o = ops[i]
            chain.append((tuple(np.round(o["origin_quat"], 6)), tuple(np.round(o["origin_xyz"], 6)), int(o["kind"]),
                          tuple(np.round(o["axis"], 6)), int(o["q_index"]), int(o["flags"]) & 1))
            i = int(o["parent"])
        rows.append(tuple(chain))
    return rows


@pytest.mark.parametrize("name", ["panda_dualbars", "fetch_table", "pr2_cloud20k", "openarm_dual"])
def test_reference_side_flatten_equals_host_mirror_tables(name):
    """The C++ flatten a plainmp maintainer would compile into the reference (run here on the compiled
    reference's own SphereCollisionCst / KinematicModel / SDF objects) produces the same C-ABI tables
    as this repo's Python host mirror: same ops (per link chain), groups, spheres (reference's
    re-ordered order), pairs and primitive records."""
    from plainmp_b200 import tables

    maker, seed, n = MAKERS[name]
    spec, cst = maker()
    flat = ref_bridge.RefCst(cst).flatten()
    prog = tables.LinkProgram(cst.get_kin(), cst._control_joint_ids, cst._group_links)
    assert flat["dof"] == len(cst._control_joint_ids) and flat["base_type"] == int(cst._base_type)
    assert len(flat["ops"]) == len(prog.ops)
    assert sorted(_canon_ops(flat["ops"])) == sorted(_canon_ops(prog.ops))
    # groups: link chain of the op + centre + radius + flag + member spheres in order
    ref_chain, host_chain = _canon_ops(flat["ops"]), _canon_ops(prog.ops)
    assert len(flat["groups"]) == len(cst._group_links)
    for g, G in enumerate(flat["groups"]):
        host_op = prog.op_of_link[cst._group_links[g]]
        assert ref_chain[int(G["op"])] == host_chain[host_op]
        centre, radius = tables.group_bounding_sphere(cst._group_spheres[g])
        np.testing.assert_allclose(G["center"], centre, atol=1e-7)
        assert abs(float(G["radius"]) - radius) < 1e-7
        assert bool(G["only_self_collision"]) == cst._group_only_self[g]
        b, c = int(G["sphere_begin"]), int(G["sphere_count"])
        np.testing.assert_allclose(flat["spheres"][b : b + c], cst._group_spheres[g].astype(np.float32), atol=1e-7)
    assert [tuple(p) for p in flat["pairs"]] == [tuple(p) for p in cst._pairs]
    prims, cloud = tables.flatten_primitives(cst._leaves())
    assert len(flat["prims"]) == len(prims)
    for a, b in zip(flat["prims"], prims):
        for k in ("type", "flags"):
            assert int(a[k]) == int(b[k]), k
        for k in ("position", "rot_col0", "rot_col1", "rot_col2", "param0", "param1", "param2", "lb", "ub"):
            np.testing.assert_allclose(a[k], b[k], atol=1e-6, err_msg=k)
    if cloud is not None:
        assert flat["cloud_radius"] == pytest.approx(cloud.radius)
        np.testing.assert_allclose(np.sort(flat["cloud_points"], axis=0), np.sort(cloud.points.astype(np.float32), axis=0), atol=0)
    else:
        assert len(flat["cloud_points"]) == 0


# ---- IneqCompositeCst and the per-edge call counts --------------------------------------------------
def _composite_case(moving_base=False):
    from plainmp_b200.constraint import IneqCompositeCst

    if moving_base:
        spec = PR2RarmSpec(base_type=BaseType.PLANAR)
        a = spec.create_collision_const(self_collision=True, use_cache=False)
        b = spec.create_collision_const(self_collision=False, use_cache=False)
        a.set_sdf(psdf.BoxSDF([1, 1, 1], psdf.Pose([0.9, 0.5, 0.5])))
        b.set_sdf(psdf.UnionSDF([psdf.CylinderSDF(0.2, 1.0, psdf.Pose([-0.6, 0.2, 0.6])), psdf.GroundSDF(0.0)]))
    else:
        spec, a = workloads.make_panda()
        b = spec.create_collision_const(self_collision=False, use_cache=False)
        b.set_sdf(psdf.UnionSDF([psdf.SphereSDF(0.25, psdf.Pose([0.45, 0.1, 0.55])),
                                 psdf.BoxSDF([0.3, 0.3, 0.3], psdf.Pose([-0.5, 0.3, 0.4], rpy_matrix(0.3, 0.2, 0.1)))]))
    c = spec.create_collision_const(self_collision=True, use_cache=False)  # self-collision only: no scene at all
    return spec, IneqCompositeCst([a, b, c])


@pytest.mark.parametrize("moving_base", [False, True])
def test_composite_constraint_matches_reference(moving_base):
    """IneqCompositeCst::is_valid / evaluate (constraints/composite_constraint.hpp:39-61, 98-105): ONE
    update_kintree through the first member, AND of is_valid_dirty, rows stacked in member order."""
    spec, comp = _composite_case(moving_base)
    rc, oc = ref_bridge.RefComposite(comp), bridge.OracleComposite(comp, n_threads=4)
    rng = np.random.default_rng(31)
    if moving_base:
        Q = rng.standard_normal((3000, oc.q_dim))
        Q[:, 7:] *= 0.4
    else:
        Q = workloads.random_configs(spec, 3000, 31).astype(np.float64)
    v_r, v_o = rc.is_valid_batch(Q), oc.is_valid_batch(Q)
    assert (v_r == v_o).all() and 0.05 < v_r.mean() < 0.95
    # the composite is the AND of its members (each driven by the same q)
    members = np.all([ref_bridge.RefCst(c).is_valid_batch(Q) for c in comp.constraints], axis=0)
    assert (v_r == members).all()
    vals_r, jac_r = rc.evaluate_batch(Q[:400], with_jac=True)
    vals_o = oc.evaluate_batch(Q[:400]).reshape(400, -1)
    on = ~np.isnan(vals_o[0])
    assert on.sum() == rc.dim == comp.cst_dim()
    np.testing.assert_allclose(vals_o[:, on], vals_r, atol=TOL, rtol=0)
    # ctor :27-37: members on different kinematic models are refused (a second spec owns a second model)
    from plainmp_b200.constraint import IneqCompositeCst

    other_spec = type(spec)(base_type=spec.base_type) if moving_base else type(spec)()
    other = other_spec.create_collision_const(use_cache=False)
    assert other.get_kin() is not comp.constraints[0].get_kin()
    with pytest.raises(RuntimeError):
        IneqCompositeCst([comp.constraints[0], other])


@pytest.mark.parametrize("vtype", [0, 1])
def test_per_edge_call_counts_match_reference(vtype):
    """checkMotion's is_valid call count PER EDGE (the position of the first failing waypoint in the
    visiting order of motion_validator.cpp:45-78): restatement == compiled reference, edge for edge."""
    spec, cst = workloads.make_panda(ceiled=True)
    o = bridge.OracleCst(cst, n_threads=4)
    w = np.array([0.05, 0.07, 0.05, 0.1, 0.05, 0.2, 0.05])
    for planner_like in (False, True):
        q0, q1 = workloads.random_edges(spec, 3000, 77, planner_like, (lambda q: o.is_valid_batch(q.astype(np.float64))) if planner_like else None)
        q0, q1 = q0.astype(np.float64), q1.astype(np.float64)
        v_r, c_r = ref_bridge.RefCst(cst).check_motion_batch_ex(vtype, w, 0.08, q0, q1)
        v_o, c_o, m = bridge.OracleValidator(vtype, 7, w, 0.08).check_motion_batch_ex(o, q0, q1, 4)
        assert (v_r == v_o).all() and (c_r == c_o).all()
        assert (c_r >= 1).all() and np.isfinite(m).all() and (m >= 0).all()
        v_t, c_t = ref_bridge.check_motion_batch_threads(cst, vtype, w, 0.08, q0, q1, 3)
        assert (v_t == v_r).all() and (c_t == c_r).all()


# ---- round 2: the specs and scene shapes the reference ships that round 1 left out ------------------
JAXON_BASE = ([-0.5, -1.0, 0.7, -0.3, -0.3, -1.0], [1.2, 1.0, 1.3, 0.3, 0.3, 1.0])
PR2_BASE = ([-2.0, -2.5, -3.2], [2.5, 2.5, 3.2])


def _moving_case(name):
    if name == "jaxon":
        spec, cst = workloads.make_jaxon()
        return spec, cst, workloads.random_configs_moving_base(spec, 1500, 1, *JAXON_BASE)
    if name == "jaxon_no_gripper":
        spec, cst = workloads.make_jaxon(gripper_collision=False)
        return spec, cst, workloads.random_configs_moving_base(spec, 1500, 2, *JAXON_BASE)
    spec, cst = workloads.make_pr2_base_only()
    return spec, cst, workloads.random_configs_moving_base(spec, 1500, 3, *PR2_BASE)


@pytest.mark.parametrize("name", ["jaxon", "jaxon_no_gripper", "pr2_base_only"])
def test_jaxon_and_base_only_specs_match_reference(name):
    """JaxonSpec (robot_spec.py:841-933: FLOATING base, 37 entries, foot / gripper links filtered) and
    PR2BaseOnlySpec (:720-732: no controlled joint, q = x, y, yaw)."""
    spec, cst, Q = _moving_case(name)
    Q = Q.astype(np.float64)
    o, r = bridge.OracleCst(cst), ref_bridge.RefCst(cst)
    assert (o.q_dim, o.n_spheres) == (r.q_dim, r.n_spheres)
    assert o.q_dim == {"jaxon": 37, "jaxon_no_gripper": 37, "pr2_base_only": 3}[name]
    assert o.n_spheres == {"jaxon": 73, "jaxon_no_gripper": 41, "pr2_base_only": 161}[name]
    v_o, v_r = o.is_valid_batch(Q), r.is_valid_batch(Q)
    assert (v_o == v_r).all() and 0.05 < v_r.mean() < 0.95
    vals_o, jac_o = o.evaluate_grad_batch(Q[:300])
    vals_r, jac_r = r.evaluate_grad_batch(Q[:300])
    assert (np.isnan(vals_o) == np.isnan(vals_r)).all()
    np.testing.assert_allclose(vals_o, vals_r, atol=TOL, rtol=0)
    np.testing.assert_allclose(jac_o, jac_r, atol=1e-7, rtol=0)
    np.testing.assert_allclose(o.spheres_batch(Q[:32], True), r.spheres_batch(Q[:32], True), atol=TOL, rtol=0)


def test_jaxon_spec_surface():
    from plainmp_b200.robot_spec import JaxonSpec, PR2BaseOnlySpec

    spec = JaxonSpec()
    lb, ub = spec.angle_bounds()
    assert lb.shape == ub.shape == (37,) and (lb < ub).all()
    q = spec.reset_manip_pose_q
    assert q.shape == (37,) and (q >= lb - 1e-9).all() and (q <= ub + 1e-9).all()
    links = {s.parent_link_name for s in spec.parse_sphere_specs()}
    assert "RLEG_LINK5" not in links and "RARM_LINK7" in links
    assert "RARM_LINK7" not in {s.parent_link_name for s in JaxonSpec(gripper_collision=False).parse_sphere_specs()}
    for name in ("rarm_end_coords", "lleg_end_coords"):  # conf/jaxon.yaml custom_links
        spec.get_kin().get_link_ids([name])
    base = PR2BaseOnlySpec()
    assert base.base_type == BaseType.PLANAR and base.control_joint_names == []
    assert base.q_default.shape == (0,)


@pytest.mark.parametrize("n_grid", [6, 4])
def test_attached_box_lattice_matches_reference(n_grid):
    """create_attached_box_collision_const (robot_spec.py:561-595): one group of zero-radius spheres on
    the surface shell of a grasped box (the r < 1e-6 shortcuts of every primitive)."""
    spec = PandaSpec()
    rot = rpy_matrix(0.7, 0.2, -0.4)
    box = psdf.BoxSDF([0.1, 0.06, 0.2], psdf.Pose([0.3, 0.1, 0.5], rot))
    rel = np.array([0.0, 0.0, 0.15])
    cst = spec.create_attached_box_collision_const(box, "panda_hand", rel, n_grid=n_grid)
    cst.set_sdf(workloads.panda_dual_bars_sdf())
    # the lattice the reference builds, restated with its own formulas: meshgrid -> world -> keep sdf > -1e-2
    ext = box.width
    grid = np.meshgrid(*[np.linspace(-0.5 * e, 0.5 * e, n_grid) for e in ext])
    pts = np.stack([g.flatten() for g in grid], axis=1) @ rot.T + box.pose.position
    sd = bridge.OracleSDF.from_psdf(box).evaluate_batch(pts)
    want = pts[sd > -1e-2] - box.pose.position + rel
    got = cst._group_spheres[0][:, :3] if hasattr(cst, "_group_spheres") else None
    o, r = bridge.OracleCst(cst), ref_bridge.RefCst(cst)
    assert o.n_spheres == r.n_spheres == len(want) and len(want) < n_grid ** 3
    if got is not None:  # same point set (the constructor may reorder the spheres of a group)
        key = lambda a: a[np.lexsort(np.round(a, 9).T)]
        np.testing.assert_allclose(key(np.asarray(got, dtype=np.float64)), key(want), atol=1e-12)
    Q = workloads.random_configs(spec, 3000, 2).astype(np.float64)
    v_o, v_r = o.is_valid_batch(Q), r.is_valid_batch(Q)
    assert (v_o == v_r).all() and 0.2 < v_r.mean() < 0.98
    np.testing.assert_allclose(o.evaluate_batch(Q[:300]), r.evaluate_batch(Q[:300]), atol=TOL, rtol=0)


def _pr2_several_clouds(radii=(0.002, 0.002, 0.002)):
    spec = PR2DualarmSpec()
    spec.reflect_default_joint_positions()
    cst = spec.create_collision_const(use_cache=False)
    pts = workloads.table_cloud(6000, 4)
    shelf = np.random.default_rng(0).uniform([-0.2, -1.0, 0.2], [0.4, -0.5, 1.5], size=(2000, 3))
    cst.set_sdf(
        psdf.UnionSDF(
            [
                psdf.CloudSDF(pts[:3000], radii[0]),
                psdf.BoxSDF([0.3, 0.3, 0.3], psdf.Pose([0.6, 0.6, 1.2])),
                psdf.UnionSDF([psdf.CloudSDF(pts[3000:], radii[1]), psdf.CloudSDF(shelf, radii[2])]),
            ]
        )
    )
    return spec, cst


def test_several_clouds_in_one_scene_match_reference():
    """The reference keeps every CloudSDF leaf (primitive_sphere_collision.cpp:436-448); the host mirror
    merges leaves of equal radius into one point set — same validity, same values."""
    from plainmp_b200 import tables

    spec, cst = _pr2_several_clouds()
    o, r = bridge.OracleCst(cst), ref_bridge.RefCst(cst)
    Q = workloads.random_configs(spec, 1500, 1).astype(np.float64)
    v_o, v_r = o.is_valid_batch(Q), r.is_valid_batch(Q)
    assert (v_o == v_r).all() and 0.05 < v_r.mean() < 0.95
    np.testing.assert_allclose(o.evaluate_batch(Q[:200]), r.evaluate_batch(Q[:200]), atol=TOL, rtol=0)
    prims, cloud = tables.flatten_primitives(cst._leaves())
    assert len(prims) == len(cst._leaves()) - 3 and cloud.points.shape == (8000, 3) and cloud.radius == 0.002
    # a single merged CloudSDF gives the reference the same answers: the merge is exact
    spec2 = PR2DualarmSpec()
    spec2.reflect_default_joint_positions()
    cst2 = spec2.create_collision_const(use_cache=False)
    cst2.set_sdf(psdf.UnionSDF([psdf.BoxSDF([0.3, 0.3, 0.3], psdf.Pose([0.6, 0.6, 1.2])), psdf.CloudSDF(cloud.points, 0.002)]))
    r2 = ref_bridge.RefCst(cst2)
    assert (r2.is_valid_batch(Q) == v_r).all()
    np.testing.assert_allclose(r2.evaluate_batch(Q[:200]), r.evaluate_batch(Q[:200]), atol=TOL, rtol=0)


def test_position_only_link_with_spheres_at_its_origin_matches_reference():
    from plainmp_b200 import tables

    spec, cst = workloads.make_panda_position_only_link()
    o, r = bridge.OracleCst(cst), ref_bridge.RefCst(cst)
    Q = workloads.random_configs(spec, 3000, 8).astype(np.float64)
    v_o, v_r = o.is_valid_batch(Q), r.is_valid_batch(Q)
    assert (v_o == v_r).all() and 0.1 < v_r.mean() < 0.95
    np.testing.assert_allclose(o.evaluate_batch(Q[:300]), r.evaluate_batch(Q[:300]), atol=TOL, rtol=0)
    np.testing.assert_allclose(o.spheres_batch(Q[:32], False), r.spheres_batch(Q[:32], False), atol=TOL, rtol=0)
    # flattened for the device: the group sits on the parent link at the added link's local offset
    a = tables.RobotTable.build_arrays(cst.get_kin(), cst._control_joint_ids, cst._base_type, cst._group_links, cst._group_spheres, cst._group_only_self, cst._pairs)
    g = [i for i, l in enumerate(cst._group_links) if cst.get_kin().links[l].name == "tool_tip"][0]
    hand = cst.get_kin().get_link_ids(["panda_hand"])[0]
    assert a["groups"][g]["op"] == a["program"].op_of_link[hand]
    b = a["groups"][g]["sphere_begin"]
    np.testing.assert_allclose(a["spheres"][b : b + 2, :3], [[0.02, -0.01, 0.22]] * 2, atol=1e-7)
    with pytest.raises(NotImplementedError):
        _, bad = workloads.make_panda_position_only_link(offset_spheres=True)
        tables.RobotTable.build_arrays(bad.get_kin(), bad._control_joint_ids, bad._base_type, bad._group_links, bad._group_spheres, bad._group_only_self, bad._pairs)


def test_clouds_of_different_radii_match_reference():
    """CloudSDF leaves of different radii in one scene: the oracle and the reference keep them as separate leaves;
    the host mirror hands them to the device as one point set with per-point radii (radius classes)."""
    from plainmp_b200 import tables

    spec, cst = _pr2_several_clouds(radii=(0.002, 0.03, 0.011))
    o, r = bridge.OracleCst(cst), ref_bridge.RefCst(cst)
    Q = workloads.random_configs(spec, 1500, 1).astype(np.float64)
    v_o, v_r = o.is_valid_batch(Q), r.is_valid_batch(Q)
    assert (v_o == v_r).all() and 0.03 < v_r.mean() < 0.95
    np.testing.assert_allclose(o.evaluate_batch(Q[:200]), r.evaluate_batch(Q[:200]), atol=TOL, rtol=0)
    prims, cloud = tables.flatten_primitives(cst._leaves())
    assert cloud.points.shape == (8000, 3) and cloud.radius == 0.03
    pr = cloud.point_radii
    assert pr.shape == (8000,) and set(np.unique(pr)) == {np.float32(0.002), np.float32(0.03), np.float32(0.011)}
    assert (pr[:3000] == np.float32(0.002)).all() and (pr[3000:6000] == np.float32(0.03)).all()
    five = [psdf.CloudSDF(np.random.default_rng(i).uniform(0, 1, (10, 3)), 0.01 * (i + 1)) for i in range(5)]
    with pytest.raises(NotImplementedError):
        tables.flatten_primitives(five)


def test_reference_side_flatten_of_several_clouds():
    """include/plainmp_b200_plainmp.hpp on the reference's own objects: every CloudSDF leaf of all_sdfs_cache_ goes
    into one point set; per-point radii are handed over exactly when the leaves' radii differ."""
    for radii, expect_classes in (((0.002, 0.002, 0.002), None), ((0.002, 0.03, 0.011), {0.002, 0.03, 0.011})):
        spec, cst = _pr2_several_clouds(radii=radii)
        flat = ref_bridge.RefCst(cst).flatten()
        assert flat["cloud_points"].shape == (8000, 3) and abs(flat["cloud_radius"] - max(radii)) < 1e-9
        if expect_classes is None:
            assert flat["cloud_point_radii"] is None
        else:
            pr = flat["cloud_point_radii"]
            assert pr.shape == (8000,) and {round(float(x), 6) for x in np.unique(pr)} == expect_classes
            # the kd-tree permutes the points of a leaf, never across leaves: 3000 + 3000 + 2000, in leaf order
            assert (pr[:3000] == np.float32(0.002)).all() and (pr[3000:6000] == np.float32(0.03)).all() and (pr[6000:] == np.float32(0.011)).all()
        # same point set as the host mirror's merge (order inside a leaf differs: the reference stores kd-tree order)
        from plainmp_b200 import tables

        _, cloud = tables.flatten_primitives(cst._leaves())
        key = lambda a: a[np.lexsort(a.T)]  # noqa: E731
        np.testing.assert_array_equal(key(flat["cloud_points"]), key(cloud.points.astype(np.float32)))
