"""CPU-only tests of the host side: URDF/table flattening, psdf host objects vs the oracle,
the C-ABI library surface (load + symbols, no compute), error behaviour, bit packing."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from oracle import bridge
from plainmp_b200 import _capi, psdf, tables, workloads
from plainmp_b200.constraint import SphereAttachmentSpec, SphereCollisionCst
from plainmp_b200.kinematics import BaseType, KinematicModel
from plainmp_b200.ompl import ValidatorConfig, ValidatorType, step_ratio
from plainmp_b200.robot_spec import FetchSpec, OpenArmV10DualSpec, PandaSpec, PR2DualarmSpec, rpy_matrix

ROOT = Path(__file__).resolve().parent.parent


def test_capi_library_loads_and_exports_every_declared_symbol():
    lib = _capi.load()
    header = (ROOT / "include" / "plainmp_b200.h").read_text()
    declared = set(re.findall(r"\b(pmp_[a-z0-9_]+)\s*\(", header))
    declared -= {"pmp_robot_t", "pmp_scene_t"}
    assert declared == set(_capi.SYMBOLS), declared ^ set(_capi.SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.pmp_abi_version() == 1
    assert isinstance(lib.pmp_launch_count(), int)


def test_capi_argument_errors_without_gpu():
    lib = _capi.load()
    h = C.c_void_p()
    assert lib.pmp_robot_create(None, C.byref(h)) == _capi.PMP_ERR_INVALID
    assert b"null" in lib.pmp_last_error()
    d = _capi.RobotDesc()
    d.base_type = 7
    assert lib.pmp_robot_create(C.byref(d), C.byref(h)) == _capi.PMP_ERR_INVALID
    ops = np.zeros(2, dtype=_capi.LINK_OP_DTYPE)
    ops[0]["parent"] = 1  # not topological
    d = _capi.RobotDesc()
    d.n_ops = 2
    d.ops = ops.ctypes.data
    assert lib.pmp_robot_create(C.byref(d), C.byref(h)) == _capi.PMP_ERR_INVALID
    assert b"topological" in lib.pmp_last_error()
    assert lib.pmp_check_configs(None, None, None, 4, None, None, None) == _capi.PMP_ERR_INVALID


def test_no_cpu_fallback_without_device():
    lib = _capi.load()
    if lib.pmp_device_count() > 0:
        pytest.skip("this box has a GPU")
    spec, cst = workloads.make_panda()
    with pytest.raises(_capi.PmpError):
        cst.is_valid(np.zeros(7))
    with pytest.raises(_capi.PmpError):
        psdf.BoxSDF([1, 1, 1]).evaluate([0, 0, 0])


def test_struct_layouts_match_header():
    assert _capi.LINK_OP_DTYPE.itemsize == 64
    assert _capi.GROUP_DTYPE.itemsize == 32
    assert _capi.PRIM_DTYPE.itemsize == 96
    assert _capi.PRIM_DTYPE.fields["lb"][1] == 64 and _capi.PRIM_DTYPE.fields["ub"][1] == 80


@pytest.mark.parametrize(
    "Spec,n_groups,n_spheres,n_pairs,dof",
    [(PandaSpec, 9, 72, 11, 7), (FetchSpec, 11, 69, 14, 8), (PR2DualarmSpec, 15, 161, 45, 14), (OpenArmV10DualSpec, 23, 179, 119, 14)],
)
def test_robot_tables_have_survey_sizes(Spec, n_groups, n_spheres, n_pairs, dof):
    spec = Spec()
    cst = spec.create_collision_const()
    assert len(cst._group_links) == n_groups
    assert sum(len(s) for s in cst._group_spheres) == n_spheres
    assert len(cst._pairs) == n_pairs
    assert cst.q_dim() == dof
    lb, ub = spec.angle_bounds()
    assert np.isfinite(lb).all() and (ub > lb).all()


def test_link_program_is_topological_and_ancestor_closed():
    for Spec in (PandaSpec, FetchSpec, PR2DualarmSpec, OpenArmV10DualSpec):
        spec = Spec()
        cst = spec.create_collision_const()
        kin = spec.get_kin()
        prog = tables.LinkProgram(kin, cst._control_joint_ids, cst._group_links)
        ops = prog.ops
        assert ops[0]["parent"] == -1
        for k in range(1, len(ops)):
            assert 0 <= ops[k]["parent"] < k
        for lid in cst._group_links:
            assert lid in prog.op_of_link
        movable = ops[ops["kind"] != _capi.OP_FIXED]
        assert sorted(movable["q_index"].tolist()) == list(range(cst.q_dim()))
        # quaternions of the URDF layer are normalised
        np.testing.assert_allclose(np.linalg.norm(ops["origin_quat"], axis=1), 1.0, atol=1e-6)


def test_panda_program_is_a_chain():
    spec = PandaSpec()
    cst = spec.create_collision_const()
    prog = tables.LinkProgram(spec.get_kin(), cst._control_joint_ids, cst._group_links)
    assert len(prog.ops) == 9
    assert prog.ops["parent"].tolist() == list(range(-1, 8))
    assert prog.ops["kind"].tolist() == [0] + [1] * 7 + [0]
    # joint1 and joint8 origins have no rotation -> identity flag; the others do not
    assert prog.ops["flags"].tolist() == [1, 1, 0, 0, 0, 0, 0, 0, 1]


def test_baked_joint_uses_stored_angle():
    spec = FetchSpec()
    kin = spec.get_kin()
    lid = kin.get_link_ids(["head_pan_link"])[0]
    q0, t0, ident0 = kin.local_transform(lid)
    assert ident0  # never set: URDF origin, flagged identity
    kin.set_joint_positions(kin.get_joint_ids(["head_pan_joint"]), [0.4])
    q1, t1, ident1 = kin.local_transform(lid)
    assert not ident1
    np.testing.assert_allclose(q1, [0, 0, np.sin(0.2), np.cos(0.2)], atol=1e-15)
    np.testing.assert_allclose(t1, t0)


def test_unknown_names_raise_like_the_reference():
    kin = PandaSpec().get_kin()
    with pytest.raises(ValueError):
        kin.get_link_ids(["nope"])  # std::invalid_argument -> ValueError
    with pytest.raises(ValueError):
        kin.get_joint_ids(["nope"])
    with pytest.raises(RuntimeError):
        KinematicModel("")
    spec = PandaSpec()
    specs = spec.parse_sphere_specs()
    with pytest.raises(RuntimeError):  # primitive_sphere_collision.cpp:171-174
        SphereCollisionCst(spec.get_kin(), spec.control_joint_names, BaseType.FIXED, specs, [("panda_link0", "nope")], None)
    with pytest.raises(RuntimeError):
        psdf.GroundSDF(0.0).translate([0, 0, 1])
    with pytest.raises(RuntimeError):
        psdf.CloudSDF(np.zeros((3, 3)), 0.1).clone()
    with pytest.raises(RuntimeError):
        kin.add_new_link("panda_link3", "panda_link0", [0, 0, 0], [0, 0, 0], True)


def test_psdf_host_objects_match_oracle_state():
    rng = np.random.default_rng(1)
    for i in range(30):
        rot = np.eye(3) if i == 0 else (rpy_matrix(rng.uniform(-3, 3), 0, 0) if i == 1 else rpy_matrix(*rng.uniform(-3, 3, 3)))
        pose = psdf.Pose(rng.standard_normal(3), rot)
        for host in (psdf.BoxSDF([0.5, 0.3, 0.6], pose), psdf.CylinderSDF(0.7, 0.2, pose), psdf.SphereSDF(0.5, pose)):
            o = bridge.OracleSDF.from_psdf(host)
            for step in range(2):
                lb, ub = o.aabb()
                np.testing.assert_allclose(host.lb, lb, atol=1e-14)
                np.testing.assert_allclose(host.ub, ub, atol=1e-14)
                pos, r, aa, za = o.pose()
                np.testing.assert_allclose(host.pose.position, pos, atol=1e-14)
                np.testing.assert_allclose(host.pose.rotation, r, atol=1e-14)
                assert (host.pose.axis_aligned, host.pose.z_axis_aligned) == (aa, za)
                t, a = rng.standard_normal(3), rng.uniform(-1, 1)
                host.translate(t)
                host.rotate_z(a)
                o.translate(t)
                o.rotate_z(a)


def test_clone_and_union_semantics():
    # tests/test_psdf.py:60-86
    box = psdf.BoxSDF([1, 1, 1], psdf.Pose())
    box2 = box.clone()
    box2.translate([1.0, 2.0, 3.0])
    np.testing.assert_allclose(box.pose.position, np.zeros(3))
    np.testing.assert_allclose(box2.pose.position, [1.0, 2.0, 3.0])
    u = psdf.UnionSDF([box])
    u2 = u.clone()
    u2.translate([1.0, 2.0, 3.0])
    np.testing.assert_allclose(box.pose.position, np.zeros(3))
    u.translate([1.0, 2.0, 3.0])
    np.testing.assert_allclose(box.pose.position, [1.0, 2.0, 3.0])
    import copy

    p1 = psdf.Pose(np.zeros(3), np.eye(3))
    for method in (copy.copy, copy.deepcopy):
        p2 = method(p1)
        p2.translate([1.0, 2.0, 3.0])
        np.testing.assert_allclose(p1.position, np.zeros(3))


def test_scene_flatten_order_and_flags():
    spec = FetchSpec()
    cst = spec.create_collision_const(use_cache=False)
    cst.set_sdf(workloads.fetch_table_sdf())
    prims, cloud = tables.flatten_primitives(cst._leaves())
    assert cloud is None
    # 4 anchors first (cyl, box, cyl, cyl), then table box, then ground
    assert prims["type"].tolist() == [2, 1, 2, 2, 1, 4]
    assert (prims["flags"] == 3).all()
    g = prims[-1]
    assert g["lb"][2] == 0 and g["ub"][2] == 0 and g["param0"] == 0
    b = psdf.BoxSDF([1, 1, 1], psdf.Pose([0, 0, 0], rpy_matrix(0.3, 0, 0)))
    p, _ = tables.flatten_primitives([b])
    assert p[0]["flags"] == _capi.PRIMFLAG_Z_ALIGNED
    with pytest.raises(ValueError):
        tables.flatten_primitives([psdf.BoxSDF([1, 1, 1], psdf.Pose([0, 0, 0], np.diag([1.0, 2.0, 1.0])))])


def test_sphere_reorder_matches_oracle_unordered_set_order():
    for maker in (workloads.make_panda, workloads.make_fetch, workloads.make_openarm):
        spec, cst = maker()
        order = bridge.OracleCst(cst).sphere_order()
        k = 0
        for got, orig in zip(cst._group_spheres, cst._group_spheres_unordered):
            np.testing.assert_array_equal(got, orig[order[k : k + len(orig)]])
            k += len(orig)


def test_group_bounding_sphere_matches_oracle():
    spec, cst = workloads.make_panda()
    o = bridge.OracleCst(cst)
    kin = spec.get_kin()
    gs = o.group_spheres(np.zeros(7), accurate=True)
    radii = [tables.group_bounding_sphere(s)[1] for s in cst._group_spheres]
    np.testing.assert_allclose(gs[:, 3], radii, atol=1e-15)
    # panda group radii
    np.testing.assert_allclose(radii, [0.1195, 0.171, 0.1488, 0.189, 0.1305, 0.174, 0.1142, 0.065, 0.1156], atol=5e-4)


def test_step_ratio_host_matches_oracle():
    rng = np.random.default_rng(2)
    w = rng.uniform(0.02, 0.2, 7)
    vb = ValidatorConfig(ValidatorType.BOX, 0.0, list(w))
    ve = ValidatorConfig(ValidatorType.EUCLIDEAN, 0.07, [])
    ob, oe = bridge.OracleValidator(0, 7, w), bridge.OracleValidator(1, 7, None, 0.07)
    for _ in range(200):
        a, b = rng.standard_normal(7), rng.standard_normal(7)
        assert step_ratio(vb, a, b) == ob.step_ratio(a, b)
        assert abs(step_ratio(ve, a, b) - oe.step_ratio(a, b)) < 1e-15
    assert step_ratio(vb, a, a) == float("inf") and step_ratio(ve, a, a) == float("inf")


def test_unpack_bits_roundtrip():
    rng = np.random.default_rng(0)
    for n in (0, 1, 31, 32, 33, 1000):
        v = rng.random(n) < 0.5
        words = np.zeros((n + 31) // 32, dtype=np.uint32)
        for i in np.where(v)[0]:
            words[i >> 5] |= np.uint32(1) << np.uint32(i & 31)
        assert (tables.unpack_bits(words.view(np.int32), n) == v).all()


def test_yaml_conf_loader_matches_imported_tables(tmp_path):
    """RobotSpec(conf_file=..., urdf_file=...) reads the reference's YAML schema at run time
    (include, group pairs, sphere rows).  With /root/reference mounted the reference's own YAMLs are
    loaded and must reproduce the imported JSON tables; a synthetic conf covers the schema anywhere."""
    import json

    from plainmp_b200 import robot_spec as rs

    ref_conf = Path("/root/reference/src/plainmp/conf")
    if ref_conf.is_dir():
        for name, rel in [("panda", "panda.yaml"), ("fetch", "fetch.yaml"), ("pr2_dualarm", "pr2_dualarm.yaml"),
                          ("openarm_v10_dual", "openarm_v10/openarm_v10_dual.yaml")]:
            got = rs.load_conf_yaml(ref_conf / rel)
            want = json.loads((rs.DATA_DIR / "conf" / f"{name}.json").read_text())
            for k in want:
                assert got[k] == want[k], (name, k)
    (tmp_path / "common.yaml").write_text(
        "collision_spheres:\n  panda_link1:\n    spheres:\n      - [0.0, 0.0, -0.1, 0.06]\n      - [0.0, -0.05, 0.0, 0.06]\n"
        "  panda_link5:\n    spheres:\n      - [0.0, 0.0, -0.2, 0.05]\n"
    )
    (tmp_path / "robot.yaml").write_text(
        "include: [common.yaml]\ncontrol_joint_names: [panda_joint1, panda_joint2, panda_joint3, panda_joint4, panda_joint5, panda_joint6, panda_joint7]\n"
        "self_collision_group_pairs:\n  - group1: [panda_link1]\n    group2: [panda_link5]\n"
        "only_self_collision_links: [panda_link1]\n"
        "custom_links:\n  tool:\n    parent_link: panda_hand\n    position: [0, 0, 0.1]\n    rpy: [0, 0, 0]\n"
    )
    spec = rs.RobotSpec(conf_file=tmp_path / "robot.yaml", urdf_file=rs.DATA_DIR / "urdf" / "panda.urdf")
    assert spec.get_self_collision_pairs() == [["panda_link1", "panda_link5"]]
    cst = spec.create_collision_const(use_cache=False)
    assert cst._group_link_names == ["panda_link1", "panda_link5"] and cst._group_only_self == [True, False]
    assert spec.get_kin().get_link_ids(["tool"])[0] >= 0
    (tmp_path / "dup.yaml").write_text(
        "control_joint_names: []\ncollision_spheres: {}\nself_collision_pairs: [[a, b], [b, a]]\n"
    )
    with pytest.raises(AssertionError):
        rs.load_conf_yaml(tmp_path / "dup.yaml")


def test_bench_reference_arm_prints_one_contract_line():
    """bench.py --impl reference runs on CPU only (the reference's compiled C++ if oracle/_ref is
    there, else the restatement) and prints exactly one JSON line with the driver's contract keys."""
    import json
    import subprocess
    import sys

    out = subprocess.run(
        [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-sample", "20000"],
        capture_output=True, text=True, check=True, timeout=600,
    ).stdout.strip().splitlines()
    assert len(out) == 1
    d = json.loads(out[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "collision_checked_configs_per_sec" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_queue_overflow_protocol_lane_level_model():
    """A lane-level model of phase 1's env-queue protocol (csrc/pmp_v3.cuh: v3_push_env +
    warp_check_interp): 32 lanes, a 128-entry queue, drain on overflow.  Restarting an overflowing group
    from its full mask after every drain (the original handling) livelocks when one group alone yields
    more than 128 items and no lane dies; resuming with the remaining bits from the second
    consecutive stall on terminates and drains every (lane, group, primitive) item."""
    import random

    KQE = 128

    def simulate(masks, resume_enabled, die_prob=0.0, max_iters=2000, seed=0):
        rng = random.Random(seed)
        bad, drained, queue = [False] * 32, set(), []
        pm, gi, resume, stall_gi, iters = [0] * 32, 0, False, -1, 0

        def push(g):
            for lane in range(32):
                if bad[lane]:
                    pm[lane] = 0
            while any(pm):
                cnt = sum(1 for lane in range(32) if pm[lane])
                if len(queue) + cnt > KQE:
                    return False
                for lane in range(32):
                    if pm[lane]:
                        b = (pm[lane] & -pm[lane]).bit_length() - 1
                        pm[lane] &= pm[lane] - 1
                        queue.append((lane, g, b))
            return True

        while True:
            iters += 1
            if iters > max_iters:
                return False, drained
            stalled = False
            while gi < len(masks):
                if not (resume_enabled and resume):
                    pm[:] = masks[gi]
                if not push(gi):
                    resume, stall_gi, stalled = (stall_gi == gi), gi, True
                    break
                resume = False
                gi += 1
            for it in queue:
                if not bad[it[0]]:
                    drained.add(it)
                    if rng.random() < die_prob:
                        bad[it[0]] = True
            queue.clear()
            if not stalled:
                return True, drained

    worst = [[0x7F] * 32]  # 7 primitives pass on every lane: 224 items from one group
    assert simulate(worst, resume_enabled=False)[0] is False  # the livelock
    ok, drained = simulate(worst, resume_enabled=True)
    assert ok and len(drained) == 224
    rng = random.Random(1)
    for trial in range(300):
        masks = [[rng.getrandbits(rng.choice([3, 7, 12, 31])) if rng.random() < rng.choice([0.2, 0.9, 1.0]) else 0 for _ in range(32)]
                 for _ in range(rng.randint(1, 12))]
        ok, drained = simulate(masks, True)
        want = {(lane, g, b) for g, row in enumerate(masks) for lane in range(32) for b in range(32) if (row[lane] >> b) & 1}
        assert ok and drained == want
        assert simulate(masks, True, die_prob=0.02, seed=trial)[0]


def test_ompl_adaptor_header_compiles_against_the_ompl_stand_in(tmp_path):
    """include/plainmp_b200_ompl.hpp + the test probe compile with g++ against oracle/ref_shim's
    ompl::base stand-in and link against the C-ABI library (no GPU needed to compile / load)."""
    import subprocess

    so = tmp_path / "libprobe.so"
    subprocess.run(
        ["g++", "-O1", "-std=c++17", "-fPIC", "-shared", f"-I{ROOT / 'include'}", f"-I{ROOT / 'oracle' / 'ref_shim'}",
         "-o", str(so), str(ROOT / "tests" / "cpp" / "ompl_adaptor_probe.cpp"),
         f"-L{ROOT / 'plainmp_b200'}", "-lplainmp_b200", f"-Wl,-rpath,{ROOT / 'plainmp_b200'}"],
        check=True,
    )
    _capi.load()
    assert C.CDLL(str(so)).probe_validators is not None


def test_group_run_unit_index_formula_is_exact():
    """pmp_v3.cuh, v3_group_run: unit u of a group run belongs to the (u / count)-th surviving lane; the kernel
    computes the quotient as (int)((float(u) + 0.5f) * (1.f / count)) for u < 32 * count.  FP32 arithmetic restated
    here: exact for every group size the shared-memory tables can hold (<= 12 800 spheres)."""
    for count in list(range(1, 1200)) + [2047, 2048, 4095, 4096, 5000, 8191, 12800]:
        u = np.arange(32 * count, dtype=np.int64)
        inv = np.float32(1.0) / np.float32(count)
        j = ((u.astype(np.float32) + np.float32(0.5)) * inv).astype(np.int32)
        assert (j == u // count).all(), count
