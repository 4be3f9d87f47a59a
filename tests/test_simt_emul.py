"""The kernel sources of plainmp_b200/csrc/ compiled UNCHANGED for a SIMT emulator (tests/cpu_emul/:
one OS thread per CUDA thread, warp collectives as rendezvous) and run on the CPU against the FP64
oracle.  Test infrastructure only: the product never loads this library and still refuses to compute
without a CUDA device (test_no_cpu_fallback_without_device).  What it buys: the kernels' control flow
— warp-local queues and their overflow handling, sub-warp drains, dynamic edge refill, the session's
flat waypoint kernel — is exercised in the CPU tier, a few hundred items at a time; in particular the
queue-overflow livelock found by fuzzing on the GPU is reproduced and its fix verified here.
FP32 results differ from the GPU's in the last bits (no FMA contraction on the host), which the
1e-5 contact band absorbs."""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np
import pytest

from oracle import bridge
from plainmp_b200 import _capi, psdf, tables, workloads
from plainmp_b200.constraint import SphereAttachmentSpec, SphereCollisionCst
from plainmp_b200.kinematics import BaseType
from plainmp_b200.robot_spec import PandaSpec, rpy_matrix

sys.path.insert(0, str(Path(__file__).resolve().parent))
from cpu_emul import build as emul_build  # noqa: E402
from fuzz_util import random_case, sample_q, stress_scene  # noqa: E402
from parity_util import assert_jacobian_parity  # noqa: E402


@pytest.fixture(scope="module")
def emu():
    lib = C.CDLL(os.environ.get("PMP_EMUL_LIB") or str(emul_build.build()))  # PMP_EMUL_LIB: tools/coverage_emul.sh
    for name, (res, args) in _capi.SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


class EmuTables:
    """pmp_robot_t / pmp_scene_t of the EMULATED library for a host-mirror constraint."""

    def __init__(self, lib, cst):
        self.lib = lib
        masks = None
        if hasattr(cst, "_merge_parts"):  # IneqCompositeCst: the members merged into one table with per-group scene masks
            m, first = cst._merge_parts(), cst.constraints[0]
            a = tables.RobotTable.build_arrays(first.get_kin(), first._control_joint_ids, first._base_type, m["links"], m["spheres"], m["only_self"], m["pairs"])
            masks = np.ascontiguousarray(m["masks"], dtype=np.uint32)
            leaves = m["leaves"]
        else:
            a = tables.RobotTable.build_arrays(
                cst.get_kin(), cst._control_joint_ids, cst._base_type, cst._group_links, cst._group_spheres, cst._group_only_self, cst._pairs
            )
            leaves = cst._leaves()
        self._keep, self._masks = a, masks
        d = _capi.RobotDesc()
        d.dof, d.base_type = a["dof"], int(a["base_type"])
        for i in range(7):
            d.base_pose[i] = float(a["base_pose"][i])
        ops = a["program"].ops
        d.n_ops, d.ops = len(ops), ops.ctypes.data
        d.n_groups, d.groups = len(a["groups"]), a["groups"].ctypes.data
        d.n_spheres, d.spheres = len(a["spheres"]), a["spheres"].ctypes.data
        d.n_pairs, d.pairs = len(a["pairs"]), a["pairs"].ctypes.data
        self.q_dim = a["q_dim"]
        self._desc = d
        self.robot = C.c_void_p()
        if masks is not None:
            assert lib.pmp_robot_create_ex(C.byref(d), masks.ctypes.data, C.byref(self.robot)) == 0, lib.pmp_last_error()
        else:
            assert lib.pmp_robot_create(C.byref(d), C.byref(self.robot)) == 0, lib.pmp_last_error()
        self.scene = None
        if leaves:
            prims, cloud = tables.flatten_primitives(leaves)
            self._prims = prims
            sd = _capi.SceneDesc()
            sd.n_prims, sd.prims = len(prims), (prims.ctypes.data if len(prims) else None)
            if cloud is not None:
                self._cloud = np.ascontiguousarray(cloud.points, dtype=np.float32)
                sd.n_cloud_points, sd.cloud_points = len(self._cloud), self._cloud.ctypes.data
                sd.cloud_radius, sd.cloud_cell = cloud.radius, 0.02  # coarse field: fewer cells to fill in emulation
                self._cloud_radii = getattr(cloud, "point_radii", None)
                if self._cloud_radii is not None:
                    sd.cloud_point_radii = self._cloud_radii.ctypes.data
            self.scene = C.c_void_p()
            assert lib.pmp_scene_create(C.byref(sd), C.byref(self.scene)) == 0, lib.pmp_last_error()

    def check(self, Q):
        Q = np.ascontiguousarray(Q, dtype=np.float32)
        n = Q.shape[0]
        bits = np.zeros((n + 31) // 32, dtype=np.uint32)
        cnt = C.c_uint64(0)
        rc = self.lib.pmp_check_configs_host(self.robot, self.scene, Q.ctypes.data, n, bits.ctypes.data, C.byref(cnt))
        assert rc == 0, self.lib.pmp_last_error()
        v = tables.unpack_bits(bits, n)
        assert int(cnt.value) == int(v.sum())
        return v

    def evaluate(self, Q):
        Q = np.ascontiguousarray(Q, dtype=np.float32)
        out = np.zeros((Q.shape[0], 2), dtype=np.float32)
        assert self.lib.pmp_eval_configs_host(self.robot, self.scene, Q.ctypes.data, Q.shape[0], out.ctypes.data) == 0
        return out.astype(np.float64)

    # device-pointer entry points: in the emulation "device" memory is host memory and launches are
    # synchronous, so numpy buffers go straight through
    def evaluate_grad(self, Q):
        Q = np.ascontiguousarray(Q, dtype=np.float32)
        out = np.zeros((Q.shape[0], 2), dtype=np.float32)
        jac = np.zeros((Q.shape[0], 2, self.q_dim), dtype=np.float32)
        rc = self.lib.pmp_eval_grad_configs(self.robot, self.scene, Q.ctypes.data, Q.shape[0], out.ctypes.data, jac.ctypes.data if jac.size else None, None)
        assert rc == 0, self.lib.pmp_last_error()
        return out.astype(np.float64), jac.astype(np.float64)

    def fk_spheres(self, Q, accurate):
        Q = np.ascontiguousarray(Q, dtype=np.float32)
        out = np.zeros((Q.shape[0], len(self._keep["spheres"]), 4), dtype=np.float32)
        rc = self.lib.pmp_fk_spheres(self.robot, Q.ctypes.data, Q.shape[0], int(accurate), out.ctypes.data, None)
        assert rc == 0, self.lib.pmp_last_error()
        return out

    def check_edges(self, vtype, widths, res, q0, q1):
        q0, q1 = np.ascontiguousarray(q0, dtype=np.float32), np.ascontiguousarray(q1, dtype=np.float32)
        v, keep = tables.make_validator(vtype, q0.shape[1], widths if vtype == 0 else None, res)
        n = q0.shape[0]
        bits = np.zeros((n + 31) // 32, dtype=np.uint32)
        nv, nc = C.c_uint64(0), C.c_uint64(0)
        rc = self.lib.pmp_check_edges_host(self.robot, self.scene, C.byref(v), q0.ctypes.data, q1.ctypes.data, n, bits.ctypes.data, C.byref(nv), C.byref(nc))
        assert rc == 0, self.lib.pmp_last_error()
        return tables.unpack_bits(bits, n), int(nc.value)

    def __del__(self):
        if getattr(self, "scene", None):
            self.lib.pmp_scene_destroy(self.scene)
        if getattr(self, "robot", None):
            self.lib.pmp_robot_destroy(self.robot)


def _parity(got, want, margin, what):
    bad = [int(i) for i in np.where(got != want)[0] if margin[i] > 1e-5]
    assert not bad, f"{what}: validity mismatches outside the band at {bad[:5]}"


def _edge_parity(out, n_checked, o, ov, q0, q1, what):
    """The edge rule of the GPU tier: an edge may differ from the serial walk only if one of the waypoints
    the walk visits lies in the 1e-5 contact band; when no edge does, the is_valid call count is the
    reference's, exactly."""
    want, calls, margin = ov.check_motion_batch_ex(o, q0, q1)
    clear = margin > 1e-5
    bad = np.where((np.asarray(out).astype(bool) != want) & clear)[0]
    assert len(bad) == 0, f"{what}: edges {bad[:5]} differ outside the band (margins {margin[bad[:5]]})"
    if n_checked is not None and clear.all():
        assert int(n_checked) == int(calls.sum()), f"{what}: {n_checked} is_valid calls, the reference makes {int(calls.sum())}"
    return want


@pytest.mark.parametrize("maker", ["panda", "panda_ceiled", "fetch"])
def test_emulated_validity_and_distances_match_oracle(emu, maker):
    spec, cst = {"panda": workloads.make_panda, "panda_ceiled": lambda: workloads.make_panda(ceiled=True), "fetch": workloads.make_fetch}[maker]()
    t = EmuTables(emu, cst)
    Q = workloads.random_configs(spec, 300, 11)  # two full tiles + a ragged one
    o = bridge.OracleCst(cst)
    Qd = Q.astype(np.float64)
    want = o.is_valid_batch(Qd)
    _parity(t.check(Q), want, o.margin_batch(Qd), maker)
    assert 0.1 < want.mean() < 0.9
    np.testing.assert_allclose(t.evaluate(Q[:128]), o.evaluate_batch(Qd[:128]), atol=1e-5, rtol=0)


@pytest.mark.parametrize("mode", ["1", "2"])
def test_group_larger_than_the_queue_terminates_in_the_emulation(emu, monkeypatch, mode):
    """The livelock of DESIGN section 8: a group inside the AABBs of 8 plates (8 x 32 = 256 items > the
    128-entry queue) with every lane alive.  Before the resume fix this never returned."""
    monkeypatch.setenv("PMP_V3_MODE", mode)
    spec = PandaSpec()
    one = SphereAttachmentSpec("panda_link0", np.zeros((3, 4)) + np.array([[0.0], [0.0], [0.3]]), np.full(4, 0.01), False)
    arm = SphereAttachmentSpec("panda_link5", np.array([[0.0, 0.0], [0.0, 0.05], [-0.1, 0.0]]), np.array([0.06, 0.05]), False)
    cst = SphereCollisionCst(spec.get_kin(), spec.control_joint_names, BaseType.FIXED, [one, arm], [("panda_link0", "panda_link5")], None, True)
    plates = []
    for r in (1.0, 1.3):
        for k in range(4):
            phi = 0.5 * np.pi * k
            plate = psdf.BoxSDF([0.05, 4.0, 4.0], psdf.Pose([r * np.cos(phi), r * np.sin(phi), 0.3]))
            plate.rotate_z(phi + 0.25 * np.pi)
            plates.append(plate)
    cst.set_sdf(psdf.UnionSDF(plates))
    t = EmuTables(emu, cst)
    Q = workloads.random_configs(spec, 96, 3)
    o = bridge.OracleCst(cst)
    want = o.is_valid_batch(Q.astype(np.float64))
    _parity(t.check(Q), want, o.margin_batch(Q.astype(np.float64)), "8 plates")
    assert want.mean() > 0.5


import os  # noqa: E402

_s, _c = (int(x) for x in os.environ.get("PMP_FUZZ_SEEDS", "1000:40").split(":"))


@pytest.mark.parametrize("seed", [40289] + list(range(_s, _s + _c)))
def test_emulated_random_robots(emu, seed):
    """The suite's fuzz seeds (among them 1003 / 1012 / 1018, which exposed evaluate on folded prismatic
    joints on the GPU) and 40289 (the queue livelock), replayed on the emulated kernels, clouds included
    (the index build runs in the emulation too, with a host stand-in for the one CUB sort)."""
    cst, lb, ub, desc = random_case(seed)
    t = EmuTables(emu, cst)
    Q = sample_q(lb, ub, 160, seed)
    o = bridge.OracleCst(cst)
    Qd = Q.astype(np.float64)
    _parity(t.check(Q), o.is_valid_batch(Qd), o.margin_batch(Qd), desc)
    np.testing.assert_allclose(t.evaluate(Q[:64]), o.evaluate_batch(Qd[:64]), atol=2e-5, rtol=0, equal_nan=True, err_msg=desc)


def test_emulated_edges_follow_the_reference_order(emu):
    """The lane = edge kernel with dynamic refill: validity and the NUMBER of waypoint checks equal the
    oracle's (which equals the compiled reference's), i.e. the visiting order with its early exit."""
    spec, cst = workloads.make_panda(ceiled=True)
    t = EmuTables(emu, cst)
    o = bridge.OracleCst(cst)
    widths = np.full(7, 0.05)
    q0, q1 = workloads.random_edges(spec, 80, 3, True, lambda q: o.is_valid_batch(q.astype(np.float64)))
    got, n_checks = t.check_edges(0, widths, 0.0, q0, q1)
    _edge_parity(got, n_checks, o, bridge.OracleValidator(0, 7, widths), q0.astype(np.float64), q1.astype(np.float64), "edge kernel")


def test_emulated_session_flat_and_early_exit_paths(emu):
    """pmp_session_*: FP64 states in, bytes out; the waypoint-parallel flat kernel (zero-copy and staged)
    and the early-exit kernel behind one entry point."""
    spec, cst = workloads.make_panda(ceiled=True)
    t = EmuTables(emu, cst)
    o = bridge.OracleCst(cst)
    widths = np.full(7, 0.05)
    v, keep = tables.make_validator(0, 7, widths, 0.0)
    ses = C.c_void_p()
    assert emu.pmp_session_create(t.robot, t.scene, C.byref(v), 64, C.byref(ses)) == 0, emu.pmp_last_error()
    q0, q1 = workloads.random_edges(spec, 40, 5, True, lambda q: o.is_valid_batch(q.astype(np.float64)))
    q0, q1 = np.ascontiguousarray(q0, dtype=np.float64), np.ascontiguousarray(q1, dtype=np.float64)
    out = np.zeros(40, dtype=np.uint8)
    nc = C.c_uint64(0)
    assert emu.pmp_session_check_motion(ses, q0.ctypes.data, q1.ctypes.data, 40, out.ctypes.data, C.byref(nc)) == 0
    _edge_parity(out, nc.value, o, bridge.OracleValidator(0, 7, widths), q0, q1, "session, small batch")
    one = np.zeros(1, dtype=np.uint8)
    for i in range(5):  # single edges: the zero-copy path
        assert emu.pmp_session_check_motion(ses, q0[i].ctypes.data, q1[i].ctypes.data, 1, one.ctypes.data, None) == 0
        assert bool(one[0]) == bool(out[i])
    st = np.zeros(40, dtype=np.uint8)
    assert emu.pmp_session_is_valid(ses, q1.ctypes.data, 40, st.ctypes.data) == 0
    Qd = q1
    bad = np.where(st.astype(bool) != o.is_valid_batch(Qd))[0]
    assert all(o.margin_batch(Qd)[i] < 1e-5 for i in bad)
    emu.pmp_session_destroy(ses)


@pytest.mark.parametrize("base_type", [BaseType.PLANAR, BaseType.FLOATING])
def test_emulated_moving_base(emu, base_type):
    from plainmp_b200.robot_spec import PR2RarmSpec

    spec = PR2RarmSpec(base_type=base_type)
    cst = spec.create_collision_const(self_collision=True, use_cache=False)
    cst.set_sdf(psdf.BoxSDF([1, 1, 1], psdf.Pose([0.5, 0.5, 0.5])))
    t = EmuTables(emu, cst)
    o = bridge.OracleCst(cst)
    rng = np.random.default_rng(2)
    Q = (rng.standard_normal((128, o.q_dim)) * 0.4).astype(np.float32)
    Qd = Q.astype(np.float64)
    _parity(t.check(Q), o.is_valid_batch(Qd), o.margin_batch(Qd), str(base_type))


@pytest.mark.parametrize("seed", list(range(3000, 3000 + max(6, _c // 8))))
def test_emulated_big_random_robots(emu, seed):
    """20-45 links, up to 35 groups, 28-44 primitives (the > 30-primitive overflow item, frequent queue
    overflows, 64-thread CTAs) on the emulated kernels."""
    cst, lb, ub, desc = random_case(seed, big=True)
    t = EmuTables(emu, cst)
    Q = sample_q(lb, ub, 96, seed)
    o = bridge.OracleCst(cst)
    Qd = Q.astype(np.float64)
    got, want, m = t.check(Q), o.is_valid_batch(Qd), o.margin_batch(Qd)
    bad = [int(i) for i in np.where(got != want)[0] if m[i] > 2e-5]
    assert not bad, desc


def test_emulated_cloud_scene_and_nearest_queries(emu):
    """PR2 dual arm against a 3 000-point table cloud: the cooperative 8-lane tree traversal of the
    narrow phase, the distance field, and the 8-lane nearest-point kernel."""
    spec, cst = workloads.make_pr2_cloud(3000)
    t = EmuTables(emu, cst)
    o = bridge.OracleCst(cst)
    Q = workloads.random_configs(spec, 96, 4)
    Qd = Q.astype(np.float64)
    want = o.is_valid_batch(Qd)
    _parity(t.check(Q), want, o.margin_batch(Qd), "pr2 cloud")
    assert 0.05 < want.mean() < 0.95
    np.testing.assert_allclose(t.evaluate(Q[:32]), o.evaluate_batch(Qd[:32]), atol=1e-5, rtol=0)
    pts = cst.get_sdf().points
    rng = np.random.default_rng(1)
    P = (pts.mean(axis=0) + rng.uniform(-1.2, 1.2, (200, 3))).astype(np.float32)
    sq = np.zeros(200, dtype=np.float32)
    idx = np.zeros(200, dtype=np.int32)
    assert emu.pmp_cloud_nearest(t.scene, P.ctypes.data, 200, sq.ctypes.data, idx.ctypes.data, None) == 0
    brute = ((P[:, None, :].astype(np.float64) - pts[None, :, :]) ** 2).sum(-1)
    np.testing.assert_allclose(sq, brute.min(1), atol=2e-5, rtol=1e-5)
    np.testing.assert_allclose(brute[np.arange(200), idx], brute.min(1), atol=2e-5, rtol=1e-5)


def test_emulated_far_group_centres_fp16_layout(emu, monkeypatch):
    """MODE 2 (FP16 group centres) with centres farther than the 4 m the inflation constants cover: the
    `far` flag must switch the self broad-phase culls off (same case as the GPU test)."""
    from plainmp_b200.robot_spec import PR2RarmSpec

    monkeypatch.setenv("PMP_V3_MODE", "2")
    spec = PR2RarmSpec()
    kin = spec.get_kin()
    name = "far_payload_link_emul"
    try:
        kin.get_link_ids([name])
    except ValueError:
        kin.add_new_link(name, "r_gripper_palm_link", [5.5, 0.0, 0.0], [0.0, 0.0, 0.0], True)
    rng = np.random.default_rng(71)
    specs = [
        SphereAttachmentSpec("r_gripper_palm_link", rng.uniform(-0.05, 0.05, (3, 6)), rng.uniform(0.03, 0.06, 6), False),
        SphereAttachmentSpec(name, rng.uniform(-0.3, 0.3, (3, 9)), rng.uniform(0.1, 0.3, 9), False),
        SphereAttachmentSpec("base_link", rng.uniform(-0.3, 0.3, (3, 5)) + np.array([[4.5], [0.0], [0.8]]), rng.uniform(1.0, 2.0, 5), True),
        SphereAttachmentSpec("r_forearm_link", rng.uniform(-0.1, 0.1, (3, 4)), rng.uniform(0.03, 0.06, 4), False),
    ]
    pairs = [(name, "base_link"), ("r_gripper_palm_link", "base_link"), ("r_forearm_link", name)]
    cst = SphereCollisionCst(kin, spec.control_joint_names, BaseType.FIXED, specs, pairs, None, True)
    t = EmuTables(emu, cst)
    Q = workloads.random_configs(spec, 512, 72)
    o = bridge.OracleCst(cst)
    Qd = Q.astype(np.float64)
    want = o.is_valid_batch(Qd)
    assert (~want).sum() >= 5  # some self collisions, all found through the far path
    _parity(t.check(Q), want, o.margin_batch(Qd), "far centres")


@pytest.mark.parametrize("seed", list(range(7000, 7000 + max(8, _c // 5))))
def test_emulated_edges_on_random_robots(emu, seed):
    """Both validators on random robots through the emulated lane = edge kernel (dynamic refill, the
    reference's visiting order, natural order for n_test >= 129) and through the session's flat kernel."""
    cst, lb, ub, desc = random_case(seed)
    dim = len(lb)
    if dim == 0:
        pytest.skip("no degrees of freedom")
    t = EmuTables(emu, cst)
    o = bridge.OracleCst(cst)
    rng = np.random.default_rng(seed)
    widths = rng.uniform(0.01, 0.3, dim)
    res = float(rng.uniform(0.01, 0.3))
    q0 = sample_q(lb, ub, 48, seed + 1)
    q1 = np.clip(q0.astype(np.float64) + rng.uniform(-0.8, 0.8, q0.shape) * rng.choice([0.05, 1.0, 3.0]), lb, ub).astype(np.float32)
    for vtype in (0, 1):
        got, nck = t.check_edges(vtype, widths, res, q0, q1)
        want = _edge_parity(got, nck, o, bridge.OracleValidator(vtype, dim, widths, res), q0.astype(np.float64), q1.astype(np.float64), desc)
        v, keep = tables.make_validator(vtype, dim, widths if vtype == 0 else None, res)
        ses = C.c_void_p()
        assert emu.pmp_session_create(t.robot, t.scene, C.byref(v), 64, C.byref(ses)) == 0
        a, b = np.ascontiguousarray(q0, dtype=np.float64), np.ascontiguousarray(q1, dtype=np.float64)
        out = np.zeros(48, dtype=np.uint8)
        nck = C.c_uint64(0)
        assert emu.pmp_session_check_motion(ses, a.ctypes.data, b.ctypes.data, 48, out.ctypes.data, C.byref(nck)) == 0
        emu.pmp_session_destroy(ses)
        _edge_parity(out, nck.value, o, bridge.OracleValidator(vtype, dim, widths, res), a, b, desc + " (session)")


GOLDEN = Path(__file__).parent / "golden"


@pytest.mark.parametrize("name", ["panda_dualbars_seed1", "fetch_table_seed2", "pr2_cloud20k_seed4", "openarm_dual_seed6"])
def test_emulated_kernels_against_reference_golden_vectors(emu, name):
    """The kernel source (emulated) against the committed REFERENCE output: validity bits outside the
    contact band and evaluate() distances for the head of every golden file."""
    g = np.load(GOLDEN / f"{name}.npz")
    maker = {"panda_dualbars_seed1": workloads.make_panda, "fetch_table_seed2": workloads.make_fetch,
             "pr2_cloud20k_seed4": lambda: workloads.make_pr2_cloud(20000), "openarm_dual_seed6": workloads.make_openarm}[name]
    spec, cst = maker()
    n = 512 if "cloud" not in name else 192
    Q = workloads.random_configs(spec, int(g["n"]), int(g["seed"]))[:n]
    t = EmuTables(emu, cst)
    want = np.unpackbits(g["valid_bits"], bitorder="little")[:n].astype(bool)
    _parity(t.check(Q), want, g["margin"][:n], name)
    m = min(n, 128)
    np.testing.assert_allclose(t.evaluate(Q[:m]), g["dist"][:m], atol=1e-5, rtol=0)


@pytest.mark.parametrize("mode", ["", "2"])
@pytest.mark.parametrize("seed", list(range(9000, 9000 + max(10, _c // 4))))
def test_emulated_queue_stress(emu, monkeypatch, seed, mode):
    """Queue-overflow stress on random robots (all base types, with and without a cloud leaf): scenes
    built so that single groups exceed the warp queue; the check must return and agree with the oracle."""
    if mode:
        monkeypatch.setenv("PMP_V3_MODE", mode)
    cst, lb, ub, desc = random_case(seed)
    rng = np.random.default_rng(seed)
    cst.set_sdf(stress_scene(rng, bool(seed % 3 == 0)))
    t = EmuTables(emu, cst)
    Q = sample_q(lb, ub, 128, seed)
    o = bridge.OracleCst(cst)
    Qd = Q.astype(np.float64)
    _parity(t.check(Q), o.is_valid_batch(Qd), o.margin_batch(Qd), desc)


@pytest.mark.parametrize("seed", [1003, 1012, 1018] + list(range(5000, 5000 + max(10, _c // 4))))
def test_emulated_jacobian_rows_and_sphere_placement(emu, seed):
    """evaluate() with Jacobian rows (arg-min tracking + point-Jacobian pass) and get_all_spheres in
    both FK precisions on the emulated kernels, random robots of every base type."""
    cst, lb, ub, desc = random_case(seed)
    t = EmuTables(emu, cst)
    Q = sample_q(lb, ub, 96, seed)
    Qd = Q.astype(np.float64)
    o = bridge.OracleCst(cst)
    vals, jac = t.evaluate_grad(Q)
    want_v, want_j = o.evaluate_grad_batch(Qd)
    np.testing.assert_allclose(vals, want_v, atol=2e-5, rtol=0, equal_nan=True, err_msg=desc)
    if jac.size:
        assert_jacobian_parity(o, Q, jac, want_j, desc, max_fraction=0.05)
    for accurate in (False, True):
        np.testing.assert_allclose(t.fk_spheres(Q[:32], accurate), o.spheres_batch(Qd[:32], accurate), atol=2e-5, rtol=0, err_msg=desc)


def test_emulated_scene_updates_and_sdf_queries(emu):
    """The scene update path (in-place primitive records; cloud re-index through the grow-only ping-pong
    buffers, growing, shrinking and removal), a session following a scene swap, and the batched psdf
    queries, all on the emulated kernels."""
    spec, cst = workloads.make_panda()
    rng = np.random.default_rng(5)

    def cloud(n, shift):
        return rng.uniform(-0.15, 0.15, (n, 3)) + np.array([0.45 + shift, 0.0, 0.55])

    box = psdf.BoxSDF([0.2, 0.3, 0.25], psdf.Pose([0.5, -0.4, 0.3]))
    cyl = psdf.CylinderSDF(0.12, 0.5, psdf.Pose([-0.3, 0.45, 0.25]))
    pts = cloud(700, 0.0)
    cst.set_sdf(psdf.UnionSDF([box, cyl, psdf.CloudSDF(pts, 0.01)]))
    t = EmuTables(emu, cst)
    Q = workloads.random_configs(spec, 96, 9)
    Qd = Q.astype(np.float64)

    def agree(what):
        o = bridge.OracleCst(cst)
        want = o.is_valid_batch(Qd)
        _parity(t.check(Q), want, o.margin_batch(Qd), what)
        np.testing.assert_allclose(t.evaluate(Q[:32]), o.evaluate_batch(Qd[:32]), atol=1e-5, rtol=0, err_msg=what)
        return want

    first = agree("initial")
    # SDFBase::translate / rotate_z on leaves -> new records, same count and order
    box.translate(np.array([-0.1, 0.25, 0.1]))
    cyl.rotate_z(0.7)
    cyl.translate(np.array([0.35, -0.2, 0.0]))
    prims, _ = tables.flatten_primitives(cst._leaves())
    assert emu.pmp_scene_update_prims(t.scene, prims.ctypes.data, len(prims), None) == 0, emu.pmp_last_error()
    moved = agree("moved primitives")
    assert (first != moved).any()
    assert emu.pmp_scene_update_prims(t.scene, prims.ctypes.data, len(prims) + 1, None) != 0  # count is fixed
    # new depth frames: same size, larger (buffers grow), smaller, none
    for n, shift in ((700, 0.05), (2500, -0.05), (90, 0.1), (0, 0.0)):
        leaves = [box, cyl] + ([psdf.CloudSDF(cloud(n, shift), 0.01)] if n else [])
        cst.set_sdf(psdf.UnionSDF(leaves))
        p32 = np.ascontiguousarray(leaves[-1].points, dtype=np.float32) if n else None
        rc = emu.pmp_scene_update_cloud(t.scene, p32.ctypes.data if n else None, n, 0.01, 0.02)
        assert rc == 0, emu.pmp_last_error()
        agree(f"cloud of {n} points")
    # a session keeps working across a scene swap
    cst.set_sdf(psdf.UnionSDF([box, cyl, psdf.CloudSDF(cloud(400, 0.0), 0.01)]))
    t2 = EmuTables(emu, cst)
    v, keep = tables.make_validator(0, 7, np.full(7, 0.05), 0.0)
    ses = C.c_void_p()
    assert emu.pmp_session_create(t.robot, t.scene, C.byref(v), 64, C.byref(ses)) == 0
    st = np.zeros(len(Qd), dtype=np.uint8)
    assert emu.pmp_session_is_valid(ses, Qd.ctypes.data, len(Qd), st.ctypes.data) == 0
    before = st.astype(bool).copy()
    assert emu.pmp_session_set_scene(ses, t2.scene) == 0
    assert emu.pmp_session_is_valid(ses, Qd.ctypes.data, len(Qd), st.ctypes.data) == 0
    o = bridge.OracleCst(cst)
    _parity(st.astype(bool), o.is_valid_batch(Qd), o.margin_batch(Qd), "session after set_scene")
    assert (before != st.astype(bool)).any()
    emu.pmp_session_destroy(ses)
    # UnionSDF::evaluate_batch / is_outside over primitives + cloud
    sdf = bridge.OracleSDF.from_psdf(cst.get_sdf())
    P = rng.uniform(-0.8, 0.9, (300, 3))
    P32 = np.ascontiguousarray(P, dtype=np.float32)
    out = np.zeros(300, dtype=np.float32)
    assert emu.pmp_sdf_evaluate(t2.scene, P32.ctypes.data, 300, out.ctypes.data, None) == 0
    want = sdf.evaluate_batch(P32.astype(np.float64))
    np.testing.assert_allclose(out, want, atol=1e-5, rtol=0)
    r32 = rng.uniform(0.0, 0.2, 300).astype(np.float32)
    flag = np.zeros(300, dtype=np.uint8)
    assert emu.pmp_sdf_is_outside(t2.scene, P32.ctypes.data, r32.ctypes.data, 300, flag.ctypes.data, None) == 0
    far = np.abs(want - r32) > 1e-5
    assert ((flag.astype(bool) == (want > r32)) | ~far).all()


# ---- cases added from the line-coverage report (tools/coverage_emul.sh): every kernel line the suite did
# ---- not reach before — link frames, zero-length and > 128-waypoint edges, the session's staged flat path,
# ---- a ground plane through is_outside, single-leaf cloud trees
def test_emulated_link_frames(emu):
    spec, cst = workloads.make_panda()
    t = EmuTables(emu, cst)
    kin = cst.get_kin()
    Q = workloads.random_configs(spec, 40, 21)
    n_ops = len(t._keep["program"].ops)
    ok = bridge.OracleKin(kin)
    for accurate in (True, False):
        out = np.zeros((len(Q), n_ops, 7), dtype=np.float32)
        assert emu.pmp_fk_frames(t.robot, Q.ctypes.data, len(Q), int(accurate), out.ctypes.data, None) == 0, emu.pmp_last_error()
        for i in (0, 17, 39):
            ok.set_joint_angles(cst._control_joint_ids, Q[i].astype(np.float64), accurate)
            for lid, op in t._keep["program"].op_of_link.items():
                want = ok.link_pose(lid)
                got = out[i, op].astype(np.float64)
                np.testing.assert_allclose(got[:3], want[:3], atol=2e-6, rtol=0)
                assert min(np.abs(got[3:] - want[3:]).max(), np.abs(got[3:] + want[3:]).max()) < 2e-6


def test_emulated_edge_corner_cases(emu):
    """Zero-length edges (valid without a check), edges of more than 128 waypoints (past the sequence
    table: plain forward order) next to ordinary ones, both validators, waypoint counts included."""
    spec, cst = workloads.make_panda(ceiled=True)
    t = EmuTables(emu, cst)
    o = bridge.OracleCst(cst)
    q0, q1 = workloads.random_edges(spec, 24, 8, True, lambda q: o.is_valid_batch(q.astype(np.float64)))
    q1[::4] = q0[::4]  # zero length
    for i in (1, 9):  # long: 0.3 rad over a 0.002 box -> 150+ waypoints
        q1[i] = q0[i] + np.float32(0.3) * np.sign(q1[i] - q0[i] + 1e-9)
    widths = np.full(7, 0.002)
    got, nc = t.check_edges(0, widths, 0.0, q0, q1)
    want, nc_want = bridge.OracleValidator(0, 7, widths).check_motion_batch(o, q0.astype(np.float64), q1.astype(np.float64))
    assert (got == want).all() and nc == nc_want
    assert got[::4].all()
    got, nc = t.check_edges(1, None, 0.004, q0, q1)
    want, nc_want = bridge.OracleValidator(1, 7, None, 0.004).check_motion_batch(o, q0.astype(np.float64), q1.astype(np.float64))
    assert (got == want).all() and nc == nc_want


def test_emulated_session_staged_flat_path(emu):
    """More than 4096 waypoints in one call: the flat kernel reading a staged (copied) arena and
    reporting through the per-edge bit mask instead of the zero-copy bytes."""
    spec, cst = workloads.make_panda(ceiled=True)
    t = EmuTables(emu, cst)
    o = bridge.OracleCst(cst)
    widths = np.full(7, 0.009)
    v, keep = tables.make_validator(0, 7, widths, 0.0)
    ses = C.c_void_p()
    assert emu.pmp_session_create(t.robot, t.scene, C.byref(v), 4096, C.byref(ses)) == 0, emu.pmp_last_error()
    q0, q1 = workloads.random_edges(spec, 110, 12, True, lambda q: o.is_valid_batch(q.astype(np.float64)))
    q0, q1 = np.ascontiguousarray(q0, dtype=np.float64), np.ascontiguousarray(q1, dtype=np.float64)
    out = np.zeros(len(q0), dtype=np.uint8)
    nc = C.c_uint64(0)
    assert emu.pmp_session_check_motion(ses, q0.ctypes.data, q1.ctypes.data, len(q0), out.ctypes.data, C.byref(nc)) == 0
    ov = bridge.OracleValidator(0, 7, widths)
    n_way = sum(int(np.floor(1.0 / ov.step_ratio(a, b))) + 1 for a, b in zip(q0, q1))
    assert n_way > 4096, n_way  # beyond the zero-copy limit: the staged path with per-waypoint result bits
    want, calls, margin = ov.check_motion_batch_ex(o, q0, q1)
    clear = margin > 1e-5
    assert (out.astype(bool) == want)[clear].all()
    # the session reports the REFERENCE's is_valid call count (first failing waypoint in its visiting order)
    assert clear.all() and nc.value == int(calls.sum()), (nc.value, int(calls.sum()), int((~clear).sum()))
    # and more than 2048 states in one is_valid call: staged copy + the bulk validity kernel
    Qd = np.ascontiguousarray(workloads.random_configs(spec, 2100, 13), dtype=np.float64)
    st = np.zeros(len(Qd), dtype=np.uint8)
    assert emu.pmp_session_is_valid(ses, Qd.ctypes.data, len(Qd), st.ctypes.data) == 0
    _parity(st.astype(bool), o.is_valid_batch(Qd), o.margin_batch(Qd), "staged is_valid")
    emu.pmp_session_destroy(ses)


@pytest.mark.parametrize("n_points", [1, 5, 8, 9])
def test_emulated_tiny_clouds_and_ground(emu, n_points):
    """Clouds whose box tree is a single leaf (<= 8 points) or barely more, in every traversal (narrow
    phase, evaluate, nearest point, psdf queries), and a ground plane through is_outside."""
    spec, cst = workloads.make_panda()
    rng = np.random.default_rng(n_points)
    pts = rng.uniform(-0.05, 0.05, (n_points, 3)) + np.array([0.45, 0.0, 0.5])
    cst.set_sdf(psdf.UnionSDF([psdf.GroundSDF(-0.02), psdf.CloudSDF(pts, 0.03)]))
    t = EmuTables(emu, cst)
    o = bridge.OracleCst(cst)
    Q = workloads.random_configs(spec, 96, 6)
    Qd = Q.astype(np.float64)
    want = o.is_valid_batch(Qd)
    _parity(t.check(Q), want, o.margin_batch(Qd), f"{n_points}-point cloud")
    assert 0.0 < want.mean() < 1.0
    np.testing.assert_allclose(t.evaluate(Q[:32]), o.evaluate_batch(Qd[:32]), atol=1e-5, rtol=0)
    P = (np.array([0.45, 0.0, 0.5]) + rng.uniform(-0.6, 0.6, (64, 3))).astype(np.float32)
    sq = np.zeros(64, dtype=np.float32)
    idx = np.zeros(64, dtype=np.int32)
    assert emu.pmp_cloud_nearest(t.scene, P.ctypes.data, 64, sq.ctypes.data, idx.ctypes.data, None) == 0
    brute = ((P[:, None, :].astype(np.float64) - pts[None, :, :]) ** 2).sum(-1)
    np.testing.assert_allclose(sq, brute.min(1), atol=2e-6, rtol=1e-5)
    sdf = bridge.OracleSDF.from_psdf(cst.get_sdf())
    val = np.zeros(64, dtype=np.float32)
    assert emu.pmp_sdf_evaluate(t.scene, P.ctypes.data, 64, val.ctypes.data, None) == 0
    wv = sdf.evaluate_batch(P.astype(np.float64))
    np.testing.assert_allclose(val, wv, atol=1e-5, rtol=0)
    # is_outside: many queries close to the cloud so that the field leaves some for the tree
    P = (np.array([0.45, 0.0, 0.5]) + rng.uniform(-0.25, 0.25, (512, 3))).astype(np.float32)
    wv = sdf.evaluate_batch(P.astype(np.float64))
    r32 = (np.maximum(wv, 0.0) + rng.uniform(-0.02, 0.02, 512)).clip(0.0).astype(np.float32)
    flag = np.zeros(512, dtype=np.uint8)
    assert emu.pmp_sdf_is_outside(t.scene, P.ctypes.data, r32.ctypes.data, 512, flag.ctypes.data, None) == 0
    far = np.abs(wv - r32) > 1e-5
    assert ((flag.astype(bool) == (wv > r32)) | ~far).all()
    assert 0.2 < flag.mean() < 0.8


@pytest.mark.parametrize("seed", [9001, 9004, 9006, 9007, 9010, 9013, 9016, 9019])
def test_emulated_edges_under_queue_stress(emu, seed):
    """The lane = edge kernel and the session's flat kernel on the queue-overflow scenes: every lane
    carries a different waypoint of a different edge while single groups overflow the warp queue."""
    cst, lb, ub, desc = random_case(seed)
    dim = len(lb)
    if dim == 0:
        pytest.skip("no degrees of freedom")
    rng = np.random.default_rng(seed)
    cst.set_sdf(stress_scene(rng, bool(seed % 3 == 0)))
    t = EmuTables(emu, cst)
    o = bridge.OracleCst(cst)
    widths = rng.uniform(0.02, 0.2, dim)
    q0 = sample_q(lb, ub, 64, seed + 1)
    q1 = np.clip(q0.astype(np.float64) + rng.uniform(-0.5, 0.5, q0.shape), lb, ub).astype(np.float32)
    got, nc = t.check_edges(0, widths, 0.0, q0, q1)
    want = _edge_parity(got, nc, o, bridge.OracleValidator(0, dim, widths, 0.0), q0.astype(np.float64), q1.astype(np.float64), desc)
    v, keep = tables.make_validator(0, dim, widths, 0.0)
    ses = C.c_void_p()
    assert emu.pmp_session_create(t.robot, t.scene, C.byref(v), 64, C.byref(ses)) == 0
    a, b = np.ascontiguousarray(q0, dtype=np.float64), np.ascontiguousarray(q1, dtype=np.float64)
    out = np.zeros(64, dtype=np.uint8)
    nck = C.c_uint64(0)
    assert emu.pmp_session_check_motion(ses, a.ctypes.data, b.ctypes.data, 64, out.ctypes.data, C.byref(nck)) == 0
    emu.pmp_session_destroy(ses)
    _edge_parity(out, nck.value, o, bridge.OracleValidator(0, dim, widths, 0.0), a, b, desc + " (session)")


# ---- branches the gcov branch report showed untaken --------------------------------------------------
def test_emulated_joint_angles_beyond_pi(emu):
    """Joint angles up to +-4 pi on a FIXED base: the range reduction of the approximate sin/cos
    (kinematics/algorithm.cpp:49-86) in all its branches, against the oracle's FP64 restatement."""
    spec, cst = workloads.make_panda()
    t = EmuTables(emu, cst)
    o = bridge.OracleCst(cst)
    rng = np.random.default_rng(17)
    Q = rng.uniform(-4 * np.pi, 4 * np.pi, (192, 7)).astype(np.float32)
    Qd = Q.astype(np.float64)
    _parity(t.check(Q), o.is_valid_batch(Qd), o.margin_batch(Qd), "angles beyond pi")
    np.testing.assert_allclose(t.fk_spheres(Q[:48], False), o.spheres_batch(Qd[:48], False), atol=2e-5, rtol=0)


def test_emulated_ground_as_nearest_primitive_and_optional_arguments(emu):
    """evaluate() with Jacobian rows where the ground plane is the arg-min primitive; is_outside without
    radii; nearest-point queries without the index output; duplicated cloud points (distance ties)."""
    spec, cst = workloads.make_panda()
    rng = np.random.default_rng(23)
    base = rng.uniform(-0.05, 0.05, (40, 3)) + np.array([0.5, 0.0, 0.6])
    pts = np.concatenate([base, base[:15], base[:5]])  # exact duplicates
    cst.set_sdf(psdf.UnionSDF([psdf.GroundSDF(0.08), psdf.CloudSDF(pts, 0.01)]))
    t = EmuTables(emu, cst)
    o = bridge.OracleCst(cst)
    Q = workloads.random_configs(spec, 96, 31)
    Qd = Q.astype(np.float64)
    vals, jac = t.evaluate_grad(Q)
    want_v, want_j = o.evaluate_grad_batch(Qd)
    np.testing.assert_allclose(vals, want_v, atol=1e-5, rtol=0)
    assert_jacobian_parity(o, Q, jac, want_j, "ground / duplicated cloud points", max_fraction=0.05)
    P = (np.array([0.5, 0.0, 0.4]) + rng.uniform(-0.5, 0.5, (200, 3))).astype(np.float32)
    sdf = bridge.OracleSDF.from_psdf(cst.get_sdf())
    wv = sdf.evaluate_batch(P.astype(np.float64))
    flag = np.zeros(200, dtype=np.uint8)
    assert emu.pmp_sdf_is_outside(t.scene, P.ctypes.data, None, 200, flag.ctypes.data, None) == 0
    assert ((flag.astype(bool) == (wv > 0)) | (np.abs(wv) < 1e-5)).all()
    sq = np.zeros(200, dtype=np.float32)
    assert emu.pmp_cloud_nearest(t.scene, P.ctypes.data, 200, sq.ctypes.data, None, None) == 0
    brute = ((P[:, None, :].astype(np.float64) - pts[None, :, :]) ** 2).sum(-1)
    np.testing.assert_allclose(sq, brute.min(1), atol=2e-6, rtol=1e-5)
    idx = np.zeros(200, dtype=np.int32)
    assert emu.pmp_cloud_nearest(t.scene, P.ctypes.data, 200, sq.ctypes.data, idx.ctypes.data, None) == 0
    np.testing.assert_allclose(brute[np.arange(200), idx], brute.min(1), atol=2e-6, rtol=1e-5)
    # values only, with a primitive nearer than the cloud for most configurations
    np.testing.assert_allclose(t.evaluate(Q), o.evaluate_batch(Qd), atol=1e-5, rtol=0)
    # psdf queries on a scene without a cloud; edges without the optional counters
    spec2, cst2 = workloads.make_panda(ceiled=True)
    t2 = EmuTables(emu, cst2)
    sdf2 = bridge.OracleSDF.from_psdf(cst2.get_sdf())
    val = np.zeros(200, dtype=np.float32)
    assert emu.pmp_sdf_evaluate(t2.scene, P.ctypes.data, 200, val.ctypes.data, None) == 0
    wv2 = sdf2.evaluate_batch(P.astype(np.float64))
    np.testing.assert_allclose(val, wv2, atol=1e-5, rtol=0)
    assert emu.pmp_sdf_is_outside(t2.scene, P.ctypes.data, None, 200, flag.ctypes.data, None) == 0
    assert ((flag.astype(bool) == (wv2 > 0)) | (np.abs(wv2) < 1e-5)).all()
    o2 = bridge.OracleCst(cst2)
    q0, q1 = workloads.random_edges(spec2, 40, 5, True, lambda q: o2.is_valid_batch(q.astype(np.float64)))
    vd, keep = tables.make_validator(0, 7, np.full(7, 0.05), 0.0)
    bits = np.zeros(2, dtype=np.uint32)
    assert emu.pmp_check_edges_host(t2.robot, t2.scene, C.byref(vd), q0.ctypes.data, q1.ctypes.data, 40, bits.ctypes.data, None, None) == 0
    assert (tables.unpack_bits(bits, 40) == t2.check_edges(0, np.full(7, 0.05), 0.0, q0, q1)[0]).all()


def test_emulated_session_more_edges_than_the_offset_table(emu):
    """More than 1024 (short) edges in one call: the flat kernel searches the waypoint offsets in global
    memory instead of the shared-memory copy."""
    spec, cst = workloads.make_panda(ceiled=True)
    t = EmuTables(emu, cst)
    o = bridge.OracleCst(cst)
    widths = np.full(7, 0.05)
    v, keep = tables.make_validator(0, 7, widths, 0.0)
    ses = C.c_void_p()
    assert emu.pmp_session_create(t.robot, t.scene, C.byref(v), 2048, C.byref(ses)) == 0, emu.pmp_last_error()
    rng = np.random.default_rng(3)
    q0 = workloads.random_configs(spec, 1100, 41).astype(np.float64)
    q1 = q0 + rng.uniform(-0.06, 0.06, q0.shape)
    out = np.zeros(len(q0), dtype=np.uint8)
    nc = C.c_uint64(0)
    assert emu.pmp_session_check_motion(ses, q0.ctypes.data, q1.ctypes.data, len(q0), out.ctypes.data, C.byref(nc)) == 0
    want = _edge_parity(out, nc.value, o, bridge.OracleValidator(0, 7, widths), q0, q1, "session, > 1024 edges")
    assert 0.1 < want.mean() < 0.9
    emu.pmp_session_destroy(ses)


def test_emulated_abi_argument_errors(emu):
    """The C ABI's argument validation (status code + pmp_last_error text), reachable on the CPU through
    the emulated build: malformed robot / scene tables, validators and calls are refused before any launch."""
    import copy

    spec, cst = workloads.make_panda()
    t = EmuTables(emu, cst)

    def refused(rc, text):
        assert rc != 0 and text in emu.pmp_last_error().decode(), (rc, emu.pmp_last_error())

    out = C.c_void_p()
    refused(emu.pmp_robot_create(None, C.byref(out)), "null argument")
    for field, value, text in (
        ("n_ops", -1, "negative table size"),
        ("base_type", 7, "bad base_type"),
        ("n_groups", 4096, "too many sphere groups"),
    ):
        d = _capi.RobotDesc.from_buffer_copy(t._desc)
        setattr(d, field, value)
        refused(emu.pmp_robot_create(C.byref(d), C.byref(out)), text)
    a = t._keep
    for name, idx, field, value, text in (
        ("groups", 0, "sphere_count", 10**6, "group sphere range out of bounds"),
        ("groups", 1, "op", 999, "group refers to an unknown op"),
    ):
        arr = a[name].copy()
        arr[idx][field] = value
        d = _capi.RobotDesc.from_buffer_copy(t._desc)
        d.groups = arr.ctypes.data
        refused(emu.pmp_robot_create(C.byref(d), C.byref(out)), text)
    ops = a["program"].ops.copy()
    ops[3]["parent"] = 5
    d = _capi.RobotDesc.from_buffer_copy(t._desc)
    d.ops = ops.ctypes.data
    refused(emu.pmp_robot_create(C.byref(d), C.byref(out)), "topological order")
    pairs = a["pairs"].copy()
    if len(pairs):
        pairs.reshape(-1)[0] = 500
        d = _capi.RobotDesc.from_buffer_copy(t._desc)
        d.pairs = pairs.ctypes.data
        refused(emu.pmp_robot_create(C.byref(d), C.byref(out)), "pair refers to an unknown group")
    assert emu.pmp_robot_q_dim(None) == -1 and emu.pmp_robot_q_dim(t.robot) == 7
    assert emu.pmp_robot_n_spheres(t.robot) == len(a["spheres"])
    # scenes
    sd = _capi.SceneDesc()
    sd.n_prims = -1
    refused(emu.pmp_scene_create(C.byref(sd), C.byref(out)), "negative table size")
    prims = t._prims.copy()
    prims[0]["type"] = 9
    sd = _capi.SceneDesc()
    sd.n_prims, sd.prims = len(prims), prims.ctypes.data
    refused(emu.pmp_scene_create(C.byref(sd), C.byref(out)), "bad primitive type")
    sd = _capi.SceneDesc()
    sd.n_cloud_points, sd.cloud_points, sd.cloud_radius, sd.cloud_cell = 10, None, 0.01, 0.02
    refused(emu.pmp_scene_create(C.byref(sd), C.byref(out)), "cloud_points is null")
    bad = np.zeros((10, 3), dtype=np.float32)
    bad[4, 1] = np.nan
    sd.cloud_points = bad.ctypes.data
    refused(emu.pmp_scene_create(C.byref(sd), C.byref(out)), "non-finite")
    # validators and calls
    Q = workloads.random_configs(spec, 8, 1)
    bits = np.zeros(1, dtype=np.uint32)
    v, keep = tables.make_validator(0, 7, np.full(7, 0.05), 0.0)
    v.type = 5
    refused(emu.pmp_check_edges_host(t.robot, t.scene, C.byref(v), Q.ctypes.data, Q.ctypes.data, 8, bits.ctypes.data, None, None), "unknown validator type")
    v, keep = tables.make_validator(1, 7, None, 0.1)
    v.resolution = 0.0
    refused(emu.pmp_check_edges_host(t.robot, t.scene, C.byref(v), Q.ctypes.data, Q.ctypes.data, 8, bits.ctypes.data, None, None), "resolution must be positive")
    refused(emu.pmp_check_configs_host(None, t.scene, Q.ctypes.data, 8, bits.ctypes.data, None), "")
    refused(emu.pmp_check_configs_host(t.robot, t.scene, Q.ctypes.data, -3, bits.ctypes.data, None), "")
    refused(emu.pmp_check_configs_host(t.robot, t.scene, None, 8, bits.ctypes.data, None), "")
    ses = C.c_void_p()
    v.type = 5
    refused(emu.pmp_session_create(t.robot, t.scene, C.byref(v), 64, C.byref(ses)), "unknown validator type")
    # and the handles still work afterwards
    assert t.check(Q).shape == (8,)


def test_emulated_composite_constraint_single_fk(emu):
    """IneqCompositeCst (composite_constraint.hpp:27-44,98-105): two collision constraints on one kinematic
    model, each with its own scene, merged into ONE table with per-group scene masks — one kernel, one FK
    pass — against the oracle's composite (pinned to the compiled reference's by tests/test_reference_pin.py)
    and against the AND of the members checked one by one."""
    from plainmp_b200.constraint import IneqCompositeCst

    spec, a = workloads.make_panda()
    b = spec.create_collision_const(self_collision=False, use_cache=False)
    b.set_sdf(psdf.UnionSDF([psdf.SphereSDF(0.25, psdf.Pose([0.45, 0.1, 0.55])), psdf.BoxSDF([0.3, 0.3, 0.3], psdf.Pose([-0.5, 0.3, 0.4], rpy_matrix(0.3, 0.2, 0.1)))]))
    comp = IneqCompositeCst([a, b])
    t = EmuTables(emu, comp)
    oc = bridge.OracleComposite(comp)
    Q = workloads.random_configs(spec, 384, 19)
    Qd = Q.astype(np.float64)
    got, want = t.check(Q), oc.is_valid_batch(Qd)
    _parity(got, want, oc.margin_batch(Qd), "composite")
    sep = EmuTables(emu, a).check(Q) & EmuTables(emu, b).check(Q)
    assert (got == sep).all()  # the mask keeps member b's groups away from member a's primitives and vice versa
    assert (want != bridge.OracleCst(a).is_valid_batch(Qd)).any()  # member b matters
    # edges through the same merged table
    q0, q1 = Q[:64], np.clip(Qd[:64] + np.random.default_rng(2).uniform(-0.4, 0.4, (64, 7)), *spec.angle_bounds()).astype(np.float32)
    e_got, _ = t.check_edges(0, np.full(7, 0.05), 0.0, q0, q1)
    e_sep = EmuTables(emu, a).check_edges(0, np.full(7, 0.05), 0.0, q0, q1)[0] & EmuTables(emu, b).check_edges(0, np.full(7, 0.05), 0.0, q0, q1)[0]
    assert (e_got == e_sep).all()


@pytest.mark.parametrize("n_side", [10, 11])
def test_emulated_huge_sphere_group_unit_offsets(emu, n_side):
    """create_attached_box_collision_const-style grids (robot_spec.py:561-595: 10 x 10 x 10 zero-radius
    spheres on the hand).  1000 spheres per group stay on the flat work-unit list, whose 16-bit unit offsets
    then force early drains (32 lanes x 1000 units > 60000 would not fit: the push is split); 1331 spheres
    exceed kBigGroup and take the 8-lane path instead."""
    spec = PandaSpec()
    g = np.linspace(-0.06, 0.06, n_side)
    pts = np.array([[x, y, z + 0.18] for x in g for y in g for z in g]).T
    att = SphereAttachmentSpec("panda_link8", pts, np.zeros(pts.shape[1]), False)
    cst = spec.create_collision_const(attachments=[att], use_cache=False, reorder_spheres=False)
    # plates whose AABBs cover the whole workspace: every lane queues the big group against every plate
    plates = [psdf.BoxSDF([0.02, 4.0, 4.0], psdf.Pose([1.6 + 0.1 * k, 0.0, 0.5])) for k in range(3)]
    for k, pl in enumerate(plates):
        pl.rotate_z(0.3 + 0.9 * k)
    cst.set_sdf(psdf.UnionSDF(plates + [psdf.BoxSDF([0.3, 0.3, 0.3], psdf.Pose([0.45, 0.0, 0.5]))]))
    t = EmuTables(emu, cst)
    o = bridge.OracleCst(cst)
    Q = workloads.random_configs(spec, 96, 7)
    Qd = Q.astype(np.float64)
    want = o.is_valid_batch(Qd)
    assert 0.1 < want.mean() < 0.9
    _parity(t.check(Q), want, o.margin_batch(Qd), f"{n_side ** 3}-sphere group")


def test_emulated_fp64_host_entry_points(emu):
    """pmp_check_configs_host_f64 / pmp_check_edges_host_f64: the reference's callers hold FP64 states; the
    packing threads + staged copies must give the words of the FP32 entry points on the same values."""
    spec, cst = workloads.make_panda(ceiled=True)
    t = EmuTables(emu, cst)
    Q = workloads.random_configs(spec, 300, 23)
    Qd = np.ascontiguousarray(Q, dtype=np.float64)
    want = t.check(Q)
    bits = np.zeros((300 + 31) // 32, dtype=np.uint32)
    nv = C.c_uint64(0)
    assert emu.pmp_check_configs_host_f64(t.robot, t.scene, Qd.ctypes.data, 300, bits.ctypes.data, C.byref(nv)) == 0, emu.pmp_last_error()
    assert (tables.unpack_bits(bits, 300) == want).all() and nv.value == int(want.sum())
    # a page-locked FP64 source takes the other path: the doubles are copied as they are and narrowed on the device
    os.environ["PMP_EMUL_PINNED"] = "1"
    try:
        bits[:] = 0
        assert emu.pmp_check_configs_host_f64(t.robot, t.scene, Qd.ctypes.data, 300, bits.ctypes.data, C.byref(nv)) == 0, emu.pmp_last_error()
    finally:
        del os.environ["PMP_EMUL_PINNED"]
    assert (tables.unpack_bits(bits, 300) == want).all() and nv.value == int(want.sum())
    q0, q1 = workloads.random_edges(spec, 96, 4, True, lambda q: t.check(q))
    e32, nc32 = t.check_edges(0, np.full(7, 0.05), 0.0, q0, q1)
    v, keep = tables.make_validator(0, 7, np.full(7, 0.05), 0.0)
    a, b = np.ascontiguousarray(q0, dtype=np.float64), np.ascontiguousarray(q1, dtype=np.float64)
    ebits = np.zeros(3, dtype=np.uint32)
    nv, nc = C.c_uint64(0), C.c_uint64(0)
    assert emu.pmp_check_edges_host_f64(t.robot, t.scene, C.byref(v), a.ctypes.data, b.ctypes.data, 96, ebits.ctypes.data, C.byref(nv), C.byref(nc)) == 0
    assert (tables.unpack_bits(ebits, 96) == e32).all() and nc.value == nc32 and nv.value == int(e32.sum())


# ---- round 2: JaxonSpec, PR2BaseOnlySpec, the attached-box lattice, several CloudSDFs in one scene ----
def _round2_case(name):
    if name == "jaxon":
        spec, cst = workloads.make_jaxon()
        return cst, workloads.random_configs_moving_base(spec, 192, 11, [-0.5, -1.0, 0.7, -0.3, -0.3, -1.0], [1.2, 1.0, 1.3, 0.3, 0.3, 1.0])
    if name == "pr2_base_only":
        spec, cst = workloads.make_pr2_base_only()
        return cst, workloads.random_configs_moving_base(spec, 192, 12, [-2.0, -2.5, -3.2], [2.5, 2.5, 3.2])
    if name == "position_only_link":
        spec, cst = workloads.make_panda_position_only_link()
        return cst, workloads.random_configs(spec, 192, 15)
    if name == "attached_box":
        spec = PandaSpec()
        box = psdf.BoxSDF([0.1, 0.06, 0.2], psdf.Pose([0.3, 0.1, 0.5], rpy_matrix(0.7, 0.2, -0.4)))
        cst = spec.create_attached_box_collision_const(box, "panda_hand", np.array([0.0, 0.0, 0.15]))
        cst.set_sdf(workloads.panda_dual_bars_sdf())
        return cst, workloads.random_configs(spec, 192, 13)
    from plainmp_b200.robot_spec import PR2DualarmSpec

    spec = PR2DualarmSpec()
    spec.reflect_default_joint_positions()
    cst = spec.create_collision_const(use_cache=False)
    pts = workloads.table_cloud(3000, 4)
    shelf = np.random.default_rng(0).uniform([-0.2, -1.0, 0.2], [0.4, -0.5, 1.5], size=(800, 3))
    r0, r1, r2 = (0.002, 0.03, 0.011) if name == "clouds_of_different_radii" else (0.002, 0.002, 0.002)
    cst.set_sdf(psdf.UnionSDF([psdf.CloudSDF(pts[:1500], r0), psdf.BoxSDF([0.3, 0.3, 0.3], psdf.Pose([0.6, 0.6, 1.2])),
                               psdf.UnionSDF([psdf.CloudSDF(pts[1500:], r1), psdf.CloudSDF(shelf, r2)])]))
    return cst, workloads.random_configs(spec, 128, 14)


@pytest.mark.parametrize("name", ["jaxon", "pr2_base_only", "attached_box", "several_clouds", "clouds_of_different_radii", "position_only_link"])
def test_emulated_round2_specs_and_scenes(emu, name):
    cst, Q = _round2_case(name)
    t = EmuTables(emu, cst)
    o = bridge.OracleCst(cst)
    Qd = Q.astype(np.float64)
    want = o.is_valid_batch(Qd)
    assert 0.03 < want.mean() < 0.97, want.mean()
    _parity(t.check(Q), want, o.margin_batch(Qd), name)
    vals, jac = t.evaluate_grad(Q[:96])
    want_v, want_j = o.evaluate_grad_batch(Qd[:96])
    np.testing.assert_allclose(vals, want_v, atol=1e-5, rtol=0, equal_nan=True, err_msg=name)
    assert_jacobian_parity(o, Q[:96], jac, want_j, name, max_fraction=0.05)
    np.testing.assert_allclose(t.fk_spheres(Q[:32], False), o.spheres_batch(Qd[:32], False), atol=2e-5, rtol=0)


def test_emulated_long_edges_take_cooperative_rounds(emu):
    """Edges of 30 - 200 waypoints, most of them valid or failing late: after kCoopAfter serial waypoints the
    whole warp works such an edge 32 positions of the visiting order per round.  Validity AND the is_valid
    call count must still be the serial walk's (first failing position in the reference's order)."""
    spec, cst = workloads.make_panda(ceiled=True)
    t = EmuTables(emu, cst)
    o = bridge.OracleCst(cst)
    rng = np.random.default_rng(41)
    q0, _ = workloads.random_edges(spec, 96, 9, True, lambda q: o.is_valid_batch(q.astype(np.float64)))
    lb, ub = spec.angle_bounds()
    # long moves of few joints: many stay valid for dozens of waypoints, some run into the bars late
    step = rng.uniform(-1.2, 1.2, q0.shape) * (rng.random(q0.shape) < 0.4)
    q1 = np.clip(q0.astype(np.float64) + step, lb, ub).astype(np.float32)
    for widths in (np.full(7, 0.02), np.full(7, 0.006)):  # up to ~60 and ~200 waypoints (beyond 128: natural order)
        ov = bridge.OracleValidator(0, 7, widths)
        want, calls, _ = ov.check_motion_batch_ex(o, q0.astype(np.float64), q1.astype(np.float64), 4, with_margin=False)
        assert (calls > 12 + 20).sum() >= 20 and 0.2 < want.mean() < 0.9, (calls.max(), want.mean())
        got, n_checks = t.check_edges(0, widths, 0.0, q0, q1)
        _edge_parity(got, n_checks, o, ov, q0.astype(np.float64), q1.astype(np.float64), f"long edges, width {widths[0]}")


def test_emulated_radius_classes_psdf_surface_edges_and_update(emu):
    """CloudSDF leaves of different radii behind one device index (radius classes): the psdf surface
    (evaluate / is_outside of the union), edges, and a cloud update that changes the class layout."""
    spec, cst = workloads.make_panda()
    rng = np.random.default_rng(3)
    a = rng.uniform(-0.12, 0.12, (500, 3)) + np.array([0.45, 0.0, 0.55])
    b = rng.uniform(-0.10, 0.10, (300, 3)) + np.array([0.25, 0.35, 0.35])
    c = rng.uniform(-0.15, 0.15, (200, 3)) + np.array([-0.3, -0.3, 0.6])
    union = psdf.UnionSDF([psdf.CloudSDF(a, 0.004), psdf.BoxSDF([0.2, 0.2, 0.2], psdf.Pose([0.5, -0.4, 0.2])), psdf.CloudSDF(b, 0.05), psdf.CloudSDF(c, 0.02)])
    cst.set_sdf(union)
    t = EmuTables(emu, cst)
    o = bridge.OracleCst(cst)
    sdf = bridge.OracleSDF.from_psdf(union)
    P = (np.array([0.2, 0.0, 0.45]) + rng.uniform(-0.6, 0.6, (300, 3))).astype(np.float32)
    want_v = sdf.evaluate_batch(P.astype(np.float64))
    got_v = np.zeros(300, dtype=np.float32)
    assert emu.pmp_sdf_evaluate(t.scene, P.ctypes.data, 300, got_v.ctypes.data, None) == 0, emu.pmp_last_error()
    np.testing.assert_allclose(got_v, want_v, atol=1e-5, rtol=0)
    for r in (0.0, 0.03):
        rad = np.full(300, r, dtype=np.float32)
        flag = np.zeros(300, dtype=np.uint8)
        assert emu.pmp_sdf_is_outside(t.scene, P.ctypes.data, rad.ctypes.data, 300, flag.ctypes.data, None) == 0
        assert ((flag.astype(bool) == (want_v > r)) | (np.abs(want_v - r) < 1e-5)).all()
    Q = workloads.random_configs(spec, 96, 21)
    Qd = Q.astype(np.float64)
    _parity(t.check(Q), o.is_valid_batch(Qd), o.margin_batch(Qd), "radius classes")
    q0, q1 = workloads.random_edges(spec, 48, 6, True, lambda q: o.is_valid_batch(q.astype(np.float64)))
    widths = np.full(7, 0.05)
    got, n_checks = t.check_edges(0, widths, 0.0, q0, q1)
    _edge_parity(got, n_checks, o, bridge.OracleValidator(0, 7, widths), q0.astype(np.float64), q1.astype(np.float64), "radius classes, edges")
    # a new frame with another class layout (two radii, other order) through the update entry point
    pts = np.ascontiguousarray(np.vstack([c, a]), dtype=np.float32)
    radii = np.concatenate([np.full(len(c), 0.03, dtype=np.float32), np.full(len(a), 0.01, dtype=np.float32)])
    assert emu.pmp_scene_update_cloud_ex(t.scene, pts.ctypes.data, len(pts), 0.0, radii.ctypes.data, 0.02) == 0, emu.pmp_last_error()
    cst.set_sdf(psdf.UnionSDF([psdf.BoxSDF([0.2, 0.2, 0.2], psdf.Pose([0.5, -0.4, 0.2])), psdf.CloudSDF(c, 0.03), psdf.CloudSDF(a, 0.01)]))
    o2 = bridge.OracleCst(cst)
    _parity(t.check(Q), o2.is_valid_batch(Qd), o2.margin_batch(Qd), "radius classes after an update")
    np.testing.assert_allclose(t.evaluate(Q[:48]), o2.evaluate_batch(Qd[:48]), atol=1e-5, rtol=0)
    five = np.arange(len(pts), dtype=np.float32) % 5 * 0.01 + 0.01
    assert emu.pmp_scene_update_cloud_ex(t.scene, pts.ctypes.data, len(pts), 0.0, five.ctypes.data, 0.02) == _capi.PMP_ERR_LIMIT
