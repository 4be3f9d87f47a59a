"""Differential test on RANDOM robots and scenes (tests/fuzz_util.py): the CUDA path against the
compiled reference where it travelled to the box, else the restated oracle (pinned to it by
tests/test_fuzz_cpu.py on the same seeds)."""
import os

import numpy as np
import pytest

from oracle import bridge, ref_bridge
from fuzz_util import random_case, sample_q
from parity_util import assert_edge_parity, assert_jacobian_parity
from plainmp_b200.kinematics import BaseType
from plainmp_b200.ompl import MotionValidator, ValidatorConfig, ValidatorSession, ValidatorType

# PMP_FUZZ_SEEDS=start:count widens the sweep for a one-off hunt (default: 40 fixed seeds)
_s, _c = (int(x) for x in os.environ.get("PMP_FUZZ_SEEDS", "1000:40").split(":"))
SEEDS = range(_s, _s + _c)
BIG_SEEDS = range(3000, 3000 + max(12, _c // 4))

pytestmark = pytest.mark.gpu

EPS_BAND = 1e-5
TOL = 2e-5  # random chains reach 3-4 m with up to 12 links: FP32 accumulates a little more than on the stock arms


@pytest.mark.parametrize("seed", SEEDS)
def test_cuda_equals_reference_on_random_robot(cuda_lib, seed):
    cst, lb, ub, desc = random_case(seed)
    o = bridge.OracleCst(cst, n_threads=4)
    r = ref_bridge.RefCst(cst) if ref_bridge.available() else o
    Q = sample_q(lb, ub, 3000, seed)
    Qd = Q.astype(np.float64)
    got, want = cst.is_valid_batch(Q), r.is_valid_batch(Qd)
    margin = o.margin_batch(Qd)
    bad = [int(i) for i in np.where(got != want)[0] if margin[i] > EPS_BAND]
    assert not bad, f"{desc}: {len(bad)} validity mismatches outside the band, e.g. {bad[:4]} margins {margin[bad[:4]]}"
    np.testing.assert_allclose(cst.evaluate_batch(Q[:1000]), r.evaluate_batch(Qd[:1000]), atol=TOL, rtol=0, equal_nan=True, err_msg=desc)
    for accurate in (False, True):
        np.testing.assert_allclose(cst.get_all_spheres_batch(Q[:16], accurate=accurate), r.spheres_batch(Qd[:16], accurate), atol=TOL, rtol=0, err_msg=desc)
    if ref_bridge.available():
        # the drop-in chain on the same robot: reference objects -> C++ flatten -> C ABI -> CUDA
        from plainmp_b200 import tables

        raw = tables.RawTables(r.flatten())
        assert (tables.check_configs(raw.robot, raw.scene, Q) == got).all(), desc
    # Jacobian rows of evaluate(): analytic here, forward differences in the reference; rows may differ
    # where the arg-min is nearly tied or the finite difference straddles a kink of the SDF
    vals, jac = cst.evaluate_grad_batch(Q[:600])
    want_v, want_j = o.evaluate_grad_batch(Qd[:600])
    np.testing.assert_allclose(vals, want_v, atol=TOL, rtol=0, equal_nan=True, err_msg=desc)
    if jac.size:
        # 1e-4 (the reference's own 4-decimal check); a row that differs must sit on a tie / kink (parity_util)
        assert_jacobian_parity(o, Q[:600], jac, want_j, desc, max_fraction=0.05)


@pytest.mark.parametrize("seed", range(1000, 1012))
def test_edges_and_session_on_random_robot(cuda_lib, seed):
    cst, lb, ub, desc = random_case(seed)
    dim = len(lb)
    if dim == 0:
        pytest.skip("no degrees of freedom")
    o = bridge.OracleCst(cst, n_threads=4)
    rng = np.random.default_rng(seed)
    widths = rng.uniform(0.03, 0.2, dim)
    q0 = sample_q(lb, ub, 1500, seed + 1)
    q1 = np.clip(q0.astype(np.float64) + rng.uniform(-0.6, 0.6, q0.shape), lb, ub).astype(np.float32)
    for vtype, res in ((ValidatorType.BOX, 0.0), (ValidatorType.EUCLIDEAN, 0.07)):
        cfg = ValidatorConfig(vtype, res, list(widths))
        got, n_checked, calls = MotionValidator(cst, cfg).check_motion_batch(q0, q1, with_calls=True)
        assert_edge_parity(cst, int(vtype), widths, res, q0, q1, got, n_checked, calls, f"{desc} {vtype.name}")
        ses = ValidatorSession(cst, cfg, capacity=512)
        got_s = ses.check_motion(q0.astype(np.float64), q1.astype(np.float64))
        assert_edge_parity(cst, int(vtype), widths, res, q0, q1, got_s, ses.is_valid_call_count, None, f"{desc} {vtype.name} (session)")
        ses.close()


@pytest.mark.parametrize("seed", BIG_SEEDS)
def test_cuda_equals_reference_on_big_random_robot(cuda_lib, seed):
    """20-45 links, up to 35 sphere groups, dozens of pairs, 28-44 primitives: the 64-thread CTAs, the
    queue-overflow path of phase 1 and the >30-primitive overflow item."""
    cst, lb, ub, desc = random_case(seed, big=True)
    o = bridge.OracleCst(cst, n_threads=4)
    r = ref_bridge.RefCst(cst) if ref_bridge.available() else o
    Q = sample_q(lb, ub, 1500, seed)
    Qd = Q.astype(np.float64)
    got, want = cst.is_valid_batch(Q), r.is_valid_batch(Qd)
    margin = o.margin_batch(Qd)
    bad = [int(i) for i in np.where(got != want)[0] if margin[i] > 2 * EPS_BAND]  # chains of 30+ links
    assert not bad, f"{desc}: {len(bad)} validity mismatches outside the band, e.g. {bad[:4]} margins {margin[bad[:4]]}"
    np.testing.assert_allclose(cst.evaluate_batch(Q[:300]), r.evaluate_batch(Qd[:300]), atol=5e-5, rtol=0, equal_nan=True, err_msg=desc)


@pytest.mark.parametrize("mode", ["", "2"])
def test_queue_overflow_stress_on_random_robots(cuda_lib, mode):
    """Scenes of fuzz_util.stress_scene (every group passes the cull test of 9-27 plates on nearly every
    lane, so single groups overflow the 128-entry warp queue) on random robots of every base type, in
    the default and the FP16 group-centre stash layout.  In a child process with a limit: a kernel that
    stops making progress fails this test instead of hanging the suite (tests/test_simt_emul.py runs
    the same cases on the emulated kernels in the CPU tier)."""
    import subprocess
    import sys
    import textwrap
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    code = textwrap.dedent(
        f"""
        import sys
        sys.path.insert(0, {str(root)!r}); sys.path.insert(0, {str(root / 'tests')!r})
        import numpy as np
        from oracle import bridge
        from fuzz_util import random_case, sample_q, stress_scene
        for seed in range(9000, 9024):
            cst, lb, ub, desc = random_case(seed)
            cst.set_sdf(stress_scene(np.random.default_rng(seed), seed % 3 == 0))
            Q = sample_q(lb, ub, 4096, seed)
            o = bridge.OracleCst(cst, n_threads=4)
            Qd = Q.astype(np.float64)
            got, want, m = cst.is_valid_batch(Q), o.is_valid_batch(Qd), o.margin_batch(Qd)
            bad = [int(i) for i in np.where(got != want)[0] if m[i] > 1e-5]
            assert not bad, (desc, bad[:4])
            if len(lb) and seed % 2 == 0:  # the lane = edge kernel on the same scene
                from plainmp_b200.ompl import MotionValidator, ValidatorConfig, ValidatorType
                rng = np.random.default_rng(seed)
                widths = rng.uniform(0.02, 0.2, len(lb))
                q1 = np.clip(Qd[:1024] + rng.uniform(-0.5, 0.5, (1024, len(lb))), lb, ub).astype(np.float32)
                from parity_util import assert_edge_parity
                e_got, e_n, e_calls = MotionValidator(cst, ValidatorConfig(ValidatorType.BOX, 0.0, list(widths))).check_motion_batch(Q[:1024], q1, with_calls=True)
                assert_edge_parity(cst, 0, widths, 0.0, Q[:1024], q1, e_got, e_n, e_calls, desc)
        print("OK")
        """
    )
    env = dict(os.environ)
    if mode:
        env["PMP_V3_MODE"] = mode
    try:
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=420, env=env)
    except subprocess.TimeoutExpired:
        pytest.fail("the validity kernel did not return within 420 s on a queue-overflow stress scene")
    assert r.returncode == 0 and "OK" in r.stdout, r.stderr[-2000:]


@pytest.mark.parametrize("seed", range(4000, 4020))
def test_psdf_and_kdtree_surface_on_random_scenes(cuda_lib, seed):
    """The psdf / kdtree binding surface (evaluate_batch, is_outside, KDTree.sqdist / query) on random
    unions and on degenerate clouds, against the compiled reference (else the pinned oracle)."""
    from fuzz_util import random_scene
    from plainmp_b200 import psdf
    from plainmp_b200.kdtree import KDTree

    rng = np.random.default_rng(seed)
    sdf = random_scene(rng, with_cloud=bool(seed % 2))
    if sdf.get_type() != psdf.T_UNION:
        sdf = psdf.UnionSDF([sdf])
    R = ref_bridge.RefSDF.from_psdf(sdf) if ref_bridge.available() else bridge.OracleSDF.from_psdf(sdf)
    P = rng.uniform(-1.6, 1.6, size=(4000, 3)).astype(np.float32).astype(np.float64)
    want = R.evaluate_batch(P)
    np.testing.assert_allclose(sdf.evaluate_batch(P.T), want, atol=1e-5, rtol=0)
    radii = np.where(rng.random(4000) < 0.25, 0.0, rng.uniform(0.0, 0.3, 4000))
    want_out = R.is_outside_batch(P, radii)
    got_out = np.array([sdf.is_outside(p, r) for p, r in zip(P[:400], radii[:400])])
    near = np.abs(want[:400] - radii[:400]) < 1e-5
    assert ((got_out == want_out[:400]) | near).all()
    # kd-tree on degenerate clouds
    kind = seed % 5
    n = int(rng.integers(1, 3000))
    if kind == 0:
        pts = np.repeat(rng.uniform(-1, 1, (1, 3)), n, axis=0)                       # all points identical
    elif kind == 1:
        pts = rng.uniform(-1, 1, (1, 3)) + np.outer(rng.uniform(-1, 1, n), rng.standard_normal(3))  # collinear
    elif kind == 2:
        pts = np.repeat(rng.uniform(-1, 1, (max(1, n // 7), 3)), 7, axis=0)           # heavy duplication
    elif kind == 3:
        pts = 1.0e3 + rng.uniform(-0.5, 0.5, (n, 3))                                  # far from the origin
    else:
        pts = rng.uniform(-1, 1, (n, 3)) * np.array([1.0, 1.0, 0.0])                  # planar
    pts = pts.astype(np.float32).astype(np.float64)
    tree = KDTree(pts)
    Pq = (pts.mean(axis=0) + rng.uniform(-1.5, 1.5, (300, 3))).astype(np.float32).astype(np.float64)
    brute = ((Pq[:, None, :] - pts[None, :, :]) ** 2).sum(-1).min(1)
    got = np.array([tree.sqdist(p) for p in Pq[:100]])
    scale = max(1.0, float(np.abs(pts).max()))
    tol = 2e-5 * scale * (1 + np.sqrt(brute))  # FP32 squared distances: error grows with the coordinates' magnitude
    assert (np.abs(got - brute[:100]) <= tol[:100]).all(), np.abs(got - brute[:100]).max()
    nn = np.array([tree.query(p) for p in Pq[:50]])
    assert (np.abs(((nn - Pq[:50]) ** 2).sum(1) - brute[:50]) <= tol[:50]).all()


@pytest.mark.parametrize("seed", range(4100, 4108))
def test_random_scene_mutations_follow_the_reference(cuda_lib, seed):
    """A random sequence of translate / rotate_z on leaves and on the union, set_sdf with a new scene and
    back: after every step the CUDA answer equals the oracle's for the mutated scene."""
    from fuzz_util import random_scene

    cst, lb, ub, desc = random_case(seed)
    rng = np.random.default_rng(seed)
    scene = random_scene(rng, with_cloud=False)
    cst.set_sdf(scene)
    Q = sample_q(lb, ub, 1500, seed)
    Qd = Q.astype(np.float64)
    for step in range(6):
        op = rng.integers(0, 4)
        target = scene
        if op == 0 and hasattr(scene, "sdfs"):
            movable = [s for s in scene.sdfs if hasattr(s, "_pose")]
            if movable:
                movable[int(rng.integers(0, len(movable)))].translate(rng.uniform(-0.3, 0.3, 3))
        elif op == 1 and hasattr(scene, "sdfs"):
            movable = [s for s in scene.sdfs if hasattr(s, "_pose")]
            if movable:
                movable[int(rng.integers(0, len(movable)))].rotate_z(float(rng.uniform(-1, 1)))
        elif op == 2:
            try:
                target.rotate_z(float(rng.uniform(-1, 1)))
            except RuntimeError:
                pass  # a union holding a GroundSDF refuses, like the reference
        else:
            scene = random_scene(rng, with_cloud=bool(step % 2))
            cst.set_sdf(scene)
        o = bridge.OracleCst(cst, n_threads=4)
        got, want, margin = cst.is_valid_batch(Q), o.is_valid_batch(Qd), o.margin_batch(Qd)
        bad = [int(i) for i in np.where(got != want)[0] if margin[i] > EPS_BAND]
        assert not bad, f"{desc}, step {step} op {op}: {len(bad)} mismatches"
