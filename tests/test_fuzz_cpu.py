"""Differential test on RANDOM robots and scenes (tests/fuzz_util.py): the restated oracle against the
compiled reference (oracle/_ref).  Covers what the four stock robots do not: prismatic and continuous
joints anywhere in the tree, branching, non-controlled joints baked at an angle, added links, random
self-only flags and pairs, zero-radius groups, moving bases with random q, rotated primitives."""
import os

import numpy as np
import pytest

from oracle import bridge, ref_bridge
from fuzz_util import random_case, sample_q

# PMP_FUZZ_SEEDS=start:count widens the sweep for a one-off hunt (default: 40 fixed seeds)
_s, _c = (int(x) for x in os.environ.get("PMP_FUZZ_SEEDS", "1000:40").split(":"))
SEEDS = range(_s, _s + _c)
BIG_SEEDS = range(3000, 3000 + max(12, _c // 4))

pytestmark = pytest.mark.skipif(
    not (ref_bridge.available() or ref_bridge.can_build()),
    reason="oracle/_ref/libplainmp_ref.so is absent and /root/reference is not mounted",
)


@pytest.mark.parametrize("seed", SEEDS)
def test_oracle_equals_reference_on_random_robot(seed):
    ref_bridge.build()
    cst, lb, ub, desc = random_case(seed)
    o, r = bridge.OracleCst(cst), ref_bridge.RefCst(cst)
    assert (o.q_dim, o.n_spheres) == (r.q_dim, r.n_spheres) == (len(lb), o.n_spheres), desc
    Q = sample_q(lb, ub, 400, seed).astype(np.float64)
    assert (o.is_valid_batch(Q) == r.is_valid_batch(Q)).all(), desc
    vo, vr = o.evaluate_batch(Q[:150]), r.evaluate_batch(Q[:150])
    assert (np.isnan(vo) == np.isnan(vr)).all(), desc
    np.testing.assert_allclose(vo, vr, atol=1e-12, rtol=0, err_msg=desc)
    for accurate in (False, True):
        np.testing.assert_allclose(o.spheres_batch(Q[:8], accurate), r.spheres_batch(Q[:8], accurate), atol=1e-12, rtol=0, err_msg=desc)
    # the C++ flatten of the reference objects describes the same robot as the host mirror
    flat = r.flatten()
    assert flat["dof"] == len(cst._control_joint_ids) and len(flat["groups"]) == len(cst._group_links), desc


@pytest.mark.parametrize("seed", BIG_SEEDS)
def test_oracle_equals_reference_on_big_random_robot(seed):
    ref_bridge.build()
    cst, lb, ub, desc = random_case(seed, big=True)
    o, r = bridge.OracleCst(cst), ref_bridge.RefCst(cst)
    Q = sample_q(lb, ub, 200, seed).astype(np.float64)
    assert (o.is_valid_batch(Q) == r.is_valid_batch(Q)).all(), desc
    np.testing.assert_allclose(o.evaluate_batch(Q[:60]), r.evaluate_batch(Q[:60]), atol=1e-12, rtol=0, err_msg=desc)


@pytest.mark.parametrize("seed", range(6000, 6020))
def test_validators_oracle_equals_reference_on_random_robot(seed):
    """Both motion validators on random robots: edge validity AND the number of is_valid calls (the
    waypoint order with its early exit) of the restatement equal the compiled reference's."""
    ref_bridge.build()
    cst, lb, ub, desc = random_case(seed)
    dim = len(lb)
    if dim == 0:
        pytest.skip("no degrees of freedom")
    o, r = bridge.OracleCst(cst), ref_bridge.RefCst(cst)
    rng = np.random.default_rng(seed)
    widths = rng.uniform(0.02, 0.3, dim)
    q0 = sample_q(lb, ub, 300, seed + 1).astype(np.float64)
    q1 = np.clip(q0 + rng.uniform(-0.8, 0.8, q0.shape) * rng.choice([0.05, 1.0, 3.0]), lb, ub)
    for vtype, res in ((0, 0.0), (1, float(rng.uniform(0.02, 0.3)))):
        want, n_ref = r.check_motion_batch(vtype, widths, res, q0, q1)
        got, n_orc = bridge.OracleValidator(vtype, dim, widths, res).check_motion_batch(o, q0, q1)
        assert (got == want).all() and n_orc == n_ref, desc
