"""The parity rule of BASELINE.json's north_star, as assertions shared by the GPU tests.

Configurations: validity bit-exact except items whose signed distance lies within EPS_BAND of the
contact threshold (oracle-only `validity_margin`).  Edges: checkMotion is a serial walk with early exit,
so an edge may differ from the reference — in validity OR in the number of is_valid calls — only if one
of the waypoints the reference's walk visits lies in that band; every differing edge is looked at.
"""
import os

import numpy as np

from oracle import bridge, ref_bridge

EPS_BAND = 1e-5  # metres
TOL = 1e-5       # metres: distances and sphere centres, FP32 vs FP64
THREADS = os.cpu_count() or 8


def assert_validity_parity(got, want, margin_of, what):
    """margin_of(indices) -> margins of those items (computed only for the mismatching ones)."""
    bad = np.where(np.asarray(got) != np.asarray(want))[0]
    m = margin_of(bad) if len(bad) else np.zeros(0)
    out = bad[m > EPS_BAND]
    print(f"[parity] {what}: n={len(got)} valid {np.mean(want):.4f} mismatches {len(bad)} (all in band: {len(out) == 0}; max margin {m.max() if len(bad) else 0:.2e})")
    assert len(out) == 0, f"{what}: {len(out)} validity mismatches outside the {EPS_BAND} band, e.g. items {out[:5]} margins {m[m > EPS_BAND][:5]}"
    return len(bad)


def reference_edges(cst, vtype, widths, res, q0, q1):
    """(valid, calls per edge) from the compiled reference when it travelled to the box, else from the
    restatement pinned to it (tests/test_reference_pin.py)."""
    q0, q1 = np.asarray(q0, dtype=np.float64), np.asarray(q1, dtype=np.float64)
    if ref_bridge.available():
        return ref_bridge.check_motion_batch_threads(cst, int(vtype), widths, res, q0, q1, THREADS)
    o = bridge.OracleCst(cst)
    v, c, _ = bridge.OracleValidator(int(vtype), q0.shape[1], widths, res).check_motion_batch_ex(o, q0, q1, THREADS, with_margin=False)
    return v, c


def assert_edge_parity(cst, vtype, widths, res, q0, q1, got, n_checked, calls, what):
    """got / calls: validity and is_valid calls per edge from the CUDA path (calls may be None: then only
    the total n_checked is compared, over the edges whose walk stays out of the band)."""
    q0d, q1d = np.asarray(q0, dtype=np.float64), np.asarray(q1, dtype=np.float64)
    want, calls_ref = reference_edges(cst, vtype, widths, res, q0d, q1d)
    o = bridge.OracleCst(cst)
    ov = bridge.OracleValidator(int(vtype), q0d.shape[1], widths, res)
    if calls is not None:
        assert int(n_checked) == int(calls.astype(np.int64).sum()), f"{what}: n_checked {n_checked} != sum of calls_per_edge {int(calls.sum())}"
        bad = np.where((np.asarray(got) != want) | (calls != calls_ref))[0]
    else:
        bad = np.where(np.asarray(got) != want)[0]
    m = ov.check_motion_batch_ex(o, q0d[bad], q1d[bad], THREADS)[2] if len(bad) else np.zeros(0)
    out = bad[m > EPS_BAND]
    n_val = int((np.asarray(got) != want).sum())
    print(f"[parity] {what}: n={len(got)} valid {want.mean():.4f} edges differing {len(bad)} ({n_val} in validity; all in band: {len(out) == 0}); "
          f"is_valid calls: cuda {int(n_checked)} reference {int(calls_ref.astype(np.int64).sum())}")
    assert len(out) == 0, f"{what}: {len(out)} edges differ outside the {EPS_BAND} band, e.g. {out[:5]} margins {m[m > EPS_BAND][:5]}"
    if calls is None:
        # totals must agree once the in-band edges are set aside
        mm = ov.check_motion_batch_ex(o, q0d, q1d, THREADS)[2]
        clear = mm > EPS_BAND
        if clear.all():
            assert int(n_checked) == int(calls_ref.astype(np.int64).sum()), f"{what}: {n_checked} is_valid calls, the reference makes {int(calls_ref.sum())}"
    return want


JAC_TOL = 1e-4  # the reference's own Jacobian check passes at 4 decimals (tests/test_constraints.py:36-44)


def assert_jacobian_parity(o, Q, jac, want_j, what, tol=JAC_TOL, delta=EPS_BAND, samples=16, max_fraction=0.03):
    """Rows of evaluate()'s Jacobian against the reference's (o: FP64 oracle of the same constraint).

    Every row that differs from the reference's row by more than `tol` is looked at and must be explained
    by one of two things, both properties of the REFERENCE's rows rather than of the CUDA path:
      (1) the reference differentiates the SDF by a forward difference with step 1e-6
          (primitive_sphere_collision.cpp:305-312); where a sphere centre lies within millimetres of the
          nearest feature the distance field's curvature (1 / distance) makes that difference itself
          wrong by > 1e-4.  The CUDA row (analytic gradient) must then match the derivative of the
          reference's own VALUE function, taken by central differences in FP64, within `tol`;
      (2) the rows are piecewise smooth in q and jump where the arg-min sphere / primitive face changes:
          the CUDA row must match the reference's row at some point within `delta` of q (per coordinate;
          the 1e-5 band of the validity rule, here in configuration space).
      (3) where a sphere centre lies within a fraction of a millimetre of the nearest feature the derivative
          itself is ill-conditioned (curvature ~ lever^2 / distance): if the derivative of the reference's
          values varies by S within +-1e-6 of q (per coordinate — ten FP32 ulps of a radian), a row is only
          required to lie within tol + S of it.
    Unexplained rows fail; so does an implausible number of explained ones."""
    Qd = np.asarray(Q, dtype=np.float64)
    assert jac.shape == want_j.shape, what
    assert (np.isnan(jac) == np.isnan(want_j)).all(), f"{what}: NaN pattern of disabled terms differs"
    err = np.abs(np.nan_to_num(jac) - np.nan_to_num(want_j)).max(axis=2)  # [n, terms]
    bad = np.argwhere(err > tol)
    rng = np.random.default_rng(1)
    d = Qd.shape[1]
    h = 1e-6
    n_fd, n_kink, n_cond, unexplained = 0, 0, 0, []

    def central(q, t):
        P = np.concatenate([q[None] + h * np.eye(d), q[None] - h * np.eye(d)])
        v = o.evaluate_batch(P)[:, t]
        return (v[:d] - v[d:]) / (2 * h)

    for i, t in bad:
        j_cd = central(Qd[i], t)
        e_cd = float(np.abs(j_cd - jac[i, t]).max())
        if e_cd <= tol:
            n_fd += 1
            continue
        spread = max(float(np.abs(central(Qd[i] + rng.uniform(-1e-6, 1e-6, size=d), t) - j_cd).max()) for _ in range(8))
        if e_cd <= tol + spread:
            n_cond += 1
            continue
        near = Qd[i][None] + rng.uniform(-delta, delta, size=(samples, d))
        _, jn = o.evaluate_grad_batch(near)
        e = np.abs(np.nan_to_num(jn[:, t]) - np.nan_to_num(jac[i, t])[None]).max(axis=1)
        if e.min() <= tol:
            n_kink += 1
            continue
        unexplained.append((int(i), int(t), float(err[i, t]), e_cd, float(e.min())))
    print(f"[parity] {what}: {err.size} Jacobian rows, max error {err.max() if err.size else 0:.2e}, {len(bad)} differ by more than {tol}: "
          f"{n_fd} match the central-difference derivative of the reference's values, {n_cond} are ill-conditioned, {n_kink} sit on a tie / kink, {len(unexplained)} unexplained")
    assert not unexplained, f"{what}: unexplained Jacobian rows (item, term, error vs reference, vs central difference, best nearby): {unexplained[:5]}"
    assert len(bad) <= max_fraction * max(err.size, 1), f"{what}: {len(bad)} of {err.size} rows differ — too many to be special points"
    return err
