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
