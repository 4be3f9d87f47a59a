"""GPU parity at the sizes BASELINE.json states (configs[1..3]), against the reference's own compiled
C++ (oracle/_ref, prebuilt in the authoring container) — the restatement only supplies the contact-band
margin of the items that differ.  Every mismatch is looked at; the counts are printed (run with -s).

  configs[1]  fetch table, 1 M configurations (env + anchor primitives + 14 self pairs)
  configs[2]  panda ceiled dual-bars, 1 M edges, box validator 0.05, distributions A and B
  configs[3]  PR2 dual-arm vs CloudSDF: 100 k-point cloud x 1 M configurations and 1 M-point cloud x 100 k
              configurations (the reference's kd-tree does ~240 configurations/s per core at 1 M points),
              evaluate() values and KDTree::sqdist / query at both cloud sizes
"""
import numpy as np
import pytest

from oracle import bridge, ref_bridge
from parity_util import EPS_BAND, THREADS, TOL, assert_edge_parity, assert_validity_parity
from plainmp_b200 import workloads
from plainmp_b200.kdtree import KDTree
from plainmp_b200.ompl import MotionValidator, ValidatorConfig, ValidatorType

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_bridge.available(), reason="prebuilt oracle/_ref/libplainmp_ref.so did not travel to this box")]


def _configs_vs_reference(cst, Q, what):
    got = cst.is_valid_batch(Q)
    want = ref_bridge.is_valid_batch_threads(cst, Q, THREADS)
    o = bridge.OracleCst(cst, n_threads=THREADS)
    assert_validity_parity(got, want, lambda idx: o.margin_batch(Q[idx].astype(np.float64)), what)
    return got, want


def test_fetch_table_1m_configs(cuda_lib):
    spec, cst = workloads.make_fetch()
    Q = workloads.random_configs(spec, 1_000_000, 2)
    got, want = _configs_vs_reference(cst, Q, "fetch table, 1M configs")
    assert 0.3 < want.mean() < 0.6
    # values on a 100 k slice (the reference's evaluate is ~20x slower than its is_valid)
    k = 100_000
    r = ref_bridge.RefCst(cst)
    err = np.abs(cst.evaluate_batch(Q[:k]) - r.evaluate_batch(Q[:k].astype(np.float64)))
    print(f"[parity] fetch evaluate, {k} configs: max |diff| {err.max():.2e}")
    assert err.max() < TOL


@pytest.mark.parametrize("dist", ["A", "B"])
def test_panda_ceiled_1m_edges(cuda_lib, dist):
    spec, cst = workloads.make_panda(ceiled=True)
    n = 1_000_000
    q0, q1 = workloads.random_edges(spec, n, 3, dist == "B", (lambda q: cst.is_valid_batch(q)) if dist == "B" else None)
    w = np.full(7, 0.05)
    mv = MotionValidator(cst, ValidatorConfig(ValidatorType.BOX, 0.0, list(w)))
    got, n_checked, calls = mv.check_motion_batch(q0, q1, with_calls=True)
    want = assert_edge_parity(cst, 0, w, 0.0, q0, q1, got, n_checked, calls, f"panda ceiled, 1M edges ({dist})")
    assert (want.mean() < 0.1) if dist == "A" else (want.mean() > 0.6)
    # the host-buffer entry points (FP32 and FP64 endpoints) return the same words
    import ctypes as C

    for f64 in (False, True):
        a = np.ascontiguousarray(q0, dtype=np.float64 if f64 else np.float32)
        b = np.ascontiguousarray(q1, dtype=np.float64 if f64 else np.float32)
        words = np.zeros((n + 31) // 32, dtype=np.uint32)
        nv, nc = C.c_uint64(0), C.c_uint64(0)
        fn = cuda_lib.pmp_check_edges_host_f64 if f64 else cuda_lib.pmp_check_edges_host
        assert fn(cst._robot().handle, cst._scene().handle, C.byref(mv._v), a.ctypes.data, b.ctypes.data, n, words.ctypes.data, C.byref(nv), C.byref(nc)) == 0
        assert (np.unpackbits(words.view(np.uint8), bitorder="little")[:n].astype(bool) == got).all()
        assert nv.value == int(got.sum()) and nc.value == n_checked


@pytest.mark.parametrize("n_points,n_configs", [(100_000, 1_000_000), (1_000_000, 100_000)])
def test_pr2_vs_cloud(cuda_lib, n_points, n_configs):
    spec, cst = workloads.make_pr2_cloud(n_points)
    Q = workloads.random_configs(spec, n_configs, 4)
    got, want = _configs_vs_reference(cst, Q, f"PR2 vs {n_points}-point cloud, {n_configs} configs")
    assert 0.2 < want.mean() < 0.6
    # evaluate(): min over 152 env spheres of (nearest-point distance - rho - r), cutoff 0.1
    k = 20_000 if n_points > 200_000 else 100_000
    Qd = Q[:k].astype(np.float64)
    refs = [ref_bridge.RefCst(cst) for _ in range(THREADS)]
    import threading

    want_d = np.zeros((k, 2))
    bounds = np.linspace(0, k, THREADS + 1).astype(int)
    ths = [threading.Thread(target=lambda t=t: want_d.__setitem__(slice(bounds[t], bounds[t + 1]), refs[t].evaluate_batch(Qd[bounds[t]:bounds[t + 1]]))) for t in range(THREADS)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    err = np.abs(cst.evaluate_batch(Q[:k]) - want_d)
    print(f"[parity] PR2 vs {n_points}-point cloud evaluate, {k} configs: max |diff| {np.nanmax(err):.2e}")
    assert np.nanmax(err) < TOL


@pytest.mark.parametrize("n_points", [100_000, 1_000_000])
def test_kdtree_nearest_at_baseline_cloud_sizes(cuda_lib, n_points):
    """KDTree::sqdist / query (collision/kdtree.cpp:67-96) at BASELINE's cloud sizes: sphere centres of PR2
    configurations (the queries the constraint makes) and points near the surface."""
    pts = workloads.table_cloud(n_points)
    tree, rtree = KDTree(pts), ref_bridge.RefKDTree(pts)
    rng = np.random.default_rng(8)
    spec, cst = workloads.make_pr2_cloud(100)
    centres = cst.get_all_spheres_batch(workloads.random_configs(spec, 1200, 4), accurate=True)[:, :, :3].reshape(-1, 3)
    near = pts[rng.integers(0, n_points, 50_000)] + rng.standard_normal((50_000, 3)) * 0.01
    q = np.concatenate([centres, near]).astype(np.float32).astype(np.float64)
    want = rtree.sqdist_batch(q)
    got = tree.sqdist_batch(q)
    # squared distances of FP32 coordinates: relative 1e-5 / absolute 1e-7 m^2 (sqrt: well inside 1e-5 m)
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-7)
    nn = tree.query_batch(q)
    np.testing.assert_allclose(((nn - q) ** 2).sum(axis=1), want, rtol=1e-5, atol=1e-7)
    err = np.abs(np.sqrt(got) - np.sqrt(want)).max()
    print(f"[parity] kd-tree {n_points} points, {len(q)} queries: max |dist diff| {err:.2e} m")
    assert err < TOL
