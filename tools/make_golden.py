#!/usr/bin/env python3
"""Generate the committed fixtures under tests/golden/.  Run in the authoring container.

1. sequence_table.json — produced by importing the REFERENCE's own pure-Python generator
   (/root/reference/cpp/plainmp/ompl/generate_sequence.py, find_farthest_sequence_fast) and
   calling it for n = 1..128.  This is real reference output; the oracle, the CUDA library and
   the validator tests are pinned against it.
2. *.npz — seeded vectors for the BASELINE.json configurations at sizes that keep the files
   small.  valid_bits / dist / jac / spheres_* / n_calls are REFERENCE OUTPUT: they come from
   oracle/_ref/libplainmp_ref.so, the reference's own C++ sources for the path compiled from
   /root/reference (oracle/Makefile.ref, Eigen/OMPL API stand-ins in oracle/ref_shim/).  The
   restated oracle is asserted to agree with them while they are written.  Only `margin` (distance
   of the item to its nearest decision threshold, used to apply the 1e-5 contact band) and
   `sphere_order` are oracle-side quantities the reference has no API for.
"""
import hashlib
import importlib.util
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
OUT = ROOT / "tests" / "golden"

from oracle import bridge, ref_bridge  # noqa: E402
from plainmp_b200 import workloads  # noqa: E402


def reference_sequence_table():
    path = "/root/reference/cpp/plainmp/ompl/generate_sequence.py"
    spec = importlib.util.spec_from_file_location("ref_generate_sequence", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)  # the script body is guarded by __main__
    rows = []
    for n in range(1, 129):
        r = list(mod.find_farthest_sequence_fast(n))
        rows.append(r + [0] * (128 - n))
    return rows


def pack(v: np.ndarray) -> np.ndarray:
    return np.packbits(v.astype(np.uint8), bitorder="little")


def config_vectors(name, spec, cst, n, seed, n_spheres_items=8):
    Q = workloads.random_configs(spec, n, seed)
    o = bridge.OracleCst(cst, n_threads=8)
    r = ref_bridge.RefCst(cst)
    Qd = Q.astype(np.float64)
    valid = r.is_valid_batch(Qd)
    dist, jac = r.evaluate_grad_batch(Qd)
    sph_approx = r.spheres_batch(Qd[:n_spheres_items], accurate=False)
    sph_exact = r.spheres_batch(Qd[:n_spheres_items], accurate=True)
    # the restated oracle must agree with the reference on what is about to be committed
    assert (o.is_valid_batch(Qd) == valid).all()
    np.testing.assert_allclose(o.evaluate_batch(Qd), dist, atol=1e-12, rtol=0)
    np.testing.assert_allclose(o.spheres_batch(Qd[:n_spheres_items], accurate=True), sph_exact, atol=1e-12, rtol=0)
    margin = o.margin_batch(Qd)
    np.savez_compressed(
        OUT / f"{name}.npz",
        source="reference (oracle/_ref/libplainmp_ref.so)",
        seed=seed,
        n=n,
        q_head=Q[:64],
        valid_bits=pack(valid),
        dist=dist.astype(np.float32),
        jac_head=jac[:256].astype(np.float32),
        margin=margin.astype(np.float32),
        spheres_approx=sph_approx.astype(np.float32),
        spheres_exact=sph_exact.astype(np.float32),
        sphere_order=o.sphere_order(),
    )
    print(f"{name}: n={n} valid={valid.mean():.4f} band(<1e-5)={(margin < 1e-5).sum()}")


def edge_vectors(name, spec, cst, n, seed, planner_like, widths):
    o = bridge.OracleCst(cst)
    valid_fn = (lambda q: o.is_valid_batch(q.astype(np.float64))) if planner_like else None
    q0, q1 = workloads.random_edges(spec, n, seed, planner_like, valid_fn)
    r = ref_bridge.RefCst(cst)
    valid, calls = r.check_motion_batch(0, widths, 0.0, q0.astype(np.float64), q1.astype(np.float64))
    v = bridge.OracleValidator(0, q0.shape[1], widths)
    valid_o, calls_o = v.check_motion_batch(o, q0.astype(np.float64), q1.astype(np.float64), n_threads=8)
    assert (valid_o == valid).all() and calls_o == calls
    np.savez_compressed(
        OUT / f"{name}.npz", source="reference (oracle/_ref/libplainmp_ref.so)", seed=seed, n=n,
        valid_bits=pack(valid), n_calls=calls, widths=widths,
    )
    print(f"{name}: n={n} valid={valid.mean():.4f} is_valid calls/edge={calls / n:.2f}")


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    ref_bridge.build()
    rows = reference_sequence_table()
    digest = hashlib.sha256(str([r for r in rows]).encode()).hexdigest()
    (OUT / "sequence_table.json").write_text(json.dumps({"sha256_of_str_rows": digest, "rows": rows}) + "\n")
    print("sequence table sha256", digest)

    spec, cst = workloads.make_panda()
    config_vectors("panda_dualbars_seed1", spec, cst, 10000, 1)
    spec, cst = workloads.make_fetch()
    config_vectors("fetch_table_seed2", spec, cst, 4096, 2)
    spec, cst = workloads.make_pr2_cloud(20000)
    config_vectors("pr2_cloud20k_seed4", spec, cst, 2048, 4)
    spec, cst = workloads.make_openarm()
    config_vectors("openarm_dual_seed6", spec, cst, 2048, 6)
    spec, cst = workloads.make_panda(ceiled=True)
    edge_vectors("panda_ceiled_edges_A_seed3", spec, cst, 2048, 3, False, np.ones(7) * 0.05)
    edge_vectors("panda_ceiled_edges_B_seed3", spec, cst, 2048, 3, True, np.ones(7) * 0.05)


if __name__ == "__main__":
    main()
