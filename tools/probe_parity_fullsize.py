#!/usr/bin/env python3
"""Full-size parity probe (GPU box): BASELINE configs 2-4 against the compiled reference, mismatches listed
with their contact-band margin.  The pytest version is tests/test_gpu_fullsize.py; this prints more.
  python tools/probe_parity_fullsize.py [edges] [fetch] [cloud100k] [cloud1m]
"""
import os
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import bridge, ref_bridge  # noqa: E402
from plainmp_b200 import workloads  # noqa: E402
from plainmp_b200.ompl import MotionValidator, ValidatorConfig, ValidatorType  # noqa: E402

what = sys.argv[1:] or ["edges", "fetch", "cloud100k"]
T = os.cpu_count() or 8
if "edges" in what:
    spec, cst = workloads.make_panda(ceiled=True)
    w = np.full(7, 0.05)
    mv = MotionValidator(cst, ValidatorConfig(ValidatorType.BOX, 0.0, list(w)))
    o = bridge.OracleCst(cst, n_threads=T)
    ov = bridge.OracleValidator(0, 7, w)
    for dist in "AB":
        n = 1 << 20
        q0, q1 = workloads.random_edges(spec, n, 3, dist == "B", (lambda q: cst.is_valid_batch(q)) if dist == "B" else None)
        t = time.time()
        got, n_checked = mv.check_motion_batch(q0, q1)
        t_gpu = time.time() - t
        t = time.time()
        want, calls = ref_bridge.check_motion_batch_threads(cst, 0, w, 0.0, q0, q1, T)
        t_ref = time.time() - t
        bad = np.where(got != want)[0]
        _, _, m = ov.check_motion_batch_ex(o, q0[bad], q1[bad], T)
        print(f"edges {dist}: n={n} valid {want.mean():.4f} gpu {t_gpu:.2f}s ref {t_ref:.2f}s ({T} thr) mismatches {len(bad)} margins {np.sort(m)[-8:] if len(bad) else []}"
              f" n_checked gpu {n_checked} ref {int(calls.sum())}", flush=True)
for name, mk in (("fetch", workloads.make_fetch), ("cloud100k", lambda: workloads.make_pr2_cloud(100_000)), ("cloud1m", lambda: workloads.make_pr2_cloud(1_000_000))):
    if name not in what:
        continue
    spec, cst = mk()
    n = 100_000 if name == "cloud1m" else 1 << 20  # the reference's kd-tree does ~240 configs/s per core at 1 M points
    Q = workloads.random_configs(spec, n, 2 if name == "fetch" else 4)
    t = time.time()
    got = cst.is_valid_batch(Q)
    t_gpu = time.time() - t
    t = time.time()
    want = ref_bridge.is_valid_batch_threads(cst, Q, T)
    t_ref = time.time() - t
    bad = np.where(got != want)[0]
    o = bridge.OracleCst(cst, n_threads=T)
    m = o.margin_batch(Q[bad].astype(np.float64)) if len(bad) else np.zeros(0)
    print(f"{name}: n={n} valid {want.mean():.4f} gpu {t_gpu:.2f}s ref {t_ref:.2f}s mismatches {len(bad)} max margin {m.max() if len(bad) else 0:.3e}", flush=True)
