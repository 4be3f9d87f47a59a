#!/usr/bin/env python3
"""Per-call latency of the validator session next to the reference's CPU checkMotion.

  python tools/latency_session.py            # on a GPU box; prints one JSON line

For batch sizes 1 .. 16384 edges (panda ceiled dual-bars, 0.05 box validator, planner-like edges):
wall-clock microseconds per pmp_session_check_motion call (host FP64 in, bytes out, everything
inside the call) and per is_valid call, and the same edges through the compiled reference
(oracle/_ref, single thread — the reference's design point) when it is present.
"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from plainmp_b200 import workloads  # noqa: E402
from plainmp_b200.ompl import ValidatorConfig, ValidatorSession, ValidatorType  # noqa: E402


def timeit(fn, reps):
    fn()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t) / reps


def main():
    spec, cst = workloads.make_panda(ceiled=True)
    widths = [0.05] * 7
    ses = ValidatorSession(cst, ValidatorConfig(ValidatorType.BOX, 0.0, widths), capacity=16384)
    q0, q1 = workloads.random_edges(spec, 16384, 3, True, lambda q: cst.is_valid_batch(q))
    q0, q1 = q0.astype(np.float64), q1.astype(np.float64)
    ref = None
    try:
        from oracle import ref_bridge

        if ref_bridge.available():
            ref = ref_bridge.RefCst(cst)
    except Exception:
        ref = None
    rows = []
    for n in (1, 8, 64, 512, 4096, 16384):
        reps = 2000 if n <= 64 else (200 if n <= 4096 else 30)
        a, b = q0[:n], q1[:n]
        t_edge = timeit(lambda: ses.check_motion(a, b), reps)
        t_state = timeit(lambda: ses.is_valid(b), reps)
        row = {"n": n, "check_motion_us_per_call": 1e6 * t_edge, "edges_per_s": n / t_edge,
               "is_valid_us_per_call": 1e6 * t_state, "states_per_s": n / t_state}
        if ref is not None:
            t_ref = timeit(lambda: ref.check_motion_batch(0, np.array(widths), 0.0, a, b), max(3, reps // 20))
            t_ref_s = timeit(lambda: ref.is_valid_batch(b), max(3, reps // 20))
            row["reference_cpu_check_motion_us_per_call"] = 1e6 * t_ref
            row["reference_cpu_is_valid_us_per_call"] = 1e6 * t_ref_s
        rows.append(row)
    # the plain per-item call of the host mirror, cst.is_valid(q) (routed through a cached session)
    t_single = timeit(lambda: cst.is_valid(q1[0]), 2000)
    rows.append({"cst_is_valid_single_call_us": 1e6 * t_single})
    # scene update path: in-place primitive update and cloud re-index, per call
    from plainmp_b200 import psdf

    spec2, cst2 = workloads.make_panda()
    cst2.is_valid_batch(q0[:32].astype(np.float32))
    leaf = cst2.get_sdf().sdfs[1]

    def move():
        leaf.translate([1e-3, 0.0, 0.0])
        cst2._scene()

    upd = {"prims_update_us": 1e6 * timeit(move, 200)}
    spec3, cst3 = workloads.make_pr2_cloud(1000)
    cst3.is_valid_batch(workloads.random_configs(spec3, 32, 1))
    for npts in (100000, 1000000):
        clouds = [psdf.CloudSDF(workloads.table_cloud(npts, s), 0.002) for s in (1, 2, 3)]
        it = iter(clouds * 4)

        def frame():
            cst3.set_sdf(next(it))
            cst3._scene()

        upd[f"cloud_reindex_{npts}_points_ms"] = 1e3 * timeit(frame, 6)
    print(json.dumps({"scene_update": upd, "workload": "panda ceiled dual-bars, box validator 0.05, planner-like edges (valid start, |dq| <= 0.5)", "rows": rows}))


if __name__ == "__main__":
    main()
