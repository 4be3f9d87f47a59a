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
    print(json.dumps({"workload": "panda ceiled dual-bars, box validator 0.05, planner-like edges (valid start, |dq| <= 0.5)", "rows": rows}))


if __name__ == "__main__":
    main()
