#!/usr/bin/env python3
"""Device-resident throughput of every bench workload (CUDA events, warm): quick A/B tool.
  python tools/time_workloads.py [workload ...]   (default: all of tools/prof_configs.py's)
"""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from plainmp_b200 import _capi, tables, workloads  # noqa: E402
from plainmp_b200.ompl import MotionValidator, ValidatorConfig, ValidatorType  # noqa: E402

ALL = ["panda", "panda_ceiled", "fetch", "pr2", "pr2cloud100k", "pr2cloud", "edgesA", "edgesB"]
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
lib = _capi.load()


def timeit(fn, steps=5, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3 / steps


for wl in sys.argv[1:] or ALL:
    if wl.startswith("edges"):
        n = 1 << 20
        spec, cst = workloads.make_panda(ceiled=True)
        robot, scene = cst._robot(), cst._scene()
        planner_like = wl.endswith("B")
        q0, q1 = workloads.random_edges(spec, n, 3, planner_like, (lambda q: tables.check_configs(robot, scene, q)) if planner_like else None)
        a, b = torch.from_numpy(q0).to(dev), torch.from_numpy(q1).to(dev)
        mv = MotionValidator(cst, ValidatorConfig(ValidatorType.BOX, 0.0, [0.05] * 7))
        bits = torch.empty((n + 31) // 32, dtype=torch.int32, device=dev)
        ctr = torch.zeros(2, dtype=torch.int64, device=dev)

        def run():
            _capi.check(lib.pmp_check_edges(robot.handle, scene.handle, C.byref(mv._v), a.data_ptr(), b.data_ptr(), n, bits.data_ptr(), ctr.data_ptr(), ctr.data_ptr() + 8, torch.cuda.current_stream().cuda_stream))

        dt = timeit(run)
        ctr.zero_()
        run()
        c = ctr.cpu().numpy()
        print(f"{wl:14s} n={n:9d} {n / dt:.4e} edges/s  valid {c[0] / n:.4f}  checks/edge {c[1] / n:.2f}  checks/s {c[1] / dt:.3e}", flush=True)
    else:
        n = (1 << 20) if wl.startswith("pr2") else (1 << 22)
        mk = {"panda": workloads.make_panda, "panda_ceiled": lambda: workloads.make_panda(ceiled=True), "fetch": workloads.make_fetch,
              "pr2": workloads.make_pr2_table,
              "pr2cloud": lambda: workloads.make_pr2_cloud(1_000_000), "pr2cloud100k": lambda: workloads.make_pr2_cloud(100_000)}[wl]
        spec, cst = mk()
        robot, scene = cst._robot(), cst._scene()
        q = torch.from_numpy(workloads.random_configs(spec, n, 5)).to(dev)
        bits = torch.empty((n + 31) // 32, dtype=torch.int32, device=dev)
        cnt = torch.zeros(1, dtype=torch.int64, device=dev)
        dt = timeit(lambda: tables.check_configs_device(robot, scene, q, bits, cnt))
        cnt.zero_()
        tables.check_configs_device(robot, scene, q, bits, cnt)
        print(f"{wl:14s} n={n:9d} {n / dt:.4e} configs/s  valid {cnt.item() / n:.4f}", flush=True)
