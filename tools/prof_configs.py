#!/usr/bin/env python3
"""Small driver for ncu: a few launches of the validity kernel on a resident batch.
  python tools/prof_configs.py [workload] [n] [launches]      (workload: panda | panda_ceiled | fetch | pr2cloud | pr2cloud100k | edgesA | edgesB | eval_<panda|fetch|pr2cloud|pr2cloud100k>)
"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from plainmp_b200 import _capi, tables, workloads  # noqa: E402
from plainmp_b200.ompl import MotionValidator, ValidatorConfig, ValidatorType  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "panda"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 22
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
lib = _capi.load()
if wl.startswith("edges"):
    spec, cst = workloads.make_panda(ceiled=True)
    robot, scene = cst._robot(), cst._scene()
    planner_like = wl.endswith("B")
    q0, q1 = workloads.random_edges(spec, n, 3, planner_like, (lambda q: tables.check_configs(robot, scene, q)) if planner_like else None)
    a, b = torch.from_numpy(q0).to(dev), torch.from_numpy(q1).to(dev)
    mv = MotionValidator(cst, ValidatorConfig(ValidatorType.BOX, 0.0, [0.05] * 7))
    bits = torch.empty((n + 31) // 32, dtype=torch.int32, device=dev)
    ctr = torch.zeros(2, dtype=torch.int64, device=dev)
    for _ in range(reps):
        _capi.check(lib.pmp_check_edges(robot.handle, scene.handle, C.byref(mv._v), a.data_ptr(), b.data_ptr(), n, bits.data_ptr(), ctr.data_ptr(), ctr.data_ptr() + 8, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    print(wl, n, "valid", ctr[0].item() / reps / n, "checks/edge", ctr[1].item() / reps / n)
elif wl.startswith("eval_"):
    # evaluate(): min signed distances (exact FK), e.g. eval_panda, eval_pr2cloud, evalgrad_pr2cloud100k
    base = wl.split("_", 1)[1]
    spec, cst = {"panda": workloads.make_panda, "fetch": workloads.make_fetch, "pr2cloud": lambda: workloads.make_pr2_cloud(1_000_000),
                 "pr2cloud100k": lambda: workloads.make_pr2_cloud(100_000)}[base]()
    robot, scene = cst._robot(), cst._scene()
    q = torch.from_numpy(workloads.random_configs(spec, n, 5)).to(dev)
    dist = torch.empty((n, 2), dtype=torch.float32, device=dev)
    for _ in range(reps):
        _capi.check(lib.pmp_eval_configs(robot.handle, scene.handle, q.data_ptr(), n, dist.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    print(wl, n, "mean ext", float(dist[:, 0].mean()))
else:
    spec, cst = {"panda": workloads.make_panda, "panda_ceiled": lambda: workloads.make_panda(ceiled=True), "fetch": workloads.make_fetch, "pr2": workloads.make_pr2_table,
                 "pr2cloud": lambda: workloads.make_pr2_cloud(1_000_000), "pr2cloud100k": lambda: workloads.make_pr2_cloud(100_000)}[wl]()
    robot, scene = cst._robot(), cst._scene()
    q = torch.from_numpy(workloads.random_configs(spec, n, 5)).to(dev)
    bits = torch.empty((n + 31) // 32, dtype=torch.int32, device=dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    for _ in range(reps):
        tables.check_configs_device(robot, scene, q, bits, cnt)
    torch.cuda.synchronize()
    print(wl, n, "valid fraction", cnt.item() / reps / n)
