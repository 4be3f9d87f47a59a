#!/usr/bin/env python3
"""End-to-end edge validation from pinned host buffers (pmp_check_edges_host) next to the device-resident
rate, panda ceiled, 0.05 box validator, 1 M edges: distributions A (random endpoints) and B (planner-like).
  [PMP_EDGE_CHUNK_FIRST=16 PMP_EDGE_CHUNK_LAST=19] python tools/time_edges_e2e.py
"""
import ctypes as C
import os
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from plainmp_b200 import _capi, tables, workloads  # noqa: E402
from plainmp_b200.ompl import MotionValidator, ValidatorConfig, ValidatorType  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
lib = _capi.load()
n = 1 << 20
spec, cst = workloads.make_panda(ceiled=True)
robot, scene = cst._robot(), cst._scene()
mv = MotionValidator(cst, ValidatorConfig(ValidatorType.BOX, 0.0, [0.05] * 7))
for tag in ("A", "B"):
    pl = tag == "B"
    q0, q1 = workloads.random_edges(spec, n, 3, pl, (lambda q: tables.check_configs(robot, scene, q)) if pl else None)
    h0, h1 = torch.from_numpy(q0).pin_memory(), torch.from_numpy(q1).pin_memory()
    a, b = h0.to(dev), h1.to(dev)
    bits = torch.empty((n + 31) // 32, dtype=torch.int32, device=dev)
    ctr = torch.zeros(2, dtype=torch.int64, device=dev)
    s = torch.cuda.current_stream().cuda_stream

    def resident():
        _capi.check(lib.pmp_check_edges(robot.handle, scene.handle, C.byref(mv._v), a.data_ptr(), b.data_ptr(), n, bits.data_ptr(), ctr.data_ptr(), ctr.data_ptr() + 8, s))

    for _ in range(2):
        resident()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        resident()
    torch.cuda.synchronize()
    t_res = (time.perf_counter() - t0) / 5
    hbits = torch.empty((n + 31) // 32, dtype=torch.int32, pin_memory=True)
    nv, nc = C.c_uint64(0), C.c_uint64(0)

    def host():
        _capi.check(lib.pmp_check_edges_host(robot.handle, scene.handle, C.byref(mv._v), h0.data_ptr(), h1.data_ptr(), n, hbits.data_ptr(), C.byref(nv), C.byref(nc)))

    for _ in range(2):
        host()
    t0 = time.perf_counter()
    for _ in range(5):
        host()
    t_host = (time.perf_counter() - t0) / 5
    assert torch.equal(hbits.to(dev), bits)
    print(f"edges{tag} chunks 2^{os.environ.get('PMP_EDGE_CHUNK_FIRST', '16')}..2^{os.environ.get('PMP_EDGE_CHUNK_LAST', '19')}: resident {n / t_res:.3e} edges/s ({t_res * 1e3:.2f} ms), "
          f"end to end {n / t_host:.3e} edges/s ({t_host * 1e3:.2f} ms)", flush=True)
