#!/usr/bin/env python3
"""bench.py — collision-checked panda configurations per second (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the hot path over one batch of synthetic configurations already resident
in HBM: `--per-gpu` panda dual-bars configurations per rank (BASELINE config 5, weak scaling),
validity bits + valid count out; for N > 1 the step also all-gathers the packed bits and
all-reduces the count over NCCL.  The batch (28 B/config) is far larger than the 126 MB L2, so
no flush is needed between iterations.  `e2e` is the same metric through the host-buffer C-ABI
entry point (pmp_check_configs_host) with pinned host inputs: H2D + kernel + D2H inside the
timed region.  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

_REAL_STDOUT = None
METRIC = "collision_checked_configs_per_sec"
UNIT = "configs/s"
# dense algorithmic work of one configuration (flop, no culling, no early exit) and bytes
F_CFG_DENSE = {"panda_dual_bars": 14.7e3, "panda_ceiled": 16.4e3, "fetch_table": 17.3e3}
# the SAME text in both arms (the driver compares config.workload of the two lines)
WORKLOAD_TEXT = {
    "panda_dual_bars": "panda_dual_bars: uniform random 7-DOF panda configs, FK + SphereCollisionCst env (box + 2 cylinders) + 11 self pairs, validity bits + count",
    "panda_ceiled": "panda_ceiled: uniform random 7-DOF panda configs, FK + SphereCollisionCst env (2 boxes + 2 cylinders) + 11 self pairs, validity bits + count",
    "fetch_table": "fetch_table: uniform random 8-DOF fetch configs, FK + SphereCollisionCst env (table box + ground + 4 anchor primitives) + 14 self pairs, validity bits + count",
}
# ncu figures are READ from the file tools/ncu_head_json.py writes out of one `ncu --set full` pass at the
# commit named inside it (never pasted): per-item warp instructions, DRAM bytes, pipe / issue / occupancy
NCU_HEAD = ROOT / "profiles" / "r02_ncu_head.json"
NCU_KEY = {"panda_dual_bars": "panda", "panda_ceiled": "panda_ceiled", "fetch_table": "fetch"}


def ncu_record(workload):
    try:
        d = json.loads(NCU_HEAD.read_text())
        rec = dict(d["workloads"][NCU_KEY[workload]])
        rec["commit"] = d.get("commit")
        rec["source"] = str(NCU_HEAD.relative_to(ROOT))
        return rec
    except Exception:
        return None


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--per-gpu", type=int, default=1 << 24, help="configurations per GPU per step")
    p.add_argument("--workload", default="panda_dual_bars", choices=["panda_dual_bars", "panda_ceiled", "fetch_table"])
    p.add_argument("--extras", type=int, default=1, help="also measure edges/s, fetch and cloud side metrics (N=1)")
    p.add_argument("--cpu-sample", type=int, default=1 << 24, help="configurations per CPU-baseline pass (about 10 s of CPU work on 16 cores)")
    p.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="weak: --per-gpu configs on every rank; strong: --total configs split over the ranks")
    p.add_argument("--host-mem", default="pinned", choices=["pinned", "wc"], help="e2e host input buffer: torch pinned memory, or write-combined pinned pages from pmp_host_alloc")
    p.add_argument("--total", type=int, default=1 << 24, help="global configurations per step with --scaling strong")
    p.add_argument("--sweep-out", default="", help="also run BASELINE config 5's sweep (strong scaling of 1M / 16M / 256M total, weak at the same per-GPU sizes) in this process and append one JSON line per point to this file")
    return p.parse_args()


def make_workload(name):
    from plainmp_b200 import workloads

    if name == "panda_dual_bars":
        return workloads.make_panda()
    if name == "panda_ceiled":
        return workloads.make_panda(ceiled=True)
    return workloads.make_fetch()


# ---- clocks ---------------------------------------------------------------------------------
class ClockSampler:
    """SM clock, power and throttle reasons sampled DURING the timed region: NVML polled every ~2 ms
    from a thread (tens of samples even for a 50 ms region); `nvidia-smi -lms` as the fallback."""

    FIELDS = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int, uuid: str | None = None):
        self.rows = []  # (sm_mhz, max_mhz, power_w, reasons-set)
        self.proc = None
        self.gpu = gpu_index
        self.uuid = uuid
        self.nvml = None
        self._stop = threading.Event()
        self.source = None

    def _nvml_open(self):
        import pynvml

        pynvml.nvmlInit()
        h = None
        if self.uuid:
            try:
                u = self.uuid if str(self.uuid).startswith("GPU-") else "GPU-" + str(self.uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(u.encode() if isinstance(u, str) else u)
            except Exception:
                h = None
        if h is None:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu])
                except Exception:
                    pass
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        self.nvml, self.h = pynvml, h
        self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

    def _poll_nvml(self):
        n = self.nvml
        bits = {
            "hw_slowdown": getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
                pw = n.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                try:
                    mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                except Exception:
                    mask = 0
                self.rows.append((sm, self.max_mhz, pw, {k for k, b in bits.items() if mask & b}))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        try:
            self._nvml_open()
            self.source = "nvml"
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            try:
                reasons = {nm for nm, v in zip(names, r[5:9]) if v.lower().startswith("active")}
                pw = float(r[3]) if r[3].replace(".", "").isdigit() else None
                self.rows.append((float(r[1]), float(r[2]), pw, reasons))
            except Exception:
                pass

    def stop(self):
        if self.nvml is None and not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi and NVML unavailable"]}
        self._stop.set()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        else:
            self.thread.join(timeout=1)
        sm = [r[0] for r in self.rows]
        reasons = set().union(*[r[3] for r in self.rows]) if self.rows else set()
        pw = [r[2] for r in self.rows if r[2] is not None]
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(max(r[1] for r in self.rows)) if self.rows else None,
            "power_w_max": max(pw) if pw else None,
            "samples": len(sm),
            "source": self.source,
            "reasons": sorted(reasons),
        }


# ---- reference arm: the reference's own CPU implementation of the path on the host cores ----------
class CpuArm:
    """kind "reference": oracle/_ref/libplainmp_ref.so — the reference's own C++ sources for the
    path, compiled from /root/reference in the authoring container (oracle/Makefile.ref; Eigen is a
    minimal API stand-in, so absolute speed may differ somewhat from an upstream build) — one private
    kin + cst per host thread over disjoint slices, since the reference is single-threaded and
    stateful.  kind "port": the FP64 restatement (oracle/plainmp_oracle.cpp), used only when the
    prebuilt reference library is not present."""

    def __init__(self, cst, threads, kind=None):
        from oracle import bridge, ref_bridge

        self.threads = threads
        self.kind = kind or ("reference" if ref_bridge.available() else "port")
        if self.kind == "reference":
            self.insts = [ref_bridge.RefCst(cst) for _ in range(threads)]
        else:
            self.o = bridge.OracleCst(cst)

    def is_valid(self, Q):
        if self.kind == "port":
            return self.o.is_valid_batch(Q, self.threads)
        out = np.zeros(len(Q), dtype=bool)
        bounds = np.linspace(0, len(Q), self.threads + 1).astype(int)

        def work(t):  # ctypes releases the GIL for the duration of the C++ loop
            out[bounds[t] : bounds[t + 1]] = self.insts[t].is_valid_batch(Q[bounds[t] : bounds[t + 1]])

        ths = [threading.Thread(target=work, args=(t,)) for t in range(self.threads)]
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        return out

    def rate(self, spec, n, seed=5):
        from plainmp_b200 import workloads

        Q = workloads.random_configs(spec, n, seed).astype(np.float64)
        self.is_valid(Q[: min(n, 20000)])  # warm
        t = time.perf_counter()
        v = self.is_valid(Q)
        dt = time.perf_counter() - t
        return n / dt, float(v.mean())


def work_model(cst, spec, n=20000, seed=5):
    """flop per configuration under the stated convention, from the oracle's instrumented
    counters on a sample — `executed` = what the reference's algorithm actually does for is_valid
    (lazy FK, group broad phase, AABB pre-tests, early exit), `dense` = FK + every placement (counted)
    + every env group and env sphere against every primitive + every listed pair and sphere pair."""
    from oracle import bridge
    from plainmp_b200 import psdf, workloads

    o = bridge.OracleCst(cst)
    Q = workloads.random_configs(spec, n, seed).astype(np.float64)
    ex, de = o.work_counts(Q), o.work_counts(Q, dense=True)
    env = [g for g, only in zip(cst._group_spheres, cst._group_only_self) if not only]
    g_env, s_env = len(env), sum(len(g) for g in env)
    cost = {psdf.T_BOX: 22, psdf.T_CYLINDER: 25, psdf.T_SPHERE: 11, psdf.T_GROUND: 4}
    prim_cost = sum(cost.get(leaf.get_type(), 0) for leaf in cst._leaves())
    p_g = len(cst._pairs)
    p_s = sum(len(cst._group_spheres[a]) * len(cst._group_spheres[b]) for a, b in cst._pairs)
    dense = de["flops"] / n + prim_cost * (g_env + s_env) + 11 * (p_g + p_s)
    return {"executed_flop_per_config": ex["flops"] / n, "dense_flop_per_config": dense,
            "counters_per_config": {k: ex[k] / n for k in bridge.OracleCst.WORK_FIELDS}, "sample": n}


def cpu_baseline_block(cst, spec, n, cores):
    arm = CpuArm(cst, cores)
    r_all, vf = arm.rate(spec, n)
    r_one, _ = CpuArm(cst, 1, arm.kind).rate(spec, min(n, 1 << 20))
    what = (
        "the reference's own C++ path compiled from its sources (oracle/_ref, Eigen API stand-in)"
        if arm.kind == "reference"
        else "FP64 restatement of the reference path (prebuilt oracle/_ref not present)"
    )
    block = {
        "value": r_all, "unit": UNIT, "cores": cores, "kind": arm.kind,
        "sample": f"{n} uniform random panda configs (valid fraction {vf:.3f}); {what}, one kin+cst per thread",
        "single_thread_value": r_one,
    }
    if arm.kind == "reference":  # the restatement beside it, for scale
        block["port_value"] = CpuArm(cst, cores, "port").rate(spec, n)[0]
    return block, arm


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    spec, cst = make_workload(args.workload)
    cores = os.cpu_count() or 1
    n = args.cpu_sample
    arm = CpuArm(cst, cores)
    for _ in range(args.warmup):
        arm.rate(spec, min(n, 1 << 18))
    t0 = time.perf_counter()
    rates = [arm.rate(spec, n)[0] for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    value = n * args.steps / sum(n / r for r in rates)
    single = CpuArm(cst, 1, arm.kind).rate(spec, min(n, 1 << 20))[0]
    note = (
        "the reference's own C++ sources for the path (kinematics, primitive_sdf, SphereCollisionCst) compiled from /root/reference into oracle/_ref (Eigen replaced by a minimal API stand-in; the upstream cmake/pybind build needs Eigen+OMPL+Boost, absent here)"
        if arm.kind == "reference"
        else "CPU restatement of the reference path (prebuilt oracle/_ref not present)"
    )
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(n / r for r in rates) / args.steps,  # timed region only (input generation excluded)
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD_TEXT[args.workload], "per_step_configs": n, "note": note + "; all host threads, one kin+cst per thread"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": arm.kind, "sample": f"{n} uniform random panda configs per step", "single_thread_value": single},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- B200 arm --------------------------------------------------------------------------------------
class ShardedStep:
    """One step of the sharded hot path on this rank: the validity kernel over the rank's resident
    configurations and, for N > 1, ONE NCCL collective that delivers every rank's packed words AND its
    64-bit valid count (the kernel writes the count straight behind the words, parallel.gather_validity_fused).
    The collective of step i runs on a side stream and overlaps the kernel of step i + 1 (two send /
    receive buffer pairs)."""

    def __init__(self, torch, dist, tables, parallel, robot, scene, q_dev, world, dev):
        self.torch, self.dist, self.tables, self.parallel = torch, dist, tables, parallel
        self.robot, self.scene, self.q, self.world = robot, scene, q_dev, world
        n = q_dev.shape[0]
        self.n, self.n_words = n, (n + 31) // 32
        self.per = parallel.fused_words(self.n_words)
        self.send = [torch.zeros(self.per, dtype=torch.int32, device=dev) for _ in range(2)]
        self.recv = [torch.empty(self.per * world, dtype=torch.int32, device=dev) for _ in range(2)] if world > 1 else None
        self.comm = torch.cuda.Stream(device=dev) if world > 1 else None
        self.kernel_done = [torch.cuda.Event() for _ in range(2)]
        self.comm_done = [torch.cuda.Event() for _ in range(2)]
        self.i = 0

    def bits(self, b):
        return self.send[b][: self.n_words]

    def count(self, b):
        return self.send[b][self.per - 2 :].view(self.torch.int64)

    def __call__(self, kev=None):
        torch = self.torch
        b = self.i & 1
        self.i += 1
        cur = torch.cuda.current_stream()
        if self.world > 1:
            cur.wait_event(self.comm_done[b])  # the gather of step i - 2 has read this send buffer
        self.count(b).zero_()
        if kev:
            kev[0].record()
        self.tables.check_configs_device(self.robot, self.scene, self.q, self.bits(b), self.count(b))
        if kev:
            kev[1].record()
        if self.world > 1:
            self.kernel_done[b].record(cur)
            with torch.cuda.stream(self.comm):
                self.comm.wait_event(self.kernel_done[b])
                self.dist.all_gather_into_tensor(self.recv[b], self.send[b])
                self.comm_done[b].record(self.comm)
        return b

    def join(self):
        if self.world > 1:
            cur = self.torch.cuda.current_stream()
            for e in self.comm_done:
                cur.wait_event(e)


def timed_steps(torch, dist, step, steps, warmup, world, dev, with_kernel_events=False):
    """W warm-up steps, then exactly K steps between a barrier + synchronise on both sides; device time
    (CUDA events on the launching stream), max over ranks.  Returns (ms per step, mean kernel ms, last buffer)."""

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    step.join()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)] if with_kernel_events else None
    e0.record()
    last = 0
    for i in range(steps):
        last = step(kev[i] if kev else None)
    step.join()
    e1.record()
    sync_all()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kev])) if kev else None
    return float(ms.item()) / steps, kernel_ms, last


def popcount_words(torch, words):
    lut = torch.tensor([bin(i).count("1") for i in range(256)], dtype=torch.int64, device=words.device)
    return int(lut[words.contiguous().view(torch.uint8).long()].sum().item())


def run_b200(args):
    import torch
    import torch.distributed as dist

    from plainmp_b200 import _capi, parallel, tables, workloads
    from plainmp_b200.ompl import MotionValidator, ValidatorConfig, ValidatorType

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    _capi.require_cuda()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from plainmp_b200.parallel import bind_to_gpu_numa_node

    numa_node = bind_to_gpu_numa_node(local_rank)  # before any pinned allocation
    if world > 1:
        # NCCL prints its version banner (NCCL_DEBUG=VERSION or WARN) on stdout from native code: point
        # fd 1 at stderr for the rest of the run and keep the real stdout for the one JSON line
        global _REAL_STDOUT
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    lib = _capi.load()

    spec, cst = make_workload(args.workload)
    robot, scene = cst._robot(), cst._scene()
    if args.scaling == "strong":
        lo, hi = parallel.shard_bounds(args.total, rank, world)
        n_local, n_global = hi - lo, args.total
    else:
        n_local = args.per_gpu
        n_global = n_local * world
    # synthetic inputs: uniform in the joint limits, generated on the host with a fixed seed per rank
    lb, ub = spec.angle_bounds()
    dof = lb.shape[0]
    rng = np.random.default_rng(5 + rank)
    if args.host_mem == "wc":
        # write-combined pinned pages (pmp_host_alloc): written once, sequentially, by the generator below
        wc_ptr = C.c_void_p()
        _capi.check(lib.pmp_host_alloc(n_local * dof * 4, 1, C.byref(wc_ptr)))
        q_host = torch.from_numpy(np.ctypeslib.as_array((C.c_float * (n_local * dof)).from_address(wc_ptr.value))).view(n_local, dof)
    else:
        q_host = torch.empty((n_local, dof), dtype=torch.float32, pin_memory=True)
    chunk = 1 << 20
    for i in range(0, n_local, chunk):
        m = min(chunk, n_local - i)
        q_host[i : i + m] = torch.from_numpy(rng.uniform(lb, ub, size=(m, dof)).astype(np.float32))
    q_dev = q_host.to(dev, non_blocking=True)
    step = ShardedStep(torch, dist, tables, parallel, robot, scene, q_dev, world, dev)

    try:
        gpu_uuid = str(torch.cuda.get_device_properties(dev).uuid)
    except Exception:
        gpu_uuid = None
    # warm-up outside the sampled region, then clocks sampled DURING the timed steps
    for _ in range(max(args.warmup, 3)):
        step()
    step.join()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank, gpu_uuid)
    if rank == 0:
        sampler.start()
    launches0 = lib.pmp_launch_count()
    ms_per_step, kernel_ms, last = timed_steps(torch, dist, step, args.steps, 0, world, dev, with_kernel_events=True)
    launches = lib.pmp_launch_count() - launches0
    count_local = int(step.count(last).item())
    valid_frac = count_local / max(n_local, 1)
    value = n_global / (ms_per_step * 1e-3)

    # ---- N > 1: the gathered words and counts are CHECKED, not just moved ----
    gather_check = None
    if world > 1:
        words_all, counts = parallel.gather_validity_fused(step.send[last], step.recv[last], n_global)
        table = step.recv[last].view(world, step.per)
        ok_counts = all(int(counts[r].item()) == popcount_words(torch, table[r, : step.n_words]) for r in range(world))
        # rank 0 recomputes the head of EVERY rank's shard (same seeded generator) on its own GPU
        head = min(chunk, n_local)
        ok_words = True
        if rank == 0:
            for r in range(world):
                qr = torch.from_numpy(np.random.default_rng(5 + r).uniform(lb, ub, size=(head, dof)).astype(np.float32)).to(dev)
                br, _ = tables.check_configs_device(robot, scene, qr)
                ok_words &= bool(torch.equal(br[: head // 32], table[r, : head // 32]))
        gather_check = {"counts_equal_popcount_of_gathered_words": ok_counts, "rank0_recomputed_head_of_every_shard": ok_words, "head_configs_per_rank": head,
                        "global_valid": int(counts.sum().item())}
        assert ok_counts and ok_words, f"gathered validity words / counts are wrong: {gather_check}"

    # ---- e2e through the host-buffer C ABI (pinned host in, bits + count out) ----
    words_host = torch.empty((n_local + 31) // 32, dtype=torch.int32, pin_memory=True)
    nv = C.c_uint64(0)
    e2e_steps = max(3, min(args.steps, 10))

    def e2e_step():
        _capi.check(
            lib.pmp_check_configs_host(robot.handle, scene.handle, q_host.data_ptr(), n_local, words_host.data_ptr(), C.byref(nv))
        )

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(2):
        e2e_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = n_global * e2e_steps / float(e2e_s.item())
    assert torch.equal(words_host.to(dev), step.bits(last)), "host-buffer path and device path disagree"
    assert nv.value == count_local
    # the bound of the e2e path: a bare pinned H2D copy of the same input bytes on this box
    scratch = torch.empty_like(q_dev)
    h2d_s = time_cuda(torch, lambda: scratch.copy_(q_host, non_blocking=True), steps=5, warmup=2)
    h2d_gbs = q_host.numel() * 4 / h2d_s / 1e9
    del scratch
    # what the reference's binding really holds: PAGEABLE FP64 numpy (bindings/constraint.cpp:82) — packed to
    # FP32 by host threads while staged (pmp_check_configs_host_f64)
    n64 = min(n_local, 1 << 22)
    q64 = q_dev[:n64].cpu().numpy().astype(np.float64)  # pageable (read back from the device copy: the host buffer may be write-combined)
    w64 = np.zeros((n64 + 31) // 32, dtype=np.uint32)
    f64_step = lambda: _capi.check(lib.pmp_check_configs_host_f64(robot.handle, scene.handle, q64.ctypes.data, n64, w64.ctypes.data, C.byref(nv)))  # noqa: E731
    f64_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(3):
        f64_step()
    f64_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(f64_s, op=dist.ReduceOp.MAX)
    e2e_f64_value = n64 * world * 3 / float(f64_s.item())
    assert (torch.from_numpy(w64.view(np.int32)).to(dev) == step.bits(last)[: (n64 + 31) // 32]).all()
    # ... and the same FP64 values in a PAGE-LOCKED array (a caller that reuses its state buffer registers it once):
    # the doubles cross PCIe as they are and are narrowed on the device
    q64p = torch.empty((n64, dof), dtype=torch.float64, pin_memory=True)
    q64p.copy_(torch.from_numpy(q64))
    w64p = np.zeros((n64 + 31) // 32, dtype=np.uint32)
    f64p_step = lambda: _capi.check(lib.pmp_check_configs_host_f64(robot.handle, scene.handle, q64p.data_ptr(), n64, w64p.ctypes.data, C.byref(nv)))  # noqa: E731
    f64p_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(3):
        f64p_step()
    f64p_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(f64p_s, op=dist.ReduceOp.MAX)
    e2e_f64_pinned_value = n64 * world * 3 / float(f64p_s.item())
    assert (w64p == w64).all()
    del q64p
    clocks = sampler.stop() if rank == 0 else None

    extras = {}
    cpu_baseline = None
    if rank == 0 and world == 1:
        tf = C.c_double(0.0)
        if lib.pmp_measure_fp32_peak(C.byref(tf)) == 0:
            extras["fp32_ffma_tflops_measured"] = tf.value
        if args.extras:
            extras.update(side_metrics(torch, dev, lib, tables, workloads, MotionValidator, ValidatorConfig, ValidatorType))
        # CPU baseline: the reference's own CPU path (oracle/_ref, else the restatement), bounded sample
        cpu_baseline, _ = cpu_baseline_block(cst, spec, args.cpu_sample, os.cpu_count() or 1)
        extras["work_model"] = work_model(cst, spec)
    if args.sweep_out:
        sweep(torch, dist, tables, parallel, robot, scene, lb, ub, world, rank, dev, args.sweep_out)
    if rank == 0:
        peaks = {}
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        except Exception:
            pass
        hbm_peak, hbm_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
        per_gpu_rate = n_local / (kernel_ms * 1e-3)
        bytes_per_cfg = dof * 4 + 1 / 8
        f_dense = F_CFG_DENSE[args.workload]
        fp32_peak = extras.get("fp32_ffma_tflops_measured") or 74.4
        ncu = ncu_record(args.workload)
        wm = extras.get("work_model") or {}
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        issue_peak = 148 * 4 * sm_mhz * 1e6
        roofline = {
            # the resource that bounds this path: FP32 SIMT instruction issue, not HBM and not tensor cores
            "bound": "fp32_simt", "unit": "TFLOP/s",
            "achieved": per_gpu_rate * f_dense / 1e12, "peak": fp32_peak, "frac": per_gpu_rate * f_dense / 1e12 / fp32_peak,
            "peak_source": "FFMA micro-benchmark on this GPU in this run (pmp_measure_fp32_peak)" if extras.get("fp32_ffma_tflops_measured") else "nominal 148 SM x 128 lanes x 2 x 1.965 GHz",
            "algorithmic_flop_per_config": f_dense,
            "note": "achieved = dense-equivalent flop (SURVEY 8d: no culling, no early exit) x configs/s of the kernel alone (CUDA events around each launch, mean of the timed steps); "
                    "the kernel legitimately skips most of the dense work, so `frac` says how the rate compares with a dense FP32 implementation at peak, `executed_frac` what the "
                    "reference's own algorithm executes, and `ncu` what the pipes and the issue slots really did",
            "executed_flop_per_config": wm.get("executed_flop_per_config"),
            "executed_frac": (per_gpu_rate * wm["executed_flop_per_config"] / 1e12 / fp32_peak) if wm.get("executed_flop_per_config") else None,
            "kernel_ms": kernel_ms, "configs_per_launch": n_local,
            # HBM, the secondary roof: algorithmic bytes vs measured traffic of the same kernel
            "hbm": {"achieved_gbs": per_gpu_rate * bytes_per_cfg / 1e9, "peak_gbs": hbm_peak, "frac": per_gpu_rate * bytes_per_cfg / 1e9 / hbm_peak, "peak_source": hbm_src,
                    "algorithmic_bytes_per_config": bytes_per_cfg, "algorithmic_gb_per_launch": bytes_per_cfg * n_local / 1e9},
            "traffic": (ncu["dram_bytes_per_item"] * n_local / 1e9) if ncu and ncu.get("dram_bytes_per_item") else None,
            "traffic_unit": "GB per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum per configuration x configs per launch)",
        }
        if ncu:
            roofline["ncu"] = {k: ncu.get(k) for k in ("commit", "source", "kernel", "registers_per_thread", "occupancy_limit_registers", "occupancy_limit_shared_mem",
                                                       "warps_active_pct", "issue_active_pct", "fma_pipe_pct", "alu_pipe_pct", "lsu_pipe_pct", "lanes_active_per_inst",
                                                       "warp_inst_per_item", "dram_bytes_per_item")}
            if ncu.get("warp_inst_per_item"):
                roofline["issue_slots"] = {"achieved_gwarp_inst_s": per_gpu_rate * ncu["warp_inst_per_item"] / 1e9, "peak_gwarp_inst_s": issue_peak / 1e9,
                                           "frac": per_gpu_rate * ncu["warp_inst_per_item"] / issue_peak,
                                           "note": "warp instructions per configuration (ncu, same kernel at the commit above) x this run's rate / (148 SMs x 4 schedulers x sampled SM clock)"}
        line = {
            "metric": METRIC,
            "value": value,
            "unit": UNIT,
            "n_gpus": world,
            "steps": args.steps,
            "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step,
            "higher_is_better": True,
            "scaling": args.scaling,
            "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic",
            "config": {
                "workload": WORKLOAD_TEXT[args.workload],
                "robot": spec.__class__.__name__, "per_gpu_configs": n_local, "global_configs": n_global,
                "l2": "inputs (28 B/config x per_gpu_configs) exceed the 126 MB L2; no flush needed" if n_local * dof * 4 > 126e6 else "inputs fit the L2 at this size (sweep point)",
                "parallelism": f"item-sharded x{world}, tables replicated" + (", ONE NCCL all_gather per step (words + 64-bit count), overlapped with the next step's kernel" if world > 1 else ""),
                "numa_node_rank0": numa_node, "valid_fraction": valid_frac,
            },
            "roofline": roofline,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n_local * dof * 4 * world, "d2h_bytes_per_step": (((n_local + 31) // 32) * 4 + 8) * world, "steps": e2e_steps,
                    "bound": "pcie_h2d", "h2d_copy_gbs_measured": h2d_gbs, "host_buffer": "torch pinned" if args.host_mem == "pinned" else "write-combined pinned (pmp_host_alloc)",
                    "frac_of_h2d_copy": (e2e_value / world) * dof * 4 / 1e9 / h2d_gbs,
                    "from_pageable_fp64": {"value": e2e_f64_value, "unit": UNIT, "configs_per_call": n64,
                                           "note": "pmp_check_configs_host_f64: pageable float64 numpy in (what the reference's binding holds), packed to FP32 by host threads into pinned staging; bound by the host's own memory bandwidth"},
                    "from_pinned_fp64": {"value": e2e_f64_pinned_value, "unit": UNIT, "configs_per_call": n64,
                                         "note": "the same entry point on a page-locked float64 array: the doubles are DMA'd as they are (56 B per configuration) and narrowed on the device"}},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if gather_check:
            line["gather_check"] = gather_check
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        if extras:
            line["extras"] = extras
        print(json.dumps(line), file=_REAL_STDOUT or sys.stdout, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def sweep(torch, dist, tables, parallel, robot, scene, lb, ub, world, rank, dev, out_path):
    """BASELINE config 5: 1 M - 256 M configurations sharded over the ranks with the NCCL gather of the
    validity words, STRONG scaling (fixed total split over N) next to WEAK scaling (fixed per-GPU size);
    inputs generated on the device (uniform in the joint limits).  One JSON line per point."""
    dof = lb.shape[0]
    lo_t = torch.tensor(lb, dtype=torch.float32, device=dev)
    span = torch.tensor(ub - lb, dtype=torch.float32, device=dev)
    points = [("strong", t) for t in (1 << 20, 1 << 22, 1 << 24, 1 << 26, 1 << 28)] + [("weak", t) for t in (1 << 20, 1 << 24)]
    for kind, size in points:
        if kind == "strong":
            a, b = parallel.shard_bounds(size, rank, world)
            n_local, n_global = b - a, size
        else:
            n_local, n_global = size, size * world
        gen = torch.Generator(device=dev)
        gen.manual_seed(1000 + rank)
        q = torch.rand((n_local, dof), generator=gen, device=dev, dtype=torch.float32) * span + lo_t
        step = ShardedStep(torch, dist, tables, parallel, robot, scene, q, world, dev)
        steps = int(max(5, min(200, 4e8 // max(n_local, 1))))
        ms, _, last = timed_steps(torch, dist, step, steps, 3, world, dev)
        if rank == 0:
            rec = {"scaling": kind, "n_gpus": world, "global_configs": n_global, "per_gpu_configs": n_local, "steps": steps, "ms_per_step": ms,
                   "configs_per_s": n_global / (ms * 1e-3), "collective": "one all_gather (words + count) per step, overlapped" if world > 1 else "none"}
            with open(out_path, "a") as f:
                f.write(json.dumps(rec) + "\n")
        del step, q
        torch.cuda.empty_cache()


def time_cuda(torch, fn, steps=5, warmup=2):
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


def side_metrics(torch, dev, lib, tables, workloads, MotionValidator, ValidatorConfig, ValidatorType):
    """Secondary BASELINE configs at N = 1 (device-resident inputs): edges/s for both endpoint
    distributions, fetch table configs/s, PR2-vs-cloud configs/s."""
    out = {}
    try:
        spec, cst = workloads.make_panda(ceiled=True)
        robot, scene = cst._robot(), cst._scene()
        n_e = 1 << 20
        valid_fn = lambda q: tables.check_configs(robot, scene, q)  # noqa: E731
        mv = MotionValidator(cst, ValidatorConfig(ValidatorType.BOX, 0.0, [0.05] * 7))
        for tag, planner_like in (("A_random_endpoints", False), ("B_planner_like", True)):
            q0, q1 = workloads.random_edges(spec, n_e, 3, planner_like, valid_fn if planner_like else None)
            a, b = torch.from_numpy(q0).to(dev), torch.from_numpy(q1).to(dev)
            bits = torch.empty((n_e + 31) // 32, dtype=torch.int32, device=dev)
            ctr = torch.zeros(2, dtype=torch.int64, device=dev)

            def run():
                ctr.zero_()
                _check(lib, lib.pmp_check_edges(robot.handle, scene.handle, C.byref(mv._v), a.data_ptr(), b.data_ptr(), n_e, bits.data_ptr(), ctr.data_ptr(), ctr.data_ptr() + 8, torch.cuda.current_stream().cuda_stream))

            dt = time_cuda(torch, run)
            c = ctr.cpu().numpy()
            out[f"edges_{tag}"] = {"edges_per_s": n_e / dt, "valid_fraction": float(c[0]) / n_e, "waypoint_checks_per_edge": float(c[1]) / n_e, "waypoint_checks_per_s": float(c[1]) / dt, "n": n_e}
            # the same edges end to end through the host-buffer entry point (pinned host in, bits out)
            h0, h1 = torch.from_numpy(q0).pin_memory(), torch.from_numpy(q1).pin_memory()
            hbits = torch.empty((n_e + 31) // 32, dtype=torch.int32, pin_memory=True)
            nv, nc = C.c_uint64(0), C.c_uint64(0)

            def run_host():
                _check(lib, lib.pmp_check_edges_host(robot.handle, scene.handle, C.byref(mv._v), h0.data_ptr(), h1.data_ptr(), n_e, hbits.data_ptr(), C.byref(nv), C.byref(nc)))

            run_host()
            t0 = time.perf_counter()
            for _ in range(3):
                run_host()
            out[f"edges_{tag}"]["e2e_edges_per_s"] = 3 * n_e / (time.perf_counter() - t0)
            assert torch.equal(hbits.to(dev), bits), "edge host path and device path disagree"
        spec, cst = workloads.make_fetch()
        robot, scene = cst._robot(), cst._scene()
        n_f = 1 << 22
        q = torch.from_numpy(workloads.random_configs(spec, n_f, 2)).to(dev)
        bits = torch.empty((n_f + 31) // 32, dtype=torch.int32, device=dev)
        cnt = torch.zeros(1, dtype=torch.int64, device=dev)
        dt = time_cuda(torch, lambda: tables.check_configs_device(robot, scene, q, bits, cnt))
        out["fetch_table"] = {"configs_per_s": n_f / dt, "n": n_f}
        spec, cst = workloads.make_pr2_cloud(1_000_000)
        robot, scene = cst._robot(), cst._scene()
        n_p = 1 << 20
        q = torch.from_numpy(workloads.random_configs(spec, n_p, 4)).to(dev)
        bits = torch.empty((n_p + 31) // 32, dtype=torch.int32, device=dev)
        cnt = torch.zeros(1, dtype=torch.int64, device=dev)
        dt = time_cuda(torch, lambda: tables.check_configs_device(robot, scene, q, bits, cnt), steps=3, warmup=1)
        out["pr2_cloud_1M_points"] = {"configs_per_s": n_p / dt, "n": n_p}
        # BASELINE config 4, the query itself: exact nearest-point distance (KDTree::sqdist) for the
        # sphere centres of PR2 configurations against the same 1 M-point cloud, and evaluate()
        n_s = 1 << 16
        centres = torch.empty((n_s, robot.n_spheres, 4), dtype=torch.float32, device=dev)
        _check(lib, lib.pmp_fk_spheres(robot.handle, q.data_ptr(), n_s, 1, centres.data_ptr(), torch.cuda.current_stream().cuda_stream))
        pts = centres[:, :, :3].reshape(-1, 3).contiguous()
        sq = torch.empty(pts.shape[0], dtype=torch.float32, device=dev)
        dt = time_cuda(torch, lambda: _check(lib, lib.pmp_cloud_nearest(scene.handle, pts.data_ptr(), pts.shape[0], sq.data_ptr(), None, torch.cuda.current_stream().cuda_stream)), steps=3, warmup=1)
        out["pr2_cloud_1M_points"]["nearest_point_queries_per_s"] = pts.shape[0] / dt
        # the queries evaluate() really makes: sphere centres within the 0.1 cutoff of the cloud's AABB
        # (far queries are the worst case of any exact NN structure and are culled before they happen)
        cl = torch.from_numpy(cst.get_sdf().points.astype(np.float32)).to(dev)
        lo, hi = cl.min(dim=0).values - 0.1, cl.max(dim=0).values + 0.1
        near = pts[((pts >= lo) & (pts <= hi)).all(dim=1)].contiguous()
        if near.shape[0] > 0:
            sqn = torch.empty(near.shape[0], dtype=torch.float32, device=dev)
            dt = time_cuda(torch, lambda: _check(lib, lib.pmp_cloud_nearest(scene.handle, near.data_ptr(), near.shape[0], sqn.data_ptr(), None, torch.cuda.current_stream().cuda_stream)), steps=3, warmup=1)
            out["pr2_cloud_1M_points"]["nearest_point_queries_per_s_near_cloud"] = near.shape[0] / dt
            out["pr2_cloud_1M_points"]["near_cloud_fraction"] = near.shape[0] / pts.shape[0]
        dist = torch.empty((n_s, 2), dtype=torch.float32, device=dev)
        dt = time_cuda(torch, lambda: _check(lib, lib.pmp_eval_configs(robot.handle, scene.handle, q.data_ptr(), n_s, dist.data_ptr(), torch.cuda.current_stream().cuda_stream)), steps=3, warmup=1)
        out["pr2_cloud_1M_points"]["evaluate_configs_per_s"] = n_s / dt
    except Exception as e:  # side metrics must never break the headline line
        out["error"] = repr(e)
    return out


def _check(lib, status):
    if status != 0:
        raise RuntimeError(lib.pmp_last_error().decode())


def main():
    args = parse_args()
    # a kernel that never returns must end this process (stack dump on stderr, non-zero exit) rather than
    # hold the GPU box until the caller's limit; PMP_BENCH_WATCHDOG_S=0 turns the watchdog off
    limit = float(os.environ.get("PMP_BENCH_WATCHDOG_S", "3600"))
    if limit > 0:
        import faulthandler

        faulthandler.dump_traceback_later(limit, exit=True, file=sys.stderr)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
    if limit > 0:
        faulthandler.cancel_dump_traceback_later()


if __name__ == "__main__":
    main()
