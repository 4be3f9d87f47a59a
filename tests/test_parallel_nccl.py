"""The multi-GPU path on real hardware: one process per GPU over NCCL, items sharded by
parallel.shard_bounds, packed validity bits all-gathered and the count all-reduced; rank 0 compares
with the single-GPU answer and with the oracle.  Needs >= 2 visible GPUs (skipped otherwise; run with
`gpurun --gpus 2`)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, n, seed, result_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    from plainmp_b200 import _capi, parallel, tables, workloads

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    spec, cst = workloads.make_panda()
    robot, scene = cst._robot(), cst._scene()
    Q = workloads.random_configs(spec, n, seed)
    lo, hi = parallel.shard_bounds(n, rank, world)
    q_local = torch.from_numpy(Q[lo:hi]).to(dev)
    words, total = parallel.check_configs_sharded(lambda q: tables.check_configs_device(robot, scene, q), q_local, n)
    dist_local = torch.from_numpy(tables.eval_configs(robot, scene, Q[lo:hi])).to(dev)
    mn = parallel.reduce_min(torch.nan_to_num(dist_local, nan=float("inf")).min(dim=0).values)
    if rank == 0:
        np.savez(result_path, words=words.cpu().numpy(), total=total, mn=mn.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_gather_equals_single_gpu(cuda_lib, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    n, seed = 200_003, 3  # not a multiple of 32 x world: ragged last shard
    out = tmp_path / "r.npz"
    mp.spawn(_worker, args=(2, port, n, seed, str(out)), nprocs=2, join=True)
    r = np.load(out)
    from oracle import bridge
    from plainmp_b200 import tables, workloads

    spec, cst = workloads.make_panda()
    Q = workloads.random_configs(spec, n, seed)
    single = cst.is_valid_batch(Q)
    got = tables.unpack_bits(r["words"], n)
    assert (got == single).all()
    assert int(r["total"]) == int(single.sum())
    d = cst.evaluate_batch(Q)
    np.testing.assert_allclose(r["mn"], np.nanmin(d, axis=0), atol=0)
    o = bridge.OracleCst(cst, n_threads=8)
    Qd = Q.astype(np.float64)
    bad = np.where(got != o.is_valid_batch(Qd))[0]
    m = o.margin_batch(Qd)
    assert all(m[i] < 1e-5 for i in bad)
