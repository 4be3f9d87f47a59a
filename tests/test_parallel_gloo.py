"""world_size-2 test of the sharding / gather plumbing on CPU (gloo).  The compute step is the
oracle here (tests may use it); on a GPU box it is the CUDA kernel — the plumbing is identical."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from plainmp_b200 import parallel


def test_shard_bounds_cover_and_align():
    for n in (0, 1, 31, 32, 33, 1000, 1 << 20, (1 << 20) + 7):
        for world in (1, 2, 3, 4, 8):
            spans = [parallel.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b
            for a, b in spans:
                assert a % 32 == 0 or a == n
            assert max(b - a for a, b in spans) <= parallel.shard_words(n, world) * 32


def _worker(rank, world, port, n, seed, result_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import bridge
    from plainmp_b200 import workloads

    spec, cst = workloads.make_panda()
    o = bridge.OracleCst(cst)
    Q = workloads.random_configs(spec, n, seed)
    lo, hi = parallel.shard_bounds(n, rank, world)

    def compute(q):
        v = o.is_valid_batch(q.numpy().astype(np.float64))
        words = np.zeros((len(v) + 31) // 32, dtype=np.uint32)
        idx = np.where(v)[0]
        np.bitwise_or.at(words, idx >> 5, (np.uint32(1) << (idx & 31).astype(np.uint32)))
        return torch.from_numpy(words.view(np.int32)), torch.tensor([int(v.sum())])

    words, total = parallel.check_configs_sharded(compute, torch.from_numpy(Q[lo:hi]), n)
    mn = parallel.reduce_min(torch.tensor([float(rank) - 0.5]))
    # the one-collective form: words and the 64-bit count in ONE send buffer (what bench.py's step does)
    w_local, cnt = compute(torch.from_numpy(Q[lo:hi]))
    per = parallel.fused_words(parallel.shard_words(n, world))
    send = torch.zeros(per, dtype=torch.int32)
    send[: w_local.numel()] = w_local
    send[per - 2 :] = cnt.to(torch.int64).view(torch.int32)
    recv = torch.empty(per * world, dtype=torch.int32)
    f_words, f_counts = parallel.gather_validity_fused(send, recv, n)
    if rank == 0:
        np.savez(result_path, words=words.numpy(), total=total, mn=mn.numpy(), f_words=f_words.numpy(), f_counts=f_counts.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_process(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    n, seed = 5000, 3
    out = tmp_path / "r.npz"
    mp.spawn(_worker, args=(2, port, n, seed, str(out)), nprocs=2, join=True)
    r = np.load(out)
    from oracle import bridge
    from plainmp_b200 import tables, workloads

    spec, cst = workloads.make_panda()
    want = bridge.OracleCst(cst).is_valid_batch(workloads.random_configs(spec, n, seed).astype(np.float64))
    got = tables.unpack_bits(r["words"], n)
    assert (got == want).all()
    assert int(r["total"]) == int(want.sum())
    assert float(r["mn"][0]) == -0.5
    assert (r["f_words"] == r["words"]).all() and int(r["f_counts"].sum()) == int(want.sum()) and len(r["f_counts"]) == 2
