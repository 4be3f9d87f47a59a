"""Item sharding over the GPUs of one node.

The reference has no data parallelism at all (its ``ParallelSolver`` forks N processes that race
on the same problem, src/plainmp/parallel.py:31-49).  Validity checking shards trivially:
configurations / edges are independent, robot and scene tables are replicated (KBs; a cloud grid
is tens of MB), and there is no exchange step inside the computation.  One process per GPU; NCCL
is used only to gather the bit-packed validity words and to reduce counts / minimum distances.

Every collective here takes the compute step as a callable so that the plumbing is testable on
CPU with the gloo backend (tests/test_parallel_gloo.py); the product compute is
``tables.check_configs_device``.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def bind_to_gpu_numa_node(device_index: int) -> Optional[int]:
    """Pin the calling process to the CPUs of the NUMA node its GPU hangs off, BEFORE it allocates
    pinned host buffers, so that the host side of every H2D / D2H copy is node-local.  With one
    process per GPU on a two-socket box, unbound ranks end up copying across the socket link and
    the end-to-end (host buffer) rate stops scaling past a few GPUs.  Returns the node, or None when
    it cannot be determined (single-node hosts report -1)."""
    import os

    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(device_index), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(device_index), "pci_device_id", 0)
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0"
        node = int(open(path + "/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of ceil(n / world) items rounded up to a multiple of 32, so that packed
    validity words never straddle ranks."""
    per = -(-n // world)
    per = -(-per // 32) * 32
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def shard_words(n: int, world: int) -> int:
    per = -(-n // world)
    return -(-per // 32)


def gather_validity(
    local_words: torch.Tensor, n: int, local_count: torch.Tensor, group: Optional[dist.ProcessGroup] = None
) -> Tuple[torch.Tensor, int]:
    """All-gather the per-rank packed words (padded to equal length) and sum the valid counts.
    Returns (words for all n items, total valid count) on every rank."""
    world = dist.get_world_size(group)
    w = shard_words(n, world)
    padded = torch.zeros(w, dtype=local_words.dtype, device=local_words.device)
    padded[: local_words.numel()] = local_words
    out = torch.empty(w * world, dtype=local_words.dtype, device=local_words.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    total = local_count.clone().to(torch.int64)
    dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    return out[: (n + 31) // 32], int(total.item())


def fused_words(n_local_words: int) -> int:
    """Length (int32 elements) of a rank's send buffer for ``gather_validity_fused``: the packed words,
    padded to an even count so that the 64-bit counter behind them is 8-byte aligned, plus that counter."""
    return (n_local_words + 1) // 2 * 2 + 2


def gather_validity_fused(
    send: torch.Tensor, out: torch.Tensor, n: int, group: Optional[dist.ProcessGroup] = None
) -> Tuple[torch.Tensor, torch.Tensor]:
    """ONE collective per step: every rank's send buffer holds its packed validity words followed by
    its 64-bit valid count (``fused_words`` int32 elements; the kernel writes the count straight behind
    the words), so a single all-gather delivers words and counts to every rank — no separate
    all-reduce.  ``out`` is the [world * len(send)] receive buffer.  Returns (views of) the words of
    all n items in rank order and the per-rank counts (int64[world]); no host synchronisation."""
    world = dist.get_world_size(group)
    dist.all_gather_into_tensor(out, send, group=group)
    per = send.numel()
    table = out.view(world, per)
    w = shard_words(n, world)
    words = table[:, :w].reshape(-1)[: (n + 31) // 32]
    counts = table[:, per - 2 :].contiguous().view(torch.int64).reshape(world)
    return words, counts


def reduce_min(local_min: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    out = local_min.clone()
    dist.all_reduce(out, op=dist.ReduceOp.MIN, group=group)
    return out


def check_configs_sharded(
    compute: Callable[[torch.Tensor], Tuple[torch.Tensor, torch.Tensor]],
    q_local: torch.Tensor,
    n_global: int,
    group: Optional[dist.ProcessGroup] = None,
) -> Tuple[torch.Tensor, int]:
    """``q_local`` is this rank's shard (rows shard_bounds(n_global, rank, world)).  ``compute``
    maps it to (packed words, valid count) — on a GPU box
    ``lambda q: tables.check_configs_device(robot, scene, q)``."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_bounds(n_global, rank, world)
    if q_local.shape[0] != hi - lo:
        raise ValueError(f"rank {rank} expects {hi - lo} rows, got {q_local.shape[0]}")
    words, count = compute(q_local)
    return gather_validity(words, n_global, count, group)
