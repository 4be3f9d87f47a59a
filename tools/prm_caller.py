#!/usr/bin/env python3
"""A measured CALLER for the validator session: a lazy-PRM style roadmap query.

The reference's planners call isValid / checkMotion one state and one edge at a time on one CPU core
(ompl_thin_wrap.hpp:80-125).  A sampling-based roadmap exposes the batch the GPU path wants without
changing the algorithm: sample N states, keep the valid ones, connect every node to its k nearest
neighbours, validate ALL candidate edges, search the graph.  This script runs that query twice on the same
samples — node and edge validation through the session (pmp_session_is_valid / pmp_session_check_motion, FP64
host states in, bytes out) and through the compiled reference's own isValid / checkMotion in a sequential
C++ loop on one core — and reports the time of every stage, whether the two roadmaps are identical, and the
shortest path both find.  No planner is built into the product; this is a caller-side measurement.

  python tools/prm_caller.py [n_samples=20000] [k=12]        # on a GPU box; prints one JSON line
"""
import json
import sys
import time
from pathlib import Path

import numpy as np
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import dijkstra
from scipy.spatial import cKDTree

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import bridge, ref_bridge  # noqa: E402  (checker / CPU arm only)
from plainmp_b200 import workloads  # noqa: E402
from plainmp_b200.ompl import ValidatorConfig, ValidatorSession, ValidatorType  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    spec, cst = workloads.make_panda(ceiled=True)
    widths = [0.05] * 7
    lb, ub = spec.angle_bounds()
    rng = np.random.default_rng(7)
    Q = rng.uniform(lb, ub, size=(n, 7))  # FP64 states, as the reference's callers hold them
    ses = ValidatorSession(cst, ValidatorConfig(ValidatorType.BOX, 0.0, widths), capacity=1 << 16)
    ses.is_valid(Q[:64])  # warm-up (module load, first launch)
    ses.check_motion(Q[:64], Q[64:128])

    out = {"workload": "panda ceiled dual-bars, box validator 0.05, lazy-PRM roadmap query", "n_samples": n, "k": k}
    # ---- GPU arm ----
    t0 = time.perf_counter()
    ok = ses.is_valid(Q)
    t1 = time.perf_counter()
    nodes = Q[ok]
    tree = cKDTree(nodes)
    _, nbr = tree.query(nodes, k=k + 1)
    src = np.repeat(np.arange(len(nodes)), k)
    dst = nbr[:, 1:].reshape(-1)
    keep = src < dst  # undirected: each pair once
    src, dst = src[keep], dst[keep]
    t2 = time.perf_counter()
    ses.reset_count()
    e_ok = ses.check_motion(nodes[src], nodes[dst])
    t3 = time.perf_counter()
    gpu_calls = ses.is_valid_call_count
    out["gpu"] = {"validate_nodes_s": t1 - t0, "knn_s": t2 - t1, "validate_edges_s": t3 - t2, "n_nodes": int(len(nodes)), "n_edges": int(len(src)),
                  "edges_valid": int(e_ok.sum()), "is_valid_calls": int(gpu_calls)}

    def solve(edge_ok):
        w = np.linalg.norm(nodes[src[edge_ok]] - nodes[dst[edge_ok]], axis=1)
        g = coo_matrix((w, (src[edge_ok], dst[edge_ok])), shape=(len(nodes), len(nodes)))
        a = int(np.argmin(np.linalg.norm(nodes - np.array([-1.5, 0.3, 0.0, -1.8, 0.0, 1.9, 0.8]), axis=1)))
        b = int(np.argmin(np.linalg.norm(nodes - np.array([1.5, 0.3, 0.0, -1.8, 0.0, 1.9, 0.8]), axis=1)))
        d, pred = dijkstra(g, directed=False, indices=a, return_predecessors=True)
        path = []
        if np.isfinite(d[b]):
            v = b
            while v != a:
                path.append(v)
                v = int(pred[v])
            path.append(a)
        return float(d[b]), path[::-1]

    t4 = time.perf_counter()
    cost_gpu, path_gpu = solve(e_ok)
    out["gpu"]["graph_search_s"] = time.perf_counter() - t4
    out["gpu"]["total_s"] = (t1 - t0) + (t2 - t1) + (t3 - t2) + out["gpu"]["graph_search_s"]
    out["gpu"]["path_cost"], out["gpu"]["path_nodes"] = cost_gpu, len(path_gpu)

    # ---- the reference's own code, one core, sequential (its design point) ----
    if ref_bridge.available():
        r = ref_bridge.RefCst(cst)
        t0 = time.perf_counter()
        ok_r = r.is_valid_batch(Q)
        t1 = time.perf_counter()
        e_ok_r, calls_r = r.check_motion_batch_ex(0, np.array(widths), 0.0, nodes[src], nodes[dst])
        t2 = time.perf_counter()
        cost_ref, path_ref = solve(e_ok_r)
        o = bridge.OracleCst(cst)
        node_diff = np.where(ok != ok_r)[0]
        edge_diff = np.where(e_ok != e_ok_r)[0]
        # every difference must be a contact-band item (north_star parity rule)
        nm = o.margin_batch(Q[node_diff]) if len(node_diff) else np.zeros(0)
        em = bridge.OracleValidator(0, 7, np.array(widths)).check_motion_batch_ex(o, nodes[src[edge_diff]], nodes[dst[edge_diff]], 8)[2] if len(edge_diff) else np.zeros(0)
        out["reference_cpu_1core"] = {"validate_nodes_s": t1 - t0, "validate_edges_s": t2 - t1, "is_valid_calls": int(calls_r.astype(np.int64).sum()),
                                      "total_s": (t1 - t0) + (t2 - t1) + out["gpu"]["knn_s"] + out["gpu"]["graph_search_s"], "path_cost": cost_ref, "path_nodes": len(path_ref)}
        out["agreement"] = {"nodes_differing": int(len(node_diff)), "edges_differing": int(len(edge_diff)),
                            "all_differences_inside_the_1e-5_band": bool((nm <= 1e-5).all() and (em <= 1e-5).all()),
                            "same_path": path_gpu == path_ref, "is_valid_calls_equal": int(gpu_calls) == int(calls_r.astype(np.int64).sum())}
        out["speedup"] = {"validate_nodes": (t1 - t0) / out["gpu"]["validate_nodes_s"], "validate_edges": (t2 - t1) / out["gpu"]["validate_edges_s"],
                          "whole_query": out["reference_cpu_1core"]["total_s"] / out["gpu"]["total_s"]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
