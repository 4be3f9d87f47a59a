#!/usr/bin/env python3
"""profiles/rNN_ncu_head.json -> a markdown table (one column per workload) for reading in a browser.
  python tools/ncu_head_md.py profiles/r02_ncu_head.json > profiles/r02_ncu_head.md
"""
import json
import sys

d = json.load(open(sys.argv[1]))
cols = []
for wl, rec in d["workloads"].items():
    cols.append((wl, rec))
    for i, extra in enumerate(rec.get("more_kernels", [])):
        cols.append((f"{wl} (kernel {i + 2})", extra))
rows = ["kernel", "duration_ms", "grid", "block", "registers_per_thread", "smem_dynamic_kb", "occupancy_limit_registers", "occupancy_limit_shared_mem",
        "warps_active_pct", "issue_active_pct", "lanes_active_per_inst", "fma_pipe_pct", "alu_pipe_pct", "lsu_pipe_pct", "xu_pipe_pct", "sm_throughput_pct",
        "l2_hit_rate_pct", "dram_throughput_pct", "stall_wait", "stall_short_scoreboard", "stall_long_scoreboard", "stall_branch_resolving",
        "stall_not_selected", "stall_math_pipe_throttle", "items", "warp_inst_per_item", "dram_bytes_per_item"]
print(f"# ncu --set full, one launch per workload — commit {d.get('commit')} (tag {d.get('tag')})\n")
print(d.get("how", "") + "\n")
print("| | " + " | ".join(c[0] for c in cols) + " |")
print("|---|" + "---|" * len(cols))
for r in rows:
    vals = []
    for _, rec in cols:
        v = rec.get(r, "")
        vals.append(f"{v:.4g}" if isinstance(v, float) else str(v).replace("|", "/"))
    print(f"| {r} | " + " | ".join(vals) + " |")
