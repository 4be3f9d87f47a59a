#!/usr/bin/env python3
"""profiles/rNN_ncu_head.json from the raw pages of one round of `ncu --set full` captures (tools/gpu_r2a.sh
exports `<tag>_prof_<workload>_raw.csv` on the GPU box): one record per workload with the launch shape,
the pipe / issue / occupancy figures and the DRAM bytes, normalised per item.  bench.py reads this file
(never pasted constants) and prints the commit it was captured at next to every figure it quotes.

  python tools/ncu_head_json.py <tag> <commit> > profiles/r02_ncu_head.json
"""
import csv
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag, commit = sys.argv[1], sys.argv[2]
# items per captured launch (tools/gpu_r2a.sh / tools/prof_configs.py)
ITEMS = {"panda": 4194304, "panda_ceiled": 4194304, "fetch": 4194304, "pr2": 1048576, "pr2cloud": 1048576, "pr2cloud100k": 1048576, "edgesA": 1048576,
         "edgesB": 1048576, "eval_panda": 1048576, "eval_fetch": 1048576, "eval_pr2cloud": 65536, "eval_pr2cloud100k": 65536}
UNIT = {"edgesA": "edge", "edgesB": "edge"}
KEYS = {
    "duration_ms": "gpu__time_duration.sum",
    "grid": "launch__grid_size",
    "block": "launch__block_size",
    "registers_per_thread": "launch__registers_per_thread",
    "smem_dynamic_kb": "launch__shared_mem_per_block_dynamic",
    "occupancy_limit_registers": "launch__occupancy_limit_registers",
    "occupancy_limit_shared_mem": "launch__occupancy_limit_shared_mem",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "inst_executed": "smsp__inst_executed.sum",
    "lanes_active_per_inst": "smsp__thread_inst_executed_per_inst_executed.ratio",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "fma_pipe_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "alu_pipe_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "lsu_pipe_pct": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "xu_pipe_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm_throughput_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram_read": "dram__bytes_read.sum",
    "dram_write": "dram__bytes_write.sum",
    "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l2_hit_rate_pct": "lts__t_sector_hit_rate.pct",
    "stall_wait": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "stall_short_scoreboard": "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "stall_long_scoreboard": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "stall_branch_resolving": "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "stall_not_selected": "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "stall_math_pipe_throttle": "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e-3, "ms": 1.0, "s": 1e3, "ns": 1e-6}
out = {"commit": commit, "tag": tag, "how": "ncu --set full --clock-control none, one launch per workload (tools/gpu_r2a.sh): replayed and cold-cache, use shares and per-item counts, not absolute times",
       "workloads": {}}
for f in sorted((ROOT / "gpurun_out").glob(f"{tag}_prof_*_raw.csv")):
    wl = f.name[len(tag) + 6 : -8]
    rows = list(csv.reader(open(f)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]

    def record(vals):
        rec = {"kernel": vals[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")}
        for k, m in KEYS.items():
            if m not in hdr:
                continue
            i = hdr.index(m)
            try:
                v = float(vals[i].replace(",", ""))
            except ValueError:
                continue
            u = units[i]
            if k in ("dram_read", "dram_write"):
                rec[k + "_bytes"] = v * SCALE.get(u, 1.0)
            elif k == "duration_ms":
                rec[k] = v * SCALE.get(u, 1.0)
            else:
                rec[k] = v
        n = ITEMS.get(wl)
        if n:
            rec["items"] = n
            rec["item"] = UNIT.get(wl, "config")
            if "inst_executed" in rec:
                rec["warp_inst_per_item"] = rec["inst_executed"] / n
            if "dram_read_bytes" in rec:
                rec["dram_bytes_per_item"] = (rec["dram_read_bytes"] + rec.get("dram_write_bytes", 0.0)) / n
        return rec

    rec = record(rows[2])
    if len(rows) > 3:  # a call that launches several kernels (evaluate against a cloud: pass 1 + the cooperative cloud stage)
        rec["more_kernels"] = [record(r) for r in rows[3:] if len(r) == len(hdr)]
    out["workloads"][wl] = rec
print(json.dumps(out, indent=1))
