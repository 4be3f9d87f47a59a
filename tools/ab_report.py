#!/usr/bin/env python3
"""Print the A/B summary written by tools/gpu_ab.sh:  python tools/ab_report.py <tag> v2 v3 ..."""
import json, sys
from pathlib import Path
tag = sys.argv[1]
out = Path(__file__).resolve().parent.parent / "gpurun_out"
for f in sorted(out.glob(f"{tag}_pytest_*.log")):
    print(f.name, " | ".join(f.read_text().strip().splitlines()[-2:]))
for v in sys.argv[2:]:
    try:
        d = json.loads((out / f"{tag}_bench_{v}.json").read_text().strip().splitlines()[-1])
        e = d.get("extras", {})
        g = lambda k, kk: e.get(k, {}).get(kk, float("nan"))
        print(f"{v:5s} {d['value']:.3e} cfg/s  e2e {d['e2e']['value']:.3e}  kern {d['roofline']['kernel_ms']:.3f} ms  edgesA {g('edges_A_random_endpoints','edges_per_s'):.3e} edgesB {g('edges_B_planner_like','edges_per_s'):.3e} fetch {g('fetch_table','configs_per_s'):.3e} pr2cloud {g('pr2_cloud_1M_points','configs_per_s'):.3e}  valid {d['config']['valid_fraction']:.4f}")
    except Exception as ex:
        print(v, "ERR", ex, (out / f"{tag}_bench_{v}.err").read_text()[-600:])
