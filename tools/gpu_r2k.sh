#!/bin/bash
# r2k (1 GPU): the bench line with the current kernels, the write-combined host buffer variant, the 1-GPU sweep, the PRM caller
tag=${1:-r2k}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 300 python bench.py --steps 20 --warmup 3 --extras 0 --host-mem wc --sweep-out gpurun_out/${tag}_sweep_n1.jsonl > gpurun_out/${tag}_bench_wc.json 2> gpurun_out/${tag}_bench_wc.err
timeout 600 python tools/prm_caller.py 20000 12 > gpurun_out/${tag}_prm_caller.json 2> gpurun_out/${tag}_prm_caller.err
cat gpurun_out/${tag}_prm_caller.json; tail -3 gpurun_out/${tag}_prm_caller.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2k_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.3e ms %.3f e2e %.3e h2d %.1f GB/s f64 %.3e" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["h2d_copy_gbs_measured"], d["e2e"]["from_pageable_fp64"]["value"]))
        for k, v in d.get("extras", {}).items():
            if isinstance(v, dict): print("   ", k, {a: ("%.3e" % b) for a, b in v.items() if "per_s" in a})
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -3 gpurun_out/${tag}_bench.err
