#!/bin/bash
# r2u: GPU fuzz hunt (500 random robots incl. scenes with several clouds / radius classes), the FP64 entry point on
# pageable and page-locked arrays, bench line with the pinned-FP64 figure
tag=${1:-r2u}
mkdir -p gpurun_out
PMP_FUZZ_SEEDS=${PMP_FUZZ_SEEDS:-60000:500} timeout 2400 python -m pytest tests/test_fuzz_gpu.py -m gpu -q -x > gpurun_out/${tag}_fuzz.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_fuzz.log
tail -4 gpurun_out/${tag}_fuzz.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fp64 or f64 or doubles" 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 3 --extras 0 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2u_bench.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("value %.3e e2e %.3e (h2d %.1f GB/s) pageable f64 %.3e pinned f64 %.3e traffic %s" % (d["value"], e["value"], e["h2d_copy_gbs_measured"], e["from_pageable_fp64"]["value"], e["from_pinned_fp64"]["value"], d["roofline"].get("traffic")))
print(d["roofline"].get("issue_slots"))
PY
tail -3 gpurun_out/${tag}_bench.err
