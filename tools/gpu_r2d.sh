#!/bin/bash
tag=${1:-r2d}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_fuzz_gpu.py -m gpu -x -q -k "livelock or queue_overflow_stress" > gpurun_out/${tag}_pytest_first.log 2>&1
rc=$?; echo "rc=$rc" >> gpurun_out/${tag}_pytest_first.log
if [ $rc -ne 0 ]; then echo "hang-prone tests failed (rc=$rc)"; tail -30 gpurun_out/${tag}_pytest_first.log; exit $rc; fi
bash tools/gpu_ab.sh $tag plainmp_b200/libplainmp_b200_base.so plainmp_b200/libplainmp_b200.so panda edgesA edgesB
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -3 gpurun_out/${tag}_bench.err
TAG=$tag python - <<'PY'
import json, os
d = json.loads(open(f"gpurun_out/{os.environ['TAG']}_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "f64", d["e2e"]["from_pageable_fp64"]["value"], "roofline", {k: d["roofline"][k] for k in ("bound","achieved","peak","frac","executed_frac")})
for k, v in d.get("extras", {}).items():
    if isinstance(v, dict): print(k, {a: b for a, b in v.items() if "per_s" in a})
PY
