#!/bin/bash
# r3x: session latency path — completion by spinning on the result bytes in mapped memory instead of cudaStreamSynchronize,
# small is_valid batches through the flat waypoint kernel: session / constraint tests, latency table, PRM caller
tag=${1:-r3x}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "session or is_valid or adaptor or ompl or single or scene_update or composite" > gpurun_out/${tag}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
timeout 600 python tools/latency_session.py > gpurun_out/${tag}_latency.json 2> gpurun_out/${tag}_latency.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r3x_latency.json").read().strip().splitlines()[-1])
for r in d["rows"]:
    if "n" in r: print(r["n"], "check_motion %.1f us  is_valid %.1f us   (ref %.1f / %.1f)" % (r["check_motion_us_per_call"], r["is_valid_us_per_call"], r.get("reference_cpu_check_motion_us_per_call", 0), r.get("reference_cpu_is_valid_us_per_call", 0)))
    else: print(r)
PY
timeout 600 python tools/prm_caller.py 2> gpurun_out/${tag}_prm.err | tail -c 600
