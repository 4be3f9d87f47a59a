#!/bin/bash
# r2h: edge e2e under different chunk schedules, session latency with the persistent packing pool
tag=${1:-r2h}
mkdir -p gpurun_out
for cfg in "16 19" "17 19" "18 20" "20 20" "15 18" "16 17"; do
  set -- $cfg
  PMP_EDGE_CHUNK_FIRST=$1 PMP_EDGE_CHUNK_LAST=$2 timeout 300 python tools/time_edges_e2e.py >> gpurun_out/${tag}_edges_e2e.txt 2>&1
done
cat gpurun_out/${tag}_edges_e2e.txt
timeout 600 python tools/latency_session.py > gpurun_out/${tag}_latency.json 2> gpurun_out/${tag}_latency.err
python - <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/%s_latency.json" % sys.argv[1] if len(sys.argv)>1 else "gpurun_out/r2h_latency.json").read().strip().splitlines()[-1])
    for r in d.get("rows", d if isinstance(d,list) else []):
        print({k:(round(v,1) if isinstance(v,float) else v) for k,v in r.items()})
except Exception as e:
    print("latency parse failed", e)
PY
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "round2" > gpurun_out/${tag}_pytest_new.log 2>&1
echo "rc=$?" >> gpurun_out/${tag}_pytest_new.log
grep -E "Jacobian rows|passed|failed|rc=" gpurun_out/${tag}_pytest_new.log | grep -v "^        " | cut -c1-330 | tail -12
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
