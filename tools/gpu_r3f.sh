#!/bin/bash
# r3f: lockstep nearest-point searches (cloud_nn8x4) in evaluate's cloud stage and in pmp_cloud_nearest
tag=${1:-r3f}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "cloud or nearest or evaluate or jacobian or kdtree or psdf" > gpurun_out/${tag}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
echo "== prev" > gpurun_out/${tag}_eval.txt
PMP_B200_LIB=$PWD/plainmp_b200/libplainmp_b200_prev.so timeout 600 python tools/time_evaluate.py >> gpurun_out/${tag}_eval.txt 2>&1
echo "== new" >> gpurun_out/${tag}_eval.txt
timeout 600 python tools/time_evaluate.py >> gpurun_out/${tag}_eval.txt 2>&1
cat gpurun_out/${tag}_eval.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<'PY'
import json,sys
d=json.loads(open(sys.argv[1] if len(sys.argv)>1 else "gpurun_out/r3f_bench.json").read().strip().splitlines()[-1])
print("value %.3e e2e %.3e" % (d["value"], d["e2e"]["value"]))
print(json.dumps(d.get("extras",{}),indent=0)[:3000])
PY
