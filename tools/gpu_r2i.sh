#!/bin/bash
# r2i: cooperative rounds for long edges — A/B against the previous build, e2e chunk schedules, edge + new tests
tag=${1:-r2i}
mkdir -p gpurun_out
bash tools/gpu_ab.sh $tag plainmp_b200/libplainmp_b200_base.so plainmp_b200/libplainmp_b200.so edgesA edgesB > /dev/null 2>&1
cat gpurun_out/${tag}_ab.txt
for cfg in "16 19" "18 20" "20 20"; do
  set -- $cfg
  PMP_EDGE_CHUNK_FIRST=$1 PMP_EDGE_CHUNK_LAST=$2 timeout 300 python tools/time_edges_e2e.py >> gpurun_out/${tag}_edges_e2e.txt 2>&1
done
cat gpurun_out/${tag}_edges_e2e.txt
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
