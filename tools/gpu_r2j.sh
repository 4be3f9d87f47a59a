#!/bin/bash
# r2j: e2e edges — both builds on the same box, more chunk schedules; the long-edge GPU test
tag=${1:-r2j}
mkdir -p gpurun_out
for lib in plainmp_b200/libplainmp_b200_base.so plainmp_b200/libplainmp_b200.so; do
  for cfg in "16 19" "17 18" "18 18" "18 19" "19 19"; do
    set -- $cfg
    echo "== $lib" >> gpurun_out/${tag}_edges_e2e.txt
    PMP_B200_LIB=$PWD/$lib PMP_EDGE_CHUNK_FIRST=$1 PMP_EDGE_CHUNK_LAST=$2 timeout 300 python tools/time_edges_e2e.py >> gpurun_out/${tag}_edges_e2e.txt 2>&1
  done
done
cat gpurun_out/${tag}_edges_e2e.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "long_edges" 2>&1 | grep -E "parity\]|passed|failed" | cut -c1-300
