#!/bin/bash
# r2s: edge host path with a dedicated copy stream and a ring of compute slots — chunk cap and edges-per-lane sweep, timeline, edge tests
tag=${1:-r2s}
mkdir -p gpurun_out
for epl in 0 4 8; do
  for cfg in "16 18" "16 17" "17 19"; do
    set -- $cfg
    echo "== epl $epl" >> gpurun_out/${tag}_edges_e2e.txt
    PMP_EDGE_EPL=$epl PMP_EDGE_CHUNK_FIRST=$1 PMP_EDGE_CHUNK_LAST=$2 timeout 300 python tools/time_edges_e2e.py >> gpurun_out/${tag}_edges_e2e.txt 2>&1
  done
done
cat gpurun_out/${tag}_edges_e2e.txt
PMP_TRACE_EDGES=1 timeout 300 python tools/time_edges_e2e.py 2>&1 | grep "pmp trace" | sed -n 36,44p
timeout 900 python -m pytest tests -m gpu -q -x -k "edge or host or session" 2>&1 | tail -3
