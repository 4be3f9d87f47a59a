#!/bin/bash
# r2r: grids of the chunked edge host path sized for a few edges per lane (PMP_EDGE_EPL), chunk schedules
tag=${1:-r2r}
mkdir -p gpurun_out
for epl in 0 2 4 8; do
  for cfg in "16 19" "17 19"; do
    set -- $cfg
    echo "== epl $epl" >> gpurun_out/${tag}_edges_e2e.txt
    PMP_EDGE_EPL=$epl PMP_EDGE_CHUNK_FIRST=$1 PMP_EDGE_CHUNK_LAST=$2 timeout 300 python tools/time_edges_e2e.py >> gpurun_out/${tag}_edges_e2e.txt 2>&1
  done
done
cat gpurun_out/${tag}_edges_e2e.txt
PMP_TRACE_EDGES=1 PMP_EDGE_EPL=4 timeout 300 python tools/time_edges_e2e.py 2>&1 | grep "pmp trace" | sed -n 30,36p
