#!/bin/bash
# r2q: device timeline of the chunked edge host path (PMP_TRACE_EDGES), evaluate rates after the kernel clean-up
tag=${1:-r2q}
mkdir -p gpurun_out
PMP_TRACE_EDGES=1 timeout 300 python tools/time_edges_e2e.py > gpurun_out/${tag}_trace.txt 2>&1
grep -n "edgesA\|edgesB" gpurun_out/${tag}_trace.txt
grep "pmp trace" gpurun_out/${tag}_trace.txt | tail -28
timeout 600 python tools/time_evaluate.py > gpurun_out/${tag}_eval.txt 2>&1
cat gpurun_out/${tag}_eval.txt
