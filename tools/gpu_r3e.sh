#!/bin/bash
# r3e: the build to commit (in-place environment narrow phase, moving-base code compiled out of the FIXED-base kernels,
# overflow primitives out of line, up to 6 CTAs/SM) against HEAD's build; then the whole GPU suite
tag=${1:-r3e}
mkdir -p gpurun_out
rm -f gpurun_out/${tag}_ab.txt
for rep in 1 2; do
  for lib in libplainmp_b200_prev.so libplainmp_b200.so; do
    echo "== $lib (pass $rep)" >> gpurun_out/${tag}_ab.txt
    PMP_B200_LIB=$PWD/plainmp_b200/$lib timeout 600 python tools/time_workloads.py >> gpurun_out/${tag}_ab.txt 2>&1
  done
done
cat gpurun_out/${tag}_ab.txt
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
