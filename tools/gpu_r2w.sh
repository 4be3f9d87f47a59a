#!/bin/bash
# r2w: cloud narrow phase as dense (configuration, sphere) units per group + lockstep tree walks (v3_cloud_run):
# parity tests that touch clouds, then the PR2-vs-cloud rates against the previous build
tag=${1:-r2w}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "cloud" > gpurun_out/${tag}_pytest_cloud.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_pytest_cloud.log
tail -5 gpurun_out/${tag}_pytest_cloud.log
rm -f gpurun_out/${tag}_ab.txt
for rep in 1 2; do
  for lib in libplainmp_b200_prev.so libplainmp_b200.so; do
    echo "== $lib (pass $rep)" >> gpurun_out/${tag}_ab.txt
    PMP_B200_LIB=$PWD/plainmp_b200/$lib timeout 600 python tools/time_workloads.py pr2cloud100k pr2cloud panda fetch >> gpurun_out/${tag}_ab.txt 2>&1
  done
done
cat gpurun_out/${tag}_ab.txt
