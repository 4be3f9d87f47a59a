#!/bin/bash
# r3a: in-place environment narrow phase (v3_group_run) with the moving-base code compiled out of the FIXED-base
# kernels (76 registers): 5 CTAs/SM (PMP_V3_MODE=1) vs 6 CTAs/SM with FP16 group centres (auto = MODE 2) vs previous build
tag=${1:-r3a}
mkdir -p gpurun_out
rm -f gpurun_out/${tag}_ab.txt
for rep in 1 2; do
  echo "== prev (pass $rep)" >> gpurun_out/${tag}_ab.txt
  PMP_B200_LIB=$PWD/plainmp_b200/libplainmp_b200_prev.so timeout 600 python tools/time_workloads.py >> gpurun_out/${tag}_ab.txt 2>&1
  echo "== new, auto mode (pass $rep)" >> gpurun_out/${tag}_ab.txt
  timeout 600 python tools/time_workloads.py >> gpurun_out/${tag}_ab.txt 2>&1
  echo "== new, PMP_V3_MODE=1 (pass $rep)" >> gpurun_out/${tag}_ab.txt
  PMP_V3_MODE=1 timeout 600 python tools/time_workloads.py panda panda_ceiled fetch edgesA edgesB >> gpurun_out/${tag}_ab.txt 2>&1
done
cat gpurun_out/${tag}_ab.txt
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
