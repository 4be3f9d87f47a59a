#!/bin/bash
tag=${1:-r3b}
mkdir -p gpurun_out
rm -f gpurun_out/${tag}_ab.txt
for rep in 1 2; do
  echo "== prev (pass $rep)" >> gpurun_out/${tag}_ab.txt
  PMP_B200_LIB=$PWD/plainmp_b200/libplainmp_b200_prev.so timeout 600 python tools/time_workloads.py panda fetch >> gpurun_out/${tag}_ab.txt 2>&1
  echo "== new, auto mode (pass $rep)" >> gpurun_out/${tag}_ab.txt
  timeout 600 python tools/time_workloads.py panda panda_ceiled fetch pr2 pr2cloud edgesA edgesB >> gpurun_out/${tag}_ab.txt 2>&1
  echo "== new, PMP_V3_MODE=1 (pass $rep)" >> gpurun_out/${tag}_ab.txt
  PMP_V3_MODE=1 timeout 600 python tools/time_workloads.py panda fetch edgesB >> gpurun_out/${tag}_ab.txt 2>&1
done
cat gpurun_out/${tag}_ab.txt
