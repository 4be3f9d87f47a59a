#!/bin/bash
tag=${1:-r3c}
mkdir -p gpurun_out
rm -f gpurun_out/${tag}_ab.txt
for rep in 1 2; do
  for lib in libplainmp_b200.so libplainmp_b200_va.so libplainmp_b200_vb.so libplainmp_b200_vc.so; do
    echo "== $lib auto (pass $rep)" >> gpurun_out/${tag}_ab.txt
    PMP_B200_LIB=$PWD/plainmp_b200/$lib timeout 600 python tools/time_workloads.py panda fetch >> gpurun_out/${tag}_ab.txt 2>&1
    echo "== $lib MODE=1 (pass $rep)" >> gpurun_out/${tag}_ab.txt
    PMP_V3_MODE=1 PMP_B200_LIB=$PWD/plainmp_b200/$lib timeout 600 python tools/time_workloads.py panda fetch >> gpurun_out/${tag}_ab.txt 2>&1
  done
done
cat gpurun_out/${tag}_ab.txt
