#!/bin/bash
tag=${1:-r3d}
mkdir -p gpurun_out
rm -f gpurun_out/${tag}_ab.txt
for rep in 1 2; do
  for lib in libplainmp_b200.so libplainmp_b200_va.so libplainmp_b200_vd.so libplainmp_b200_ve.so; do
    echo "== $lib auto (pass $rep)" >> gpurun_out/${tag}_ab.txt
    PMP_B200_LIB=$PWD/plainmp_b200/$lib timeout 600 python tools/time_workloads.py panda fetch edgesA edgesB >> gpurun_out/${tag}_ab.txt 2>&1
  done
done
cat gpurun_out/${tag}_ab.txt
