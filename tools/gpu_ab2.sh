#!/bin/bash
# A/B of two builds on chosen workloads:  tools/gpu_ab2.sh <tag> <workloads...>   (libplainmp_b200_prev.so vs libplainmp_b200.so)
tag=$1; shift
mkdir -p gpurun_out; rm -f gpurun_out/${tag}_ab.txt
for rep in 1 2; do
  for lib in libplainmp_b200_prev.so libplainmp_b200.so; do
    echo "== $lib (pass $rep)" >> gpurun_out/${tag}_ab.txt
    PMP_B200_LIB=$PWD/plainmp_b200/$lib timeout 600 python tools/time_workloads.py "$@" >> gpurun_out/${tag}_ab.txt 2>&1
  done
done
cat gpurun_out/${tag}_ab.txt
