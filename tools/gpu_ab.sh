#!/bin/bash
# Kernel A/B on the GPU box: the same Python, two builds of libplainmp_b200 (PMP_B200_LIB), device-resident rates.
#   tools/gpu_ab.sh <tag> <libA> <libB> [workloads...]
tag=$1; a=$2; b=$3; shift 3
mkdir -p gpurun_out
for rep in 1 2; do
  for lib in $a $b; do
    echo "== $lib (pass $rep)" >> gpurun_out/${tag}_ab.txt
    PMP_B200_LIB=$PWD/$lib timeout 600 python tools/time_workloads.py "$@" >> gpurun_out/${tag}_ab.txt 2>&1
  done
done
cat gpurun_out/${tag}_ab.txt
