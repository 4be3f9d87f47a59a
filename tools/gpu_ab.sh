#!/bin/bash
# A/B on the GPU box: parity tests + bench per kernel variant, then one ncu --set full capture.
#   tools/gpu_ab.sh <tag> "<variants for pytest>" "<variants for bench>" [ncu_variant]
tag=$1; tv=$2; bv=$3; nv=$4
mkdir -p gpurun_out
for v in $tv; do PMP_KERNEL=$v python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_$v.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_pytest_$v.log; done
for v in $bv; do PMP_KERNEL=$v python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench_$v.json 2> gpurun_out/${tag}_bench_$v.err; done
if [ -n "$nv" ]; then
  PMP_KERNEL=$nv ncu --set full --clock-control none --import-source on -k regex:k_check_configs -s 1 -c 1 -f -o gpurun_out/${tag}_prof_$nv python tools/prof_configs.py panda 4194304 3 > gpurun_out/${tag}_prof.log 2>&1
fi
