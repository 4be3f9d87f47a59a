#!/bin/bash
# r2e: register-bound A/B of the validity kernel, evaluate() rates with the cooperative cloud stage, ncu of the evaluate kernels
tag=${1:-r2e}
mkdir -p gpurun_out
bash tools/gpu_ab.sh $tag plainmp_b200/libplainmp_b200_base.so plainmp_b200/libplainmp_b200.so panda panda_ceiled fetch pr2 pr2cloud > /dev/null 2>&1
timeout 600 python tools/time_evaluate.py > gpurun_out/${tag}_eval.txt 2>&1
cat gpurun_out/${tag}_eval.txt
SKIP_TIMES=1 timeout 1500 bash tools/gpu_r2a.sh $tag eval_panda eval_pr2cloud pr2cloud > /dev/null 2>&1
cat gpurun_out/${tag}_ab.txt
ls gpurun_out | grep $tag
