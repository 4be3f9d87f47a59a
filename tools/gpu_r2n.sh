#!/bin/bash
# r2n: radius classes in the cloud index — A/B of the cloud kernels against the previous build, evaluate rates, whole GPU suite
tag=${1:-r2n}
mkdir -p gpurun_out
bash tools/gpu_ab.sh $tag plainmp_b200/libplainmp_b200_base.so plainmp_b200/libplainmp_b200.so pr2cloud pr2cloud100k panda > /dev/null 2>&1
cat gpurun_out/${tag}_ab.txt
timeout 600 python tools/time_evaluate.py > gpurun_out/${tag}_eval.txt 2>&1
cat gpurun_out/${tag}_eval.txt
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_pytest.log
tail -6 gpurun_out/${tag}_pytest.log
