#!/bin/bash
# r2f: the round-2 specs / scenes and the strict Jacobian rule on hardware, then the whole GPU suite
tag=${1:-r2f}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_fuzz_gpu.py -m gpu -q -s -k "round2 or jacobian or Jacobian or random_robot" > gpurun_out/${tag}_pytest_new.log 2>&1
echo "rc=$?" >> gpurun_out/${tag}_pytest_new.log
grep -E "parity\]|passed|failed|Error|rc=" gpurun_out/${tag}_pytest_new.log | tail -60
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
