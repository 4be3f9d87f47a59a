#!/bin/bash
# r3z: evaluate's self term with the inner loop in B's link frame and a conservative prune before the square root
tag=${1:-r3z}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "evaluate or jacobian or dist or fuzz or golden or smoke" > gpurun_out/${tag}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
timeout 600 python tools/time_evaluate.py > gpurun_out/${tag}_eval.txt 2>&1
cat gpurun_out/${tag}_eval.txt
