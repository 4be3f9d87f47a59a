#!/bin/bash
# r3u: last evidence with the final kernels — 1-GPU sweep of BASELINE config 5 (1 M ... 256 M per launch), the PRM caller,
# a fuzz hunt with both group-centre layouts forced (FP32: PMP_V3_MODE=1, FP16: 2) incl. big robots, memcheck over the
# default fuzz seeds
tag=${1:-r3u}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --extras 0 --sweep-out gpurun_out/${tag}_sweep_n1.jsonl > gpurun_out/${tag}_bench_sweep.json 2> gpurun_out/${tag}_bench_sweep.err
cat gpurun_out/${tag}_sweep_n1.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['scaling'], d['n_gpus'], d['global_configs'], '%.3e' % d['configs_per_s'], '%.3f ms' % d['ms_per_step'])"
timeout 600 python tools/prm_caller.py > gpurun_out/${tag}_prm_caller.json 2> gpurun_out/${tag}_prm_caller.err; tail -c 900 gpurun_out/${tag}_prm_caller.json
for mode in 1 2; do
  PMP_V3_MODE=$mode PMP_FUZZ_SEEDS=80000:300 timeout 1200 python -m pytest tests/test_fuzz_gpu.py -m gpu -q -x > gpurun_out/${tag}_fuzz_mode$mode.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_fuzz_mode$mode.log
  tail -3 gpurun_out/${tag}_fuzz_mode$mode.log
done
timeout 1500 compute-sanitizer --tool memcheck --target-processes all python -m pytest tests/test_fuzz_gpu.py -m gpu -q -x -k "not stress and not livelock" > gpurun_out/${tag}_fuzz_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_fuzz_memcheck.log
grep -E "passed|failed|ERROR SUMMARY|rc=" gpurun_out/${tag}_fuzz_memcheck.log | tail -4
