#!/bin/bash
# r3m (8-GPU box): the driver's bench command at N = 8 / 4 / 2 with the final kernels, the strong / weak sweep of
# BASELINE config 5 at N = 8, the 2-GPU NCCL correctness test
tag=${1:-r3m}
mkdir -p gpurun_out
run() { n=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 20 --warmup 3 --extras 0 "$@"
}
run 8 --sweep-out gpurun_out/${tag}_sweep_n8.jsonl > gpurun_out/${tag}_bench_n8.json 2> gpurun_out/${tag}_bench_n8.err
run 4 > gpurun_out/${tag}_bench_n4.json 2> gpurun_out/${tag}_bench_n4.err
run 2 > gpurun_out/${tag}_bench_n2.json 2> gpurun_out/${tag}_bench_n2.err
timeout 300 python bench.py --steps 20 --warmup 3 --extras 0 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
timeout 300 python -m pytest tests/test_parallel_nccl.py -m gpu -q > gpurun_out/${tag}_nccl_test.log 2>&1
TAG=$tag python - <<'PY'
import json, glob, os
tag = os.environ.get("TAG", "r3m")
for f in sorted(glob.glob(f"gpurun_out/{tag}_bench_n*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.3e ms %.3f e2e %.3e h2d %.1f GB/s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["h2d_copy_gbs_measured"]), d.get("gather_check"))
    except Exception as e:
        print(f, "unreadable", e)
for f in sorted(glob.glob(f"gpurun_out/{tag}_sweep_n*.jsonl")):
    for l in open(f):
        d = json.loads(l); print(d["scaling"], d["n_gpus"], d["global_configs"], "%.3e" % d["configs_per_s"], "%.3f ms" % d["ms_per_step"])
PY
tail -3 gpurun_out/${tag}_nccl_test.log
