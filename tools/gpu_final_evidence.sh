#!/bin/bash
# r4p: the round's final evidence pass on one GPU — ncu --set full of every workload's dominant kernel (CSV
# exports), the ncu launch list of the bench command, the bench itself and the reference arm, GPU tests, smoke.
tag=${1:-r4p}
mkdir -p gpurun_out
SKIP_TIMES=1 timeout 2400 bash tools/gpu_r2a.sh $tag panda panda_ceiled fetch pr2 pr2cloud pr2cloud100k edgesA edgesB eval_panda eval_pr2cloud > gpurun_out/${tag}_prof.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --extras 0 > gpurun_out/${tag}_launches.log 2>&1
timeout 600 python tools/time_workloads.py > gpurun_out/${tag}_times.txt 2>&1
timeout 600 python tools/time_evaluate.py > gpurun_out/${tag}_eval.txt 2>&1
timeout 600 python tools/latency_session.py > gpurun_out/${tag}_latency.json 2> gpurun_out/${tag}_latency.err
timeout 900 python tools/probe_parity_fullsize.py edges fetch cloud100k cloud1m > gpurun_out/${tag}_probe.txt 2>&1
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${tag}_smoke.log
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_pytest.log
cat gpurun_out/${tag}_times.txt gpurun_out/${tag}_eval.txt; tail -3 gpurun_out/${tag}_smoke.log; tail -4 gpurun_out/${tag}_pytest.log; tail -2 gpurun_out/${tag}_probe.txt
du -sh gpurun_out
