#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench (both arms), launch list, ncu --set full captures.
#   tools/gpu_round.sh <tag> [ncu workloads...]
tag=$1; shift
mkdir -p gpurun_out
# the two hang-prone cases first and under a short limit: a kernel that stops making progress then costs
# minutes of box time, not the whole call (a livelock found late in round 1 consumed 30 GPU-minutes)
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_fuzz_gpu.py -m gpu -x -q -k "livelock or queue_overflow_stress" > gpurun_out/${tag}_pytest_first.log 2>&1
rc=$?; echo "rc=$rc" >> gpurun_out/${tag}_pytest_first.log
if [ $rc -ne 0 ]; then echo "hang-prone tests failed (rc=$rc): stopping before the long steps"; tail -20 gpurun_out/${tag}_pytest_first.log; exit $rc; fi
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_pytest.log
timeout 900 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_smoke.log
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --extras 0 --cpu-sample 65536 > gpurun_out/${tag}_launches.log 2>&1
for wl in "$@"; do
  n=4194304; case $wl in pr2cloud*) n=1048576;; edges*) n=1048576;; esac
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_check_(configs|edges)' -s 1 -c 1 -f -o gpurun_out/${tag}_prof_$wl python tools/prof_configs.py $wl $n 3 > gpurun_out/${tag}_prof_$wl.log 2>&1
done
