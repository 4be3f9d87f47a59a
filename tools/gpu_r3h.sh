#!/bin/bash
# r3h: compute-sanitizer (memcheck, racecheck, synccheck) over the GPU tests that drive the in-place narrow phase
# (v3_group_run: rank table + pending list in shared memory), then a fuzz hunt over 500 random robots / scenes
tag=${1:-r3h}
mkdir -p gpurun_out
SEL="round2 or long_edges or different_radii or several or composite or edges_against or cloud_sdf or session_matches or queue_overflow or golden or more_than_30 or many_prim"
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --target-processes all python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > gpurun_out/${tag}_sanitizer_$tool.log 2>&1
  echo "rc=$?" >> gpurun_out/${tag}_sanitizer_$tool.log
  grep -E "passed|failed|ERROR SUMMARY|rc=" gpurun_out/${tag}_sanitizer_$tool.log | tail -4
done
PMP_FUZZ_SEEDS=${PMP_FUZZ_SEEDS:-70000:500} timeout 1500 python -m pytest tests/test_fuzz_gpu.py -m gpu -q -x > gpurun_out/${tag}_fuzz.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_fuzz.log
tail -4 gpurun_out/${tag}_fuzz.log
