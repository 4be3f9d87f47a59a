tag=r2c
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_fuzz_gpu.py -m gpu -x -q -k "livelock or queue_overflow_stress" > gpurun_out/${tag}_pytest_first.log 2>&1
rc=$?; echo "rc=$rc" >> gpurun_out/${tag}_pytest_first.log
if [ $rc -ne 0 ]; then echo "hang-prone tests failed (rc=$rc)"; tail -30 gpurun_out/${tag}_pytest_first.log; exit $rc; fi
bash tools/gpu_ab.sh $tag plainmp_b200/libplainmp_b200_base.so plainmp_b200/libplainmp_b200.so
echo "== new lib, PMP_CTAS6=0" >> gpurun_out/${tag}_ab.txt
PMP_CTAS6=0 timeout 600 python tools/time_workloads.py panda panda_ceiled fetch edgesA edgesB >> gpurun_out/${tag}_ab.txt 2>&1
tail -6 gpurun_out/${tag}_ab.txt
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
SKIP_TIMES=1 bash tools/gpu_r2a.sh ${tag} panda edgesB > /dev/null 2>&1
