mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2o_pytest.log
