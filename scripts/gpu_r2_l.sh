mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_golden.py -x -q -m gpu > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2l_pytest.log
