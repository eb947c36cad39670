mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q -m gpu -k "sweep or calls or prescreen or critical or config3" > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2g_pytest.log
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-config-legs --no-e2e-text"
timeout 300 $B | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('SWEEP', {k:round(v,3) for k,v in d['noise_floor_sweep'].items() if 'ms' in k}, d['noise_floor_sweep']['calls_per_value_rank0'])"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'call_' -c 12 --csv --log-file gpurun_out/r2g_launches.csv $B > gpurun_out/r2g_ncu.log 2>&1
python scripts/ncu_summary.py --launches gpurun_out/r2g_launches.csv | tail -5
