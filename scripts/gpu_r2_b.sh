mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "noise" > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2b_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline"
for nk in 1 7 8 9; do timeout 300 $B --noise-kernel $nk | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('NOISE', $nk, d['kernel_ms'], round(d['roofline_noise']['frac'],4), {k:round(v,3) for k,v in d['noise_floor_sweep'].items() if 'ms' in k})"; done
