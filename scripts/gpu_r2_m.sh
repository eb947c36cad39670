timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "noise" 2>&1 | tail -2
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-config-legs --no-e2e-text --no-sweep"
timeout 300 $B | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('NOISE', d['kernel_ms'], round(d['roofline_noise']['frac'],4))"
