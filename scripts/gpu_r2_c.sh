B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
for nk in 7 8 9 10 11; do timeout 300 $B --noise-kernel $nk | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('NOISE', $nk, round(d['kernel_ms']['noise_model'],4), round(d['roofline_noise']['frac'],4), {k:round(v,3) for k,v in d['noise_floor_sweep'].items() if 'noise_ms' in k})"; done
