# round 2, first pass on the GPU: parity of the new kernels, then timings of the variants (run under gpurun, ONE GPU)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2a_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 $B > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2a_bench.json'))
print('DEFAULT', d['kernel_ms'], 'roof', round(d['roofline']['frac'],3), round(d['roofline_noise']['frac'],3))
print('SWEEP', {k:v for k,v in d['noise_floor_sweep'].items() if k.endswith('ms') or 'ms_' in k})
PY
for nk in 7 8 9; do timeout 300 $B --no-sweep --noise-kernel $nk | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('NOISE', $nk, d['kernel_ms'], round(d['roofline_noise']['frac'],4))"; done
for ck in 20; do timeout 300 $B --no-sweep --call-kernel $ck | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('CALL', $ck, d['kernel_ms'], round(d['roofline']['frac'],4), d['config']['calls_per_step_rank0'])"; done
