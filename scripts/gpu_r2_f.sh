mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2f_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 $B > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_bench.json'))
print('DEFAULT', d['kernel_ms'], 'roof', round(d['roofline']['frac'],3), round(d['roofline_noise']['frac'],3))
print('SWEEP', {k:round(v,3) for k,v in d['noise_floor_sweep'].items() if 'ms' in k})
PY
