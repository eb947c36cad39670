mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2_final_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time timeout 900 python bench.py > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err ) 2>&1 | tail -3
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_bench_ref.json 2> gpurun_out/r2_final_bench_ref.err ) 2>&1 | tail -3
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_final_bench_n1.json'))
print('value', d['value'], 'ms', d['ms_per_step'], d['kernel_ms'], 'launches', d['gpu_launches'], d['clocks'])
print('roofline', d['roofline']['frac'], d['roofline']['traffic'], d['roofline_noise']['frac'], d['roofline_noise']['traffic'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
print('sweep', {k:(round(v,3) if isinstance(v,float) else v) for k,v in d['noise_floor_sweep'].items() if 'ms' in k}, d['noise_floor_sweep']['roofline_sweep']['caller']['frac'], d['noise_floor_sweep']['roofline_sweep']['noise']['frac'], d['noise_floor_sweep']['roofline_sweep']['caller']['traffic'])
print('c3', d['config_legs']['config3_sweep_step']['ms_per_step'], d['config_legs']['config3_sweep_step']['noise_ms'], d['config_legs']['config3_sweep_step']['caller_ms'], 'c4', d['config_legs']['config4_caller']['caller_ms'])
for k,v in d['e2e_text'].items():
    if isinstance(v,dict): print(k, 'ours', round(v['ours_wall_s'],2), 'ref', round(v.get('reference_wall_s',0),2), 'x', round(v.get('speedup_wall',0),1))
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
print(open('gpurun_out/r2_final_bench_ref.json').read()[:300])
PY
