mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sort_calls_dev" 2>&1 | tail -2
( time timeout 900 python bench.py > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err ) 2>&1 | tail -3; echo "bench rc=$?"
tail -3 gpurun_out/r2i_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2i_bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], d['kernel_ms'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
print('sweep', {k:(round(v,3) if isinstance(v,float) else v) for k,v in d['noise_floor_sweep'].items() if 'ms' in k})
print('roofline_sweep', json.dumps(d['noise_floor_sweep']['roofline_sweep'])[:600])
print('legs', json.dumps(d['config_legs'])[:1500])
for k,v in d['e2e_text'].items():
    if isinstance(v,dict): print(k, v['shape'], 'ours', round(v['ours_wall_s'],2), 'ref', round(v.get('reference_wall_s',0),2), 'x', round(v.get('speedup_wall',0),1), v['ours']['ee_phases_s'], v['ours']['vc_phases_s'])
print('cpu', d['cpu_baseline'])
PY
