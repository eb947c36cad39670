mkdir -p gpurun_out
( time timeout 900 python bench.py > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err ) 2>&1 | tail -3
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_final_bench_n1.json'))
print('value', d['value'], 'ms', d['ms_per_step'], d['kernel_ms'], 'launches', d['gpu_launches'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
print('sweep', {k:(round(v,3) if isinstance(v,float) else v) for k,v in d['noise_floor_sweep'].items() if 'ms' in k})
for k,v in d['e2e_text'].items():
    if isinstance(v,dict): print(k, 'ours', [round(x,2) for x in v['ours_wall_s_runs']], 'ref', round(v.get('reference_wall_s',0),2), 'x', round(v.get('speedup_wall',0),1), {a:round(b,2) for a,b in v['ours']['ee_phases_s'].items() if b>0.05}, {a:round(b,2) for a,b in v['ours']['vc_phases_s'].items() if b>0.05})
PY
