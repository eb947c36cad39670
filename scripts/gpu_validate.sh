# Everything the driver runs at round end, in one go on a GPU box (under gpurun, ONE GPU):
#   pytest -m gpu, smoke(), the default bench line and the reference arm.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/validate_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/validate_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time timeout 900 python bench.py > gpurun_out/validate_bench_n1.json 2> gpurun_out/validate_bench_n1.err ) 2>&1 | tail -3
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/validate_bench_ref.json 2> gpurun_out/validate_bench_ref.err ) 2>&1 | tail -3
python - <<'PY'
import json
d = json.load(open('gpurun_out/validate_bench_n1.json'))
print('value', d['value'], 'ms', d['ms_per_step'], d['kernel_ms'], 'launches', d['gpu_launches'], d['clocks'])
print('roofline', d['roofline']['frac'], d['roofline_noise']['frac'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
s = d['noise_floor_sweep']
print('sweep', {k: round(v, 3) for k, v in s.items() if 'ms' in k}, s['roofline_sweep']['caller']['frac'], s['roofline_sweep']['noise']['frac'])
for k, v in d['config_legs'].items():
    print(k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items() if a.endswith('ms') or a == 'ms_per_step'})
for k, v in d['e2e_text'].items():
    if isinstance(v, dict):
        print(k, 'ours', [round(x, 2) for x in v['ours_wall_s_runs']], 'served', [round(x, 2) for x in v.get('ours_served', {}).get('wall_s_runs', [])], 'ref', round(v.get('reference_wall_s', 0), 2), 'x', round(v.get('speedup_wall', 0), 1), 'x served', round(v.get('speedup_wall_served', 0), 1))
PY
python - <<'PY'
import json
d = json.load(open('gpurun_out/validate_bench_n1.json'))
p = d.get('pileup', {})
print('pileup', {k: p.get(k) for k in ('identical_to_numpy_pileup', 'reads', 'error')}, p.get('pileup_kernel', {}).get('ms'), p.get('phases_s', {}).get('inflate_busy'))
PY
