mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2k_topo.txt 2>&1
nproc > gpurun_out/r2k_nproc.txt; free -g >> gpurun_out/r2k_nproc.txt
timeout 300 scripts/h2d_matrix > gpurun_out/r2k_h2d_matrix.json 2> gpurun_out/r2k_h2d.err; echo "h2d rc=$?"
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2k_pytest.log
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2k_bench_n8.json 2> gpurun_out/r2k_bench_n8.err ) 2>&1 | tail -3
tail -3 gpurun_out/r2k_bench_n8.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2k_bench_n8.json'))
print('value', d['value'], 'ms', d['ms_per_step'], d['kernel_ms'], 'gather', d.get('calls_gather',{}).get('ms'), d.get('calls_gather',{}).get('calls_total'), d.get('job_ms'))
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('h2d_gbs_achieved'))
for k,v in d['e2e_text'].items():
    if isinstance(v,dict): print(k, 'ours', round(v['ours_wall_s'],2), v.get('identical_to_one_gpu'))
m=json.load(open('gpurun_out/r2k_h2d_matrix.json'))
for c in m['cases']:
    if c['case']!='pair': print(c)
import collections
pairs=[c for c in m['cases'] if c['case']=='pair']
print('pairs min/max aggregate', min(c['aggregate_gbs'] for c in pairs), max(c['aggregate_gbs'] for c in pairs))
PY
