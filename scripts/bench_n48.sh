mkdir -p gpurun_out
for n in 4 8; do
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/final_bench_n$n.json 2> gpurun_out/final_bench_n$n.err ) 2>&1 | tail -3
python - <<PY
import json
d=json.load(open('gpurun_out/final_bench_n$n.json'))
print('N=$n value', d['value'], 'ms', d['ms_per_step'], 'gather', d.get('calls_gather',{}).get('ms'), d.get('job_ms',{}).get('value'), 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('h2d_gbs_achieved'))
for k,v in d['e2e_text'].items():
    if isinstance(v,dict): print(' ', k, 'ours', [round(x,2) for x in v['ours_wall_s_runs']], v.get('identical_to_one_gpu'))
PY
done
