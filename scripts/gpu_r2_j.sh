mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2j_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2j_pytest.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2j_bench_n2.json 2> gpurun_out/r2j_bench_n2.err ) 2>&1 | tail -3
tail -5 gpurun_out/r2j_bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2j_bench_n2.json'))
print('value', d['value'], 'ms', d['ms_per_step'], d['kernel_ms'], 'gather', d.get('calls_gather'), d.get('job_ms'))
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('h2d_gbs_achieved'))
for k,v in d['e2e_text'].items():
    if isinstance(v,dict): print(k, 'ours', round(v['ours_wall_s'],2), v.get('identical_to_one_gpu'), v['ours']['vc_phases_s'])
PY
