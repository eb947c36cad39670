# Two-GPU check of the product's multi-GPU path (under gpurun --gpus 2): the as_create_multi tests and a short N=2 bench line
# with the e2e and e2e_text legs (programs on both GPUs, cold and through the resident service).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/validate_pytest_n2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/validate_pytest_n2.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-sweep --no-config-legs --no-pileup-leg --no-cpu-baseline > gpurun_out/validate_bench_n2.json 2> gpurun_out/validate_bench_n2.err ) 2>&1 | tail -3
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/validate_bench_n2.json') if l.startswith('{')][-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d.get('calls_gather'), d.get('job_ms'))
for k, v in d.get('e2e_text', {}).items():
    if isinstance(v, dict):
        print(k, 'ours', [round(x, 2) for x in v.get('ours_wall_s_runs', [])], 'served', [round(x, 3) for x in v.get('ours_served', {}).get('wall_s_runs', [])], 'ref', v.get('reference_wall_s'), v.get('error'))
    else:
        print(k, v)
PY
