AS_TIMING=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-sweep --no-config-legs --no-e2e-text --no-cpu-baseline --e2e-steps 2 > gpurun_out/r2n.json 2> gpurun_out/r2n.err
grep AS_TIMING gpurun_out/r2n.err | tail -24
python -c "
import json; d=json.load(open('gpurun_out/r2n.json')); print(d['e2e']['ms_per_step'], d['e2e']['noise_call_ms'], d['e2e']['caller_call_ms'], d['e2e']['pcie_h2d_gbs_measured'], d['e2e']['h2d_gbs_achieved'])"
