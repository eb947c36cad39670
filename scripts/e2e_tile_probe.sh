mkdir -p gpurun_out
for TM in 64 256; do
  AS_HOST_TILE_MB=$TM timeout 90 python bench.py --steps 3 --warmup 3 --e2e-steps 3 --no-sweep --no-config-legs --no-e2e-text --no-pileup-leg --no-cpu-baseline > gpurun_out/tile_$TM.json 2> gpurun_out/tile_$TM.err
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/tile_$TM.json') if l.startswith('{')][-1]); e=d['e2e']; print('TILE_MB $TM e2e ms', e['ms_per_step'], e.get('noise_call_ms'), e.get('caller_call_ms'), e.get('pcie_h2d_gbs_measured'))
" 2>&1 | tail -1
done
