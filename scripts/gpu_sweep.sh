# kernel-variant sweep on the bench workload (device-resident step only); run under gpurun
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
for nk in ${NOISE_VARIANTS:-0 1 2 6}; do timeout 300 $B --noise-kernel $nk --call-kernel 3 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('NOISE', $nk, d['kernel_ms'], d['roofline_noise']['frac'])"; done
for ck in ${CALL_VARIANTS:-1 3 7 8 9}; do timeout 300 $B --noise-kernel 1 --call-kernel $ck | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('CALL', $ck, d['kernel_ms'], d['roofline']['frac'], d['config']['calls_per_step_rank0'])"; done
