# caller-kernel variants on the c3 bench workload and on the c5 shape (device-resident step only); run under gpurun
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
C5="--slots 100000 --normals 100 --tumours 10000 --depth 50000 --somatic-rate 0.0005 --vaf 0.005 0.01"
for ck in ${CALL_VARIANTS:-11 13 14 15 16 3}; do
  timeout 300 $B --call-kernel $ck | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C3 CALL', $ck, round(d['kernel_ms']['caller'],4), round(d['roofline']['frac'],4), d['config']['calls_per_step_rank0'])"
  timeout 300 $B $C5 --call-kernel $ck | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C5 CALL', $ck, round(d['kernel_ms']['caller'],4), round(d['roofline']['frac'],4), d['config']['calls_per_step_rank0'])"
done
