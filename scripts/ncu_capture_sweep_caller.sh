# ncu --set full of the noise-floor sweep caller (run under gpurun, ONE GPU); summarise with scripts/ncu_summary.py
set -e
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:'call_staged_kernelILi3ELi2ELb1ELb1' \
    -s 2 -c 1 -f -o gpurun_out/prof_sweep_caller $CMD > gpurun_out/ncu_sweep_caller.log 2>&1
ls -la gpurun_out/prof_sweep_caller.ncu-rep
