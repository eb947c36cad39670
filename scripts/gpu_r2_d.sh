mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --noise-kernel 7"
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:'noise_pattern_kernelILi5' -s 2 -c 1 -f -o gpurun_out/prof_pat5 $CMD > gpurun_out/ncu_pat5.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:'noise_pattern_kernelILi1' -s 6 -c 1 -f -o gpurun_out/prof_pat1 $CMD > gpurun_out/ncu_pat1.log 2>&1
ls -la gpurun_out/*.ncu-rep
