mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:'call_scan_kernel' -s 2 -c 1 -f -o gpurun_out/prof_scan $CMD > gpurun_out/ncu_scan.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'call_series_kernel' -s 2 -c 1 -f -o gpurun_out/prof_series $CMD > gpurun_out/ncu_series.log 2>&1
ls -la gpurun_out/prof_scan.ncu-rep gpurun_out/prof_series.ncu-rep
