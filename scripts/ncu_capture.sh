# ncu evidence of the bench step (run under gpurun, ONE GPU): launch list of a short bench run and --set full captures of
# the two streaming kernels of the step and of the sweep kernels.  Outputs under gpurun_out/; summarise here with
# scripts/ncu_summary.py (profiles/README.md lists the files).
set -e
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-config-legs --no-e2e-text --no-pileup-leg"
$CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
# the caller and the noise kernel of one timed step (3 warm-up steps = 6 matching launches skipped)
ncu --set full --clock-control none --import-source on -k regex:'call_staged_kernel|noise_staged_kernel' -s 6 -c 2 \
    -f -o gpurun_out/prof_main $CMD > gpurun_out/ncu_main.log 2>&1
# the noise-floor sweep: single-pass noise kernel for five values, and the deferred caller (scan / resolve / series).  Matching
# launches in bench order: 7 x noise_pattern (2 warm-ups + 5 timed), then 7 x (scan, resolve, series): skip six, take the last
# pattern launch and the first two rounds of the caller's three kernels
ncu --set full --clock-control none --import-source on -k regex:'noise_pattern_kernel|call_scan_kernel|call_resolve_kernel|call_series_kernel' \
    -s 6 -c 7 -f -o gpurun_out/prof_sweep $CMD > gpurun_out/ncu_sweep.log 2>&1
ls -la gpurun_out/*.ncu-rep
