mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'call_' -c 60 --csv --log-file gpurun_out/r2e_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2e_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2e_launches.csv')) if len(r)>5]
h=next(r for r in rows if 'Kernel Name' in r); i={k:n for n,k in enumerate(h)}
for r in rows:
    if r is h: continue
    print(r[i['ID']], r[i['Kernel Name']][:60], r[i['Metric Name']], r[i['Metric Value']], r[i['Metric Unit']])
PY
