"""Summarises ncu reports brought back in gpurun_out/ (run in the dev container, no GPU needed):

    python scripts/ncu_summary.py gpurun_out/prof_main.ncu-rep [gpurun_out/prof_sweep.ncu-rep ...] > profiles/<name>.txt
    python scripts/ncu_summary.py --launches gpurun_out/launches.csv > profiles/<name>_summary.txt
    python scripts/ncu_summary.py --traffic gpurun_out/prof_main.ncu-rep     # rewrites profiles/ncu_traffic.json
"""
import csv
import io
import json
import re
import subprocess
import sys
from collections import OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "sm__cycles_elapsed.avg", "launch__grid_size",
        "launch__block_size", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units = rows[0], rows[1]
    return head, units, rows[2:]


def to_bytes(value, unit):
    v = float(value.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def summarise(reps):
    for rep in reps:
        head, units, rows = raw_rows(rep)
        idx = {h: i for i, h in enumerate(head)}
        print(f"# ncu --set full --clock-control none --import-source on, summary of {Path(rep).name}")
        for r in rows:
            print(f"\nKernel Name = {r[idx['Kernel Name']]}")
            for k in KEEP:
                if k in idx:
                    print(f"{k} = {r[idx[k]]} {units[idx[k]]}")


def traffic(reps):
    kernels = OrderedDict()
    rows_all = []
    for rep in reps:
        head, units, rows = raw_rows(rep)
        rows_all += [(head, units, r) for r in rows]
    for head, units, r in rows_all:
        idx = {h: i for i, h in enumerate(head)}
        name = re.sub(r"^void\s+", "", r[idx["Kernel Name"]]).split("<")[0].split("(")[0].replace("asdev::", "")
        kernels[name] = {"dram_bytes_read": to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]),
                         "dram_bytes_write": to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]]),
                         "dram_pct_of_peak": float(r[idx["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]]),
                         "duration_ms": float(r[idx["gpu__time_duration.sum"]]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[idx["gpu__time_duration.sum"]], 1.0),
                         "kernel": r[idx["Kernel Name"]].split("(")[0]}
    doc = {"workload": "c3: synthetic 500-gene panel shard per GPU",
           "source": f"{', '.join(Path(r).name for r in reps)} (ncu --set full --clock-control none, one launch each; scripts/ncu_capture.sh)",
           "kernels": kernels}
    (ROOT / "profiles" / "ncu_traffic.json").write_text(json.dumps(doc, indent=1))
    print(json.dumps(doc, indent=1))


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    head = next(r for r in rows if "Kernel Name" in r)
    idx = {h: i for i, h in enumerate(head)}
    acc = OrderedDict()
    for r in rows:
        if r is head or r[idx["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]])
        v = float(r[idx["Metric Value"]].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[idx["Metric Unit"]], 1e-6)
        acc.setdefault(name, []).append(v)
    print("# ncu --metrics gpu__time_duration.sum --clock-control none: per-kernel launches and mean ms of\n"
          "# `bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline` (cold-cache, serialised launches: compare SHARES, not absolutes)\n")
    for k, v in acc.items():
        print(f"{k[:110]:110s} launches={len(v):3d} mean_ms={sum(v) / len(v):9.4f}")


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2])
    elif sys.argv[1] == "--traffic":
        traffic(sys.argv[2:])
    else:
        summarise(sys.argv[1:])
