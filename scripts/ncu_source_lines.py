"""Ranks CUDA source lines of an ncu report (captured with --import-source on, -lineinfo) by stall samples and
executed warp instructions:   python scripts/ncu_source_lines.py <report.ncu-rep> [top_n] [kernel-name substring]
(of a report with several kernels, the first launch whose name contains the substring)"""
import csv
import io
import subprocess
import sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
want = sys.argv[3] if len(sys.argv) > 3 else ""
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-c", "1"]
if want:
    cmd += ["-k", "regex:" + want]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, data, head = "?", [], None
kern = want or "first launch of the report"
for r in rows:
    if len(r) >= 2 and r[0] in ("File Name", "File Path"):
        fname = r[1].split("/")[-1]
    elif len(r) > 8 and r[0] == "Line No":
        head = r
        S, I = r.index("# Samples"), r.index("Instructions Executed")
    elif head and len(r) > max(S, I) and r[0].isdigit():
        try:
            data.append((int(r[S] or 0), int(r[I] or 0), f"{fname}:{r[0]}", r[1]))
        except ValueError:
            pass
tot, toti = sum(d[0] for d in data) or 1, sum(d[1] for d in data) or 1
print(f"# {kern[:150]}\n# total stall samples {tot}, warp instructions {toti}\n# samples  %smp  warp-inst  %inst  line")
for d in sorted(data, reverse=True)[:top]:
    print(f"{d[0]:8d} {100 * d[0] / tot:5.1f} {d[1]:11d} {100 * d[1] / toti:5.1f}  {d[2]:22s} {d[3].strip()[:120]}")
