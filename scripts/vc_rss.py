#!/usr/bin/env python
"""Peak resident memory of the two programs on a panel of configs[2]'s size (2,000,000 slots): VERDICT r01 item 7
("host structures that break at configs[2]/[3]").  The caller program holds the panel (O(slots)) and TWO sample groups of
pinned counts (AS_GROUP_SAMPLES / AS_GROUP_MB), not the run: its peak RSS must not grow with the number of tumours.

    python scripts/vc_rss.py [out.json]        (GPU box; ~3 minutes, most of it writing the text inputs)
"""
import json
import os
import resource
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from scripts import c2_cli_parity as cli  # noqa: E402


def run_child(cmd, cwd, env):
    """wall seconds and peak RSS (MB) of one child process"""
    before = resource.getrusage(resource.RUSAGE_CHILDREN).ru_maxrss
    t = time.perf_counter()
    r = subprocess.run(["/bin/sh", "-c", "exec \"$@\"", "sh"] + cmd, cwd=cwd, env=env, capture_output=True, text=True)
    wall = time.perf_counter() - t
    assert r.returncode == 0, r.stdout[-2000:]
    peak = resource.getrusage(resource.RUSAGE_CHILDREN).ru_maxrss      # max over all children so far (kB)
    return wall, max(peak, before) / 1024.0, cli.timings(r.stderr)


def main():
    out = {}
    with tempfile.TemporaryDirectory(prefix="rss_", dir="/tmp") as td:
        shape = dict(n_amplicons=16000, n_normals=4, n_tumours=8, depth=2000, amp_len=(125, 125), seed=20183,
                     chroms=tuple(f"chr{i}" for i in range(1, 23)) + ("chrX",), somatic_rate=2e-4)
        out["shape"] = cli.stage(td, **shape)
        env = dict(os.environ, AS_TIMING="1", AS_DEVICES="0")
        ee = [str(cli.BIN / "AmpliSolveErrorEstimation"), "panel_design=panel.bed", "reference_genome=ref.fa", "germline_dir=N",
              "C_value=0.002", "coverage_cutoff=100", "default_error=0.01", "output_dir=o"]
        w, rss, ph = run_child(ee, td, env)
        out["error_estimation"] = {"wall_s": w, "peak_rss_mb_so_far": rss, "phases": ph}
        vc = [str(cli.BIN / "AmpliSolveVariantCalling"), "errorFile=o/positionSpecificNoise_0.0020.txt", "tumour_dir=T", "output_dir=v",
              "coverage_cutoff=100", "p_value=0.05"]
        for tag, extra in (("groups_of_2", {"AS_GROUP_SAMPLES": "2"}), ("one_group", {"AS_GROUP_SAMPLES": "8"})):
            # a fresh interpreter per measurement: RUSAGE_CHILDREN is a running maximum
            code = ("import json,resource,subprocess,sys,time;t=time.perf_counter();"
                    "r=subprocess.run(sys.argv[1:],capture_output=True,text=True);"
                    "print(json.dumps({'rc':r.returncode,'wall_s':time.perf_counter()-t,"
                    "'peak_rss_mb':resource.getrusage(resource.RUSAGE_CHILDREN).ru_maxrss/1024.0}))")
            r = subprocess.run([sys.executable, "-c", code] + vc, cwd=td, env=dict(env, **extra), capture_output=True, text=True)
            res = json.loads(r.stdout.strip().splitlines()[-1])
            assert res["rc"] == 0
            out["variant_calling_" + tag] = res
    print(json.dumps(out, indent=1))
    if len(sys.argv) > 1:
        Path(sys.argv[1]).write_text(json.dumps(out, indent=1) + "\n")


if __name__ == "__main__":
    main()
