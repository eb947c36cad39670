#!/usr/bin/env python
"""Instruction-level evidence for profiles/ (runs in the dev container, no GPU): the SASS of the shipped streaming kernels
around their bulk-copy (TMA, 1-D form: UBLKCP) and mbarrier (SYNCS) instructions, and the ptxas register / spill table of
every kernel in libamplisolve_b200.so.

    python scripts/sass_proof.py        # writes profiles/r02_sass_excerpts.txt and profiles/r02_ptxas_registers.txt
"""
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from amplisolve_b200 import build as as_build  # noqa: E402

OBJ = ROOT / "amplisolve_b200" / "lib" / "obj"
KERNELS = [("as_kernels.o", "noise_staged_kernelILi4ELi3E"), ("as_kernels.o", "call_staged_kernelILi3ELi2ELb1ELb0E"),
           ("as_call_deferred.o", "call_scan_kernelILi3ELi2E"), ("as_noise_pattern.o", "noise_pattern_kernelILi5ELi4ELi3ELi3ELb1E")]
WANT = re.compile(r"UBLKCP|SYNCS|LDS\.128|BAR\.|ATOMG|RED\.|UTMA|ELECT")


def sass(obj, pattern):
    names = subprocess.run(["cuobjdump", "-sass", str(OBJ / obj)], capture_output=True, text=True).stdout
    fn = next(l.split(":", 1)[1].strip() for l in names.splitlines() if "Function :" in l and pattern in l)
    out = subprocess.run(["cuobjdump", "-sass", "-fun", fn, str(OBJ / obj)], capture_output=True, text=True).stdout
    lines = [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l).rstrip() for l in out.splitlines() if re.search(r"/\*[0-9a-f]{4}\*/", l)]
    return fn, lines


def main():
    as_build.build()
    doc = ["# cuobjdump -sass of the shipped streaming kernels (sm_100a): every bulk copy global -> shared (UBLKCP = the 1-D form of",
           "# cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes, issued by the producer thread), every mbarrier operation",
           "# (SYNCS.*: init / arrive.expect_tx / try_wait / arrive) and the first 128-bit shared-memory reads of the consumers, with",
           "# instruction counts per mnemonic class.  Regenerate: python scripts/sass_proof.py", ""]
    for obj, pat in KERNELS:
        fn, lines = sass(obj, pat)
        ops = {}
        for l in lines:
            m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
            if m:
                ops[m.group(1).split(".")[0]] = ops.get(m.group(1).split(".")[0], 0) + 1
        doc.append(f"## {fn}")
        doc.append(f"   {len(lines)} instructions; " + ", ".join(f"{k} {v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:14]))
        shown = 0
        for l in lines:
            if WANT.search(l) and (("LDS.128" not in l) or shown < 40):
                if "LDS.128" in l and sum(1 for d in doc[-12:] if "LDS.128" in d) >= 4:
                    continue
                doc.append(l)
                shown += 1
        doc.append("")
    (ROOT / "profiles" / "r02_sass_excerpts.txt").write_text("\n".join(doc) + "\n")
    # ptxas -v of every kernel
    rows = ["# ptxas -v (nvcc -Xptxas=-v, sm_100a) of every kernel of libamplisolve_b200.so: registers, spills, static shared memory.",
            "# Regenerate: python scripts/sass_proof.py", ""]
    for src in as_build.CU_SOURCES:
        cmd = [as_build.nvcc(), *as_build.NVCC_FLAGS, "-Xptxas=-v", "-c", "-o", "/dev/null", str(as_build.CSRC / src), "-I", str(ROOT / "include")]
        err = subprocess.run(cmd, capture_output=True, text=True).stderr
        name = None
        for l in err.splitlines():
            m = re.search(r"Compiling entry function '(\S+)'", l)
            if m:
                name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", l)
            if m and name:
                spill = m.groups()
            m = re.search(r"Used (\d+) registers(?:, used \d+ barriers)?(.*)", l)
            if m and name:
                rows.append(f"{src:22s} {name[:100]:100s} regs {int(m.group(1)):3d}  stack {spill[0]:>4s} B  spill st/ld {spill[1]:>4s}/{spill[2]:>4s} B {m.group(2).strip(', ')}")
                name = None
    (ROOT / "profiles" / "r02_ptxas_registers.txt").write_text("\n".join(rows) + "\n")
    print(f"{len(doc)} SASS lines, {len(rows) - 3} kernels")


if __name__ == "__main__":
    main()
