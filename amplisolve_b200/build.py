"""Build libamplisolve_b200.so (CUDA kernels + C ABI + host code) in-tree for sm_100a.

    python -m amplisolve_b200.build [--force]

nvcc cross-compiles without a GPU.  -fmad=false: the reference runs on x86-64 without FMA and the
parity-critical fp32/fp64 expressions must not be contracted (they also use explicit _rn intrinsics).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB = PKG / "lib" / "libamplisolve_b200.so"
BIN = PKG / "bin"
CU_SOURCES = ["as_kernels.cu", "as_capi.cu", "as_sort.cu", "as_fisher.cu"]
CXX_SOURCES = ["as_host.cpp"]
HEADERS = ["as_device.cuh", "as_noise.cuh", "as_pipeline.cuh", "as_kernels.h", "as_wire.h", "../../include/amplisolve_b200.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC,-O2,-ffp-contract=off", "-cudart", "static"]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(exe).exists():
        raise RuntimeError("nvcc not found: libamplisolve_b200.so cannot be built (there is no CPU fallback)")
    return exe


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    srcs = [CSRC / s for s in CU_SOURCES + CXX_SOURCES]
    deps = srcs + [CSRC / h for h in HEADERS] + [Path(__file__)]
    if force or _stale(LIB, deps):
        LIB.parent.mkdir(parents=True, exist_ok=True)
        cmd = [nvcc(), *NVCC_FLAGS, "-shared", "-o", str(LIB), *map(str, srcs), "-I", str(ROOT / "include")]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed building libamplisolve_b200.so")
    mains = CSRC / "as_main.cpp"
    if mains.exists():
        BIN.mkdir(exist_ok=True)
        for prog, macro in (("AmpliSolveErrorEstimation", "AS_MAIN_EE"), ("AmpliSolveVariantCalling", "AS_MAIN_VC")):
            out = BIN / prog
            if force or _stale(out, [mains, LIB]):
                cmd = ["g++", "-O2", "-std=c++17", f"-D{macro}", "-o", str(out), str(mains), "-I", str(ROOT / "include"),
                       f"-L{LIB.parent}", "-lamplisolve_b200", "-Wl,-rpath,$ORIGIN/../lib"]
                r = subprocess.run(cmd, capture_output=True, text=True)
                if r.returncode != 0:
                    sys.stderr.write(r.stdout + r.stderr)
                    raise RuntimeError(f"g++ failed building {prog}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
