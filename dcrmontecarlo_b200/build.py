"""Builds libwost.so (CUDA kernels + C ABI) in-tree for sm_100a.

    python -m dcrmontecarlo_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  ``-fmad=false`` is part of the numerical contract (see
csrc/wost_device.cuh): the reference computes its fp32 geometry without fused multiply-adds.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
LIB = PKG / "libwost.so"
SOURCES = [PKG / "csrc" / "wost_lib.cu"]
HEADERS = [PKG / "csrc" / "wost_device.cuh", ROOT / "include" / "wost.h"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def nvcc_cmd(out: Path, extra=()) -> list[str]:
    return [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
            "-Xcompiler", "-fPIC,-O2", "-shared", "-I", str(ROOT / "include"), "-I", str(PKG / "csrc"),
            *extra, "-o", str(out), *map(str, SOURCES)]


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(p.stat().st_mtime > t for p in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> Path:
    if force or needs_build():
        cmd = nvcc_cmd(LIB, ["-Xptxas", "-v"] if verbose else [])
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
