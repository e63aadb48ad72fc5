"""Compile the CUDA extension in-tree: in_cwave_b200/libicw_b200.so (sm_100a only).

The library is the product: kernels + C ABI (include/icw_b200.h).  It cross-compiles without a
GPU.  -fmad=false is load-bearing: the half-band recurrences and the modulator arithmetic must
not be contracted into FMAs (DESIGN.md "numerics"); every fused operation in the sources is an
explicit fma().
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libicw_b200.so"
SOURCES = ["icw_api.cu", "icw_kernels.cu", "icw_fused.cu", "icw_scan.cu", "icw_mt.cu", "icw_crc.cu", "icw_chainmt.cu"]
NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
]


def nvcc_path() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built")
    return cand


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + \
        list(CSRC.glob("*.inc")) + [PKG.parent / "include" / "icw_b200.h", Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        if not PLUGIN.exists() or PLUGIN.stat().st_mtime < (PKG / "host" / "icw_plugin.c").stat().st_mtime:
            build_plugin()
        return LIB
    extra = os.environ.get("ICW_NVCC_EXTRA", "").split()      # A/B experiments: -DICW_APPLY_THREADS=384 ...
    cmd = [nvcc_path(), *NVCC_FLAGS, *extra, "-o", str(LIB)] + [str(CSRC / s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    (PKG / "build.log").write_text(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed; see in_cwave_b200/build.log")
    if verbose:
        print(log)
    build_plugin()
    return LIB


PLUGIN = PKG / "libicw_plugin.so"


def build_plugin() -> Path:
    """The host C layer (reference entry points over the C ABI): plain gcc, links libicw_b200.so."""
    cmd = ["gcc", "-std=c99", "-O2", "-Wall", "-fPIC", "-shared", "-I", str(PKG.parent / "include"),
           "-o", str(PLUGIN), str(PKG / "host" / "icw_plugin.c"), str(PKG / "host" / "icw_config.c"), "-L", str(PKG), "-licw_b200",
           "-Wl,-rpath,$ORIGIN"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("gcc failed on host/icw_plugin.c / icw_config.c")
    return PLUGIN


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print("built", LIB)
