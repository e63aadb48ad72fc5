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
SOURCES = ["icw_api.cu", "icw_kernels.cu", "icw_fused.cu", "icw_split.cu", "icw_scan.cu", "icw_mt.cu", "icw_crc.cu", "icw_chainmt.cu", "icw_comm.cu", "icw_sfused.cu"]
HOST_SOURCES = ["icw_hbconv.cpp"]      # plain g++ (binary128 arithmetic: nvcc's front end does not take __float128)
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
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cpp")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + \
        list(CSRC.glob("*.inc")) + [PKG.parent / "include" / "icw_b200.h", Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps)


def _compile_one(args):
    src, obj, extra = args
    if src.suffix == ".cpp":
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-c", "-o", str(obj), str(src)]
    else:
        cmd = [nvcc_path(), *[f for f in NVCC_FLAGS if f != "-shared"], *extra, "-c", "-o", str(obj), str(src)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    return src.name, " ".join(cmd), res.returncode, res.stdout + res.stderr


def build(force: bool = False, verbose: bool = False) -> Path:
    """One nvcc per translation unit, in parallel (objects cached under in_cwave_b200/build/), then one link.
    The units share no device symbols, so no relocatable device code is needed."""
    if not force and not needs_build():
        if not PLUGIN.exists() or PLUGIN.stat().st_mtime < max((PKG / "host" / f).stat().st_mtime for f in PLUGIN_SOURCES):
            build_plugin()
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    extra = os.environ.get("ICW_NVCC_EXTRA", "").split()      # A/B experiments: -DICW_APPLY_THREADS=384 ...
    objdir = PKG / "build"
    objdir.mkdir(exist_ok=True)
    stamp = objdir / "flags.txt"
    flags_now = " ".join(NVCC_FLAGS + extra)
    if not stamp.exists() or stamp.read_text() != flags_now:
        force = True
    hdr_t = max(d.stat().st_mtime for d in list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list(CSRC.glob("*.inc")) +
                [PKG.parent / "include" / "icw_b200.h", Path(__file__)])
    jobs, objs = [], []
    for s in SOURCES + HOST_SOURCES:
        src, obj = CSRC / s, objdir / (Path(s).stem + ".o")
        objs.append(obj)
        if force or not obj.exists() or obj.stat().st_mtime < max(src.stat().st_mtime, hdr_t):
            jobs.append((src, obj, extra))
    logs = []
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for name, cmd, rc, log in ex.map(_compile_one, jobs):
            logs.append(f"### {name}\n{cmd}\n{log}")
            if rc != 0:
                (PKG / "build.log").write_text("\n".join(logs))
                sys.stderr.write(log)
                raise RuntimeError(f"nvcc failed on {name}; see in_cwave_b200/build.log")
    stamp.write_text(flags_now)
    cmd = [nvcc_path(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB)] + [str(o) for o in objs] + LINK_LIBS
    res = subprocess.run(cmd, capture_output=True, text=True)
    logs.append("### link\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    prev = (PKG / "build.log").read_text() if (PKG / "build.log").exists() and not force else ""
    (PKG / "build.log").write_text("\n".join(logs) + ("\n### earlier\n" + prev[:200000] if prev else ""))
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("link failed; see in_cwave_b200/build.log")
    if verbose:
        print("\n".join(logs))
    build_plugin()
    return LIB


LINK_LIBS: list[str] = ["-ldl"]
PLUGIN_SOURCES = ["icw_plugin.c", "icw_config.c"]
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
