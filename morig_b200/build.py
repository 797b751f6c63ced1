"""Builds `morig_b200/libmorig_b200.so` (the C-ABI library) in-tree with nvcc for sm_100a.

    python -m morig_b200.build            # build if stale
    python -m morig_b200.build --force
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.environ.get("MORIG_LIB") or os.path.join(PKG, "libmorig_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--shared"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(PKG, "..", "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    objs = []
    procs = []
    bdir = os.path.join(PKG, "build_trace" if (os.environ.get("MORIG_TRACE") == "1" or os.environ.get("MORIG_NVCC_FLAGS")) else "build")
    os.makedirs(bdir, exist_ok=True)
    for src in sources():
        obj = os.path.join(bdir, os.path.basename(src) + ".o")
        cmd = [NVCC, "-c", src, "-o", obj] + [f for f in FLAGS if f != "--shared"]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        if os.environ.get("MORIG_TRACE") == "1":         # role timeline of scripts/tc_trace.py
            cmd += ["-DMORIG_TRACE"]
        cmd += os.environ.get("MORIG_NVCC_FLAGS", "").split()      # experiment builds (e.g. -DMORIG_BSPLIT=4)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out:
            print(out)
    # no -lcuda: driver entry points (cuTensorMapEncodeTiled) are resolved at run time through cudaGetDriverEntryPoint,
    # so the library loads on the CPU-only build box too
    cmd = [NVCC, "--shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
