"""Builds libtbv_b200.so (hand-written sm_100a CUDA behind the C-ABI of include/tbv_b200.h) in-tree with nvcc.

nvcc cross-compiles without a GPU, so this runs in the CPU-only build container; the resulting .so travels to the
GPU box with the repository snapshot.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtbv_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",  # fp64 paths must keep the reference's mul-then-add rounding (no FMA contraction)
    "-Xcompiler", "-fPIC", "-shared",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


OBJ_DIR = os.path.join(HERE, "build")
COMPILE_FLAGS = [f for f in NVCC_FLAGS if f != "-shared"]


def _headers():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "tbv_b200.h")]


def _obj(src: str) -> str:
    return os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build() -> bool:
    return _stale(LIB, sources() + _headers())


def build(force: bool = False, verbose: bool = False) -> str:
    """One nvcc -c per .cu (only the stale ones, in parallel), then one link into libtbv_b200.so."""
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libtbv_b200.so")
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdrs = _headers()
    todo = [s for s in sources() if force or _stale(_obj(s), [s] + hdrs)]
    procs = [(s, subprocess.Popen([nvcc] + COMPILE_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", _obj(s), s],
                                  stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)) for s in todo]
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(out)
        elif verbose:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("nvcc failed building libtbv_b200.so")
    res = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + [_obj(s) for s in sources()],
                         capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libtbv_b200.so")
    return LIB


def _exports():
    """Every function include/tbv_b200.h declares — the C-ABI the library must export."""
    import re
    hdr = open(os.path.join(HERE, "..", "include", "tbv_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(tbv_[a-z0-9_]+)\s*\(", hdr)))


EXPORTS = _exports()


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
