"""Build libsfron_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

The library links only the CUDA runtime (statically): no torch, no Python.  It is
loaded with ctypes by `capi.py`.  nvcc cross-compiles without a GPU, so this runs in
the CPU-only build container; the resulting .so travels to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB_NAME = "libsfron_b200.so"
LIB_PATH = os.path.join(HERE, LIB_NAME)
STAMP_PATH = os.path.join(HERE, ".libsfron_b200.stamp")

SOURCES = ["api.cu", "fisher.cu", "mask.cu", "select.cu", "update.cu", "extras.cu", "peer.cu", "peer_tma.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",            # no implicit FMA contraction: explicit __fmaf_rn only
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libsfron_b200.so cannot be built (no CPU fallback exists)")


def _source_digest() -> str:
    h = hashlib.sha256()
    files = [os.path.join(CSRC, s) for s in SOURCES] + [
        os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "update_core.cuh"), os.path.join(CSRC, "peer_common.cuh"), os.path.join(INCLUDE, "sfron_b200.h"), __file__]
    for f in files:
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_fresh() -> bool:
    if not (os.path.exists(LIB_PATH) and os.path.exists(STAMP_PATH)):
        return False
    with open(STAMP_PATH) as fh:
        return fh.read().strip() == _source_digest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the library if sources changed; returns the path of the .so."""
    if not force and is_fresh():
        return LIB_PATH
    nvcc = _nvcc()
    objs = []
    build_dir = os.path.join(HERE, "build")
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(build_dir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-I", INCLUDE, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(f"--- {src}\n{out}")
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH, *objs,
            "-Xlinker", "--exclude-libs=ALL"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    with open(STAMP_PATH, "w") as fh:
        fh.write(_source_digest())
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
