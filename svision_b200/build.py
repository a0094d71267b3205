"""In-tree build of ``libsvx.so`` (hand-written sm_100a CUDA + the C-ABI) with nvcc.

``python -m svision_b200.build`` or ``__graft_entry__.build()``.  The library has no torch /
pybind dependency: it links the CUDA runtime statically and resolves the one driver symbol it
needs (``cuTensorMapEncodeTiled``) at run time, so the same ``.so`` built in the CPU-only
container runs on the GPU box."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsvx.so")
SOURCES = ["encoder.cu", "front.cu", "layer_tc.cu", "cnn_aux.cu", "svx_api.cu", "svx_multi.cpp",
           "host_bed.cpp", "host_pairs.cpp", "host_calls.cpp"]
HEADERS = ["common.cuh", "kernels.h", "encoder_bitmap.cuh", os.path.join("..", "..", "include", "svx.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    objs = []
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    common = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", *ARCH]
    if verbose:
        common += ["-Xptxas", "-v"]
    common += os.environ.get("SVX_NVCC_FLAGS", "").split()      # development: e.g. -DSVX_COLLECTOR_A=1
    procs = []
    for src in SOURCES:
        obj = os.path.join(bdir, os.path.splitext(src)[0] + ".o")
        objs.append(obj)
        cmd = [_nvcc(), *common, "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    failed = False
    for src, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {src} ---\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    link = [_nvcc(), "-shared", *ARCH, "-o", LIB, *objs, "-cudart", "static"]
    subprocess.check_call(link)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
