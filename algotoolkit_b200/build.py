"""In-tree build of libatk_cuda.so (the C-ABI CUDA library) with nvcc for sm_100a.

    python -m algotoolkit_b200.build [--force] [--verbose]

Each csrc/*.cu is compiled to build/<name>.o and the objects are linked into
algotoolkit_b200/libatk_cuda.so.  nvcc cross-compiles without a GPU; the .so travels to the GPU box with the
repository snapshot.  -lineinfo keeps source correlation for ncu; -fmad=false (plus explicit _rn intrinsics in the
kernels) keeps multiply and add separately rounded, as in the reference build.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libatk_cuda.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
HOST_CXX = "/usr/bin/g++"  # dynamic libstdc++ (the /opt/gcc wrapper links it statically, which breaks inside python)

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
    "-ccbin", HOST_CXX,
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-Wall,-Wno-unused-function",
    "-Xptxas", "-v",
    "-I", INCLUDE,
]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, checked: bool = False, profile: bool = False) -> str:
    """checked=True builds libatk_cuda_checked.so with -DATK_CHECKED (device-side assertions on tile / halo indices and on
    the peer protocols' sequence numbers); ATK_LIB=checked makes capi load it."""
    global OBJ, LIB
    if checked:
        OBJ = os.path.join(HERE, "build_checked")
        LIB = os.path.join(HERE, "libatk_cuda_checked.so")
    elif profile:  # -DATK_REDUCE_PROFILE: %globaltimer stamps at the five stations of every in-kernel reduction
        OBJ = os.path.join(HERE, "build_profile")
        LIB = os.path.join(HERE, "libatk_cuda_profile.so")
    else:
        OBJ = os.path.join(HERE, "build")
        LIB = os.path.join(HERE, "libatk_cuda.so")
    flags = NVCC_FLAGS + (["-DATK_CHECKED"] if checked else []) + (["-DATK_REDUCE_PROFILE"] if profile else [])
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(INCLUDE, "atk_cuda.h"))
    headers.append(os.path.abspath(__file__))
    jobs = []
    for s in _sources():
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s[:-3] + ".o")
        if force or _stale(obj, [src] + headers):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [NVCC] + flags + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return job, r

    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for (src, obj), r in ex.map(compile_one, jobs):
                log = os.path.join(OBJ, os.path.basename(src) + ".log")
                with open(log, "w") as f:
                    f.write(r.stdout + r.stderr)
                if r.returncode != 0:
                    sys.stderr.write(r.stdout + r.stderr)
                    raise RuntimeError(f"nvcc failed on {src}")
                if verbose:
                    sys.stderr.write(r.stderr)
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in _sources()]
    if force or jobs or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-ccbin", HOST_CXX, "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link of libatk_cuda.so failed")
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, checked="--checked" in sys.argv,
                 profile="--profile" in sys.argv)
    print(path)
