"""Builds the `fastsolver` Python module (pybind11) from algotoolkit_b200/python/pybind.cpp and the reference-named
host headers in algotoolkit_b200/host, linked against the in-tree libatk_cuda.so.

    python -m algotoolkit_b200.build_host [--force]
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "python", "pybind.cpp")
OUT = os.path.join(HERE, "fastsolver" + sysconfig.get_config_var("EXT_SUFFIX"))
HOST_CXX = "/usr/bin/g++"


def _deps():
    deps = [SRC, os.path.join(ROOT, "include", "atk_cuda.h"), os.path.abspath(__file__)]
    for base, _, files in os.walk(os.path.join(HERE, "host")):
        deps += [os.path.join(base, f) for f in files if f.endswith(".hpp")]
    return deps


def build(force: bool = False) -> str:
    import pybind11

    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in _deps()):
        return OUT
    cmd = [HOST_CXX, "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-fvisibility=hidden",
           "-I", os.path.join(ROOT, "include"), "-I", pybind11.get_include(), "-I", sysconfig.get_paths()["include"],
           SRC, "-o", OUT, "-L", HERE, "-latk_cuda", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("building the fastsolver module failed")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
