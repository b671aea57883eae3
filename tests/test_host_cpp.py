"""The reference's own gtest cases, restated against the drop-in C++ headers (tests/cpp/test_host_api.cpp)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_host_api.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "test_host_api.bin")


def build_exe():
    lib = os.path.join(ROOT, "algotoolkit_b200")
    if os.path.exists(EXE) and os.path.getmtime(EXE) >= max(
            os.path.getmtime(SRC), os.path.getmtime(os.path.join(lib, "libatk_cuda.so"))):
        return EXE
    cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"), "-I",
           os.path.join(lib, "host"), SRC, "-o", EXE, "-L", lib, "-latk_cuda", f"-Wl,-rpath,{lib}", "-Wl,-rpath,$ORIGIN/../../algotoolkit_b200"]
    subprocess.run(cmd, check=True)
    return EXE


def test_host_headers_compile():
    """CPU side: the drop-in headers compile and link against libatk_cuda.so with the reference's call shapes."""
    from algotoolkit_b200 import build as atk_build

    atk_build.build()
    assert os.path.exists(build_exe())


@pytest.mark.gpu
def test_reference_unit_tests_on_gpu():
    exe = build_exe()
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "data", "bcsstk01.mtx")], capture_output=True, text=True,
                       timeout=300)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "0 failed" in r.stdout
