"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol include/atk_cuda.h declares,
fails loudly without a device, and the product never routes through the oracle."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "atk_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(atk_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from algotoolkit_b200 import build as atk_build
    from algotoolkit_b200 import capi

    atk_build.build()
    lib = ctypes.CDLL(capi.LIB_PATH)
    declared = header_symbols()
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/atk_cuda.h but not exported"
    assert sorted(capi.EXPORTS) == declared
    assert lib.atk_abi_version() == 2
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert " T atk_cg_solve" in out and " T atk_gmres_solve" in out


def test_no_cpu_fallback_without_device():
    from algotoolkit_b200 import capi

    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        capi.Context(0)


def test_library_is_sm100a_and_has_cp_async_kernels():
    from algotoolkit_b200 import capi

    out = subprocess.run(["cuobjdump", "--list-elf", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "LDGSTS" in sass  # cp.async: the shared-memory plane pipeline of the fused CG kernel


def test_product_never_touches_the_oracle():
    """Nothing under algotoolkit_b200/ or include/ may import, link or call anything under oracle/."""
    bad = []
    for base in ("algotoolkit_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            if "build" in dirpath.split(os.sep):
                continue
            for f in files:
                if not f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    continue
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"(from|import)\s+oracle|liboracle|libatk_ref|krylov_oracle|pyoracle|oracle/|orc_[a-z]+\(|ref_[a-z_]+\(", txt):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
    deps = subprocess.run(["ldd", os.path.join(ROOT, "algotoolkit_b200", "libatk_cuda.so")], capture_output=True, text=True).stdout
    assert "oracle" not in deps and "atk_ref" not in deps


def test_fastsolver_module_builds_and_imports():
    import sys

    from algotoolkit_b200 import build_host

    build_host.build()
    sys.path.insert(0, os.path.join(ROOT, "algotoolkit_b200"))
    import fastsolver

    for name in ("Vector", "SparseMatrix", "GMRES", "ConjugateGrad", "Poisson3DSolver", "read_matrix_market", "matvec_mul",
                 "BoundaryType"):
        assert hasattr(fastsolver, name)
    A = fastsolver.SparseMatrix(3, 3)  # host-side container semantics need no GPU
    A.addValue(0, 2, 3.0)
    A.addValue(1, 0, 4.0)
    A.addValue(2, 2, 0.0)  # explicit zero: not stored
    A.finalize()
    assert A(0, 2) == 3.0 and A(1, 0) == 4.0 and A(2, 2) == 0.0 and A.nnz() == 2
    with pytest.raises(IndexError):
        A(3, 0)
    v = fastsolver.Vector(4)
    v[1] = 2.5
    assert v[1] == 2.5 and v.size() == 4
    with pytest.raises(IndexError):
        v[4]


def test_header_is_plain_c_and_a_c_program_links(tmp_path):
    """The drop-in boundary is a C ABI: include/atk_cuda.h must compile as C99 (no C++ in the signatures) and a C
    program must link against libatk_cuda.so and get a clean status without a device (no compute)."""
    import subprocess

    src = tmp_path / "abi.c"
    src.write_text(
        '#include <stdio.h>\n#include "atk_cuda.h"\n'
        "int main(void) {\n"
        "    atk_context_t ctx = 0;\n"
        "    int k0 = -1, k1 = -1;\n"
        "    if (atk_abi_version() != ATK_ABI_VERSION) return 2;\n"
        "    if (atk_slab_partition(10, 1, 3, &k0, &k1) != ATK_OK || k0 >= k1) return 3;\n"
        "    int rc = atk_context_create(0, &ctx);\n"
        '    printf("%d %s\\n", rc, rc == ATK_OK ? "device" : atk_last_error());\n'
        "    if (rc == ATK_OK) atk_context_destroy(ctx);\n"
        "    return 0;\n}\n")
    exe = tmp_path / "abi.bin"
    lib_dir = os.path.join(ROOT, "algotoolkit_b200")
    cmd = ["/usr/bin/gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o",
           str(exe), "-L", lib_dir, "-latk_cuda", "-Wl,-rpath," + lib_dir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    status = int(r.stdout.split()[0])
    assert status in (0, 4), r.stdout  # ATK_OK on a GPU box, ATK_ERR_CUDA without a device - never a crash, never a CPU path


def test_hot_kernels_do_not_spill():
    """The fused CG kernels run at a fixed register budget (7 CTAs per SM); code added to them has twice pushed ptxas into
    spilling inside the plane loop, which costs 30 % of the iteration without failing any parity test.  The build log
    (-Xptxas -v) must show at most a few bytes of spills for them."""
    import re

    from algotoolkit_b200 import build as atk_build

    atk_build.build()
    log = open(os.path.join(atk_build.OBJ, "atk_stencil.cu.log")).read()
    seen = 0
    for kernel in ("cg_persistent_kernelILi32ELi4E", "stencil7_cg_fused_smem_kernelILi32ELi4E", "stencil7_smem_kernelILi32ELi4ELb0ELb0E"):
        m = re.search(kernel + r".*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", log, re.S)
        assert m, f"{kernel} not found in the ptxas log"
        assert int(m.group(2)) <= 16 and int(m.group(3)) <= 16, (kernel, m.groups())
        seen += 1
    assert seen == 3
    # the Arnoldi step kernels sit at the 128-register cap of a 512-thread CTA (w in registers / sixteen loads in flight)
    glog = open(os.path.join(atk_build.OBJ, "atk_gmres.cu.log")).read()
    for kernel in ("arnoldi_mgs_ring_kernelILi512ELi14E", "arnoldi_mgs_kernelILi512ELb0E", "arnoldi_mgs_kernelILi512ELb1E"):
        m = re.search(kernel + r".*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", glog, re.S)
        assert m, f"{kernel} not found in the ptxas log"
        assert int(m.group(2)) == 0 and int(m.group(3)) == 0, (kernel, m.groups())
