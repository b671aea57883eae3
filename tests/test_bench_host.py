"""Host-side logic of bench.py that needs no GPU: workload naming shared by both arms, the sample-grid choice of the reference
arm, and the recognition of the GMRES path that ran (bytes moved per row and projection) from its launch count."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_bench():
    spec = importlib.util.spec_from_file_location("atk_bench", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


def test_both_arms_name_the_same_workload():
    b = load_bench()
    assert b.workload_name(512) == "poisson3d_cg_512^3"
    assert b.pick_sample_grid(128, 512) == 128 and b.pick_sample_grid(1024, 512) == 512
    assert b.pick_sample_grid(0, 64) == 64  # never larger than the workload itself


def test_gmres_path_is_recognised_from_the_launch_count():
    b = load_bench()
    K = 300
    # launch counts of one restart as measured on B200 (profiles/r02_mgs_step.txt, r02_gmres_fused_projection_n1.txt)
    assert b.gmres_projection_traffic("exact_mgs", 608, K)[0] == 8.0       # persistent step kernel
    assert b.gmres_projection_traffic("exact_mgs", 45758, K)[0] == 32.0    # fused update+products launch
    assert b.gmres_projection_traffic("exact_mgs", 90608, K)[0] == 40.0    # dot + axmy
    assert b.gmres_projection_traffic("exact_mgs", 91816, K)[0] == 40.0    # ... with an all-gather kernel per projection
    assert b.gmres_projection_traffic("batched_mgs", 1506, K)[0] == 16.0
