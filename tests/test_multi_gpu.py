"""Multi-GPU parity (one process per GPU, NCCL + CUDA-IPC peer memory): runs tools/dist_check.py under
torch.distributed.run when the box has at least two GPUs, otherwise skips.  The host-side logic of the same path is
covered on CPU by tests/test_dist_cpu.py (gloo)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("ranks", [2, 4])
def test_distributed_cg_and_gmres_match_oracle(ranks):
    from algotoolkit_b200 import capi

    if capi.device_count() < ranks:
        pytest.skip(f"needs {ranks} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ranks}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + ranks), os.path.join(ROOT, "tools", "dist_check.py"), "24", "40"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0 and "DIST_CHECK PASS" in r.stdout, tail


@pytest.mark.gpu
def test_peer_protocol_stress_with_wrapping_sequence_tags():
    """200 mixed solves / applies on one distributed context with the sequence-number base seeded just below 2^32
    (tools/stress_dist.py): every repeated case must reproduce its first result bit for bit, first results match the oracle."""
    from algotoolkit_b200 import capi

    if capi.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, ATK_EPOCH_SEED="0xFFFFFF00")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tools", "stress_dist.py"), "200"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0 and "STRESS PASS" in r.stdout, (r.stdout + r.stderr)[-3000:]
