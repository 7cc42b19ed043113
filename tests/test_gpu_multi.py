"""Multi-GPU parity (needs >= 2 visible GPUs; skipped otherwise): partitioned time stepping over
NCCL against the serial oracle, launched exactly like the driver launches bench.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import dune_copasi_b200 as D
    return D.lib().dcb_device_count()


OVERLAP = "model.time_step_operator.linear_solver.b200.overlap_halo=true"


@pytest.mark.parametrize("name,mf,extra", [("grayscott3d", "1", ""), ("grayscott3d", "0", ""), ("cell3d", "1", ""),
                                           ("two_disks", "0", ""), ("gauss3d", "1", ""), ("advection3d", "0", ""),
                                           ("grayscott3d", "1", OVERLAP), ("grayscott3d_q1", "1", ""),
                                           ("grayscott3d_q1", "0", ""), ("gauss3d_q1", "1", OVERLAP)])
def test_two_rank_time_stepping_matches_serial_oracle(name, mf, extra):
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tools", "mgpu_check.py"),
           name, "2", mf, extra]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "rel L2 err" in r.stdout


def test_two_ranks_over_nccl_only():
    """DCB_PEER_COLLECTIVES=0: the same run with every collective on NCCL."""
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29518", os.path.join(ROOT, "tools", "mgpu_check.py"),
           "grayscott3d", "2", "1", ""]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT,
                       env=dict(os.environ, DCB_PEER_COLLECTIVES="0"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "peer_memory=False" in r.stdout and "rel L2 err" in r.stdout
