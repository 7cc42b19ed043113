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
                                           ("grayscott3d_q1", "0", ""), ("gauss3d_q1", "1", OVERLAP),
                                           ("cell10_nested", "1", ""),   # unstructured tets, RCB partition, general halo plan
                                           ("tables", "1", "")])
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


TILE = "model.assembly.b200.tile=true,model.assembly.b200.tile_w=2,model.assembly.b200.tile_r=1,model.assembly.b200.tile_lz=2"


def _run(world, name, mf, extra, port, env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "mgpu_check.py"),
           name, "2", mf, extra]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert f"world={world}" in r.stdout and "rel L2 err" in r.stdout
    return r.stdout


@pytest.mark.parametrize("world", [3, 4, 8])
@pytest.mark.parametrize("name,mf,extra", [("grayscott3d", "1", ""),            # 11 vertex planes: slabs with remainder planes
                                           ("grayscott3d_aniso", "1", ""),      # 9 planes along the slab axis, anisotropic
                                           ("grayscott3d_aniso", "0", ""),      # assembled (CSR + SpMV)
                                           ("cell3d", "1", ""),                 # three compartments, general halo plan
                                           ("grayscott3d_q1", "1", ""),
                                           ("cell10_nested", "1", ""),          # BASELINE configs[4] in miniature: RCB over unstructured tets
                                           ("grayscott3d_wide", "1", TILE)])    # tile drivers + fused BiCGSTAB
def test_many_rank_time_stepping_matches_serial_oracle(world, name, mf, extra):
    """3 and 4 ranks cut the lattices of the suite into slabs of unequal thickness (down to one owned plane
    per rank at 8), exercising halo plans with remainder planes; all against the serial oracle <= 1e-10."""
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    _run(world, name, mf, extra, 29520 + world)


@pytest.mark.parametrize("name,extra", [("grayscott3d_aniso", TILE), ("grayscott3d_wide", TILE), ("gauss3d_aniso", TILE),
                                        ("grayscott3d_q1", TILE)])
def test_two_rank_tile_drivers_and_fused_bicgstab(name, extra):
    """The tile-marching kernels on slabs: the fused BiCGSTAB updates the ghost planes of r and p locally and
    exchanges v and t instead (one halo update per operator application, as the unfused loop)."""
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2, name, "1", extra, 29519)
