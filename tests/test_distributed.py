"""Multi-rank tests of the mesh-partitioned mode (ParMesh / ParFiniteElementSpace semantics,
linear_convection_diffusion_2D.cpp:300,312).  CPU: world_size 2 and 4 over gloo (partition plan,
P / P^T semantics; element work by the oracle).  GPU: NCCL, when >= 2 devices are visible."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(nproc, mode, order, n=(6, 5, 4), extra=()):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "dist_check.py"), "--mode", mode, "--order", str(order), "--mesh", *map(str, n), *extra]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


@pytest.mark.parametrize("nproc,order,mesh", [(2, 3, (6, 5, 4)), (4, 2, (6, 5, 4)), (8, 2, (4, 4, 4))])
def test_partitioned_apply_gloo(nproc, order, mesh):
    """2x1x1, 2x2x1 and 2x2x2 boxes (the last one has face, edge and corner neighbours: 7 peers per rank)"""
    out = _run(nproc, "cpu", order, n=mesh)
    assert out.count("host-emulated partitioned apply") == nproc
    assert out.count("host-emulated symmetric exchange") == nproc


@pytest.mark.parametrize("nproc,order,mesh,kind", [(2, 3, (5, 4, 3), "checker"), (4, 2, (6, 5, 4), "random"), (3, 1, (4, 4, 6), "slabs"),
                                                   (8, 2, (4, 4, 4), "checker")])
def test_elementwise_partition_gloo(nproc, order, mesh, kind):
    """cdm_mesh_partition_elements (a per-element rank array as METIS would give): the plans of the space -- owners, P / P^T
    peer lists, symmetric exchange -- reproduce the un-partitioned oracle apply on every rank; "checker" makes every element face
    a partition boundary (dofs shared by up to 8 ranks), "random" gives irregular sharing groups"""
    out = _run(nproc, "cpu", order, n=mesh, extra=("--partition", kind))
    assert out.count("host-emulated partitioned apply") == nproc
    assert out.count("host-emulated symmetric exchange") == nproc


@pytest.mark.gpu
def test_partitioned_apply_and_gmres_nccl():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run through gpurun --gpus 2)")
    nproc = 8 if n >= 8 else (4 if n >= 4 else 2)
    out = _run(nproc, "gpu", 3, n=(8, 6, 6), extra=("--p2p",))          # default protocol: symmetric peer-memory exchange
    assert out.count("cg iters") == nproc and out.count("chained apply without P") == 2 * nproc
    assert out.count("backward-Euler step") == 3 * nproc
    out = _run(nproc, "gpu", 3, n=(8, 6, 6), extra=("--halo", "0"))     # P / P^T over NCCL
    assert out.count("cg iters") == nproc
