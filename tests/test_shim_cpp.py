"""The MFEM-shaped C++ shim (include/cdm_mfem_shim.hpp) and the re-hosted steady driver
(examples/convdiff_steady.cpp, mirroring linear_convection_diffusion_2D.cpp:238-446)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "convdiff_steady")


def _build():
    import cdm_b200 as cdm
    if not os.path.exists(cdm.LIB_PATH):
        cdm.build()
    subprocess.run(["make", "-C", os.path.join(ROOT, "examples")], check=True, capture_output=True)


def test_example_builds_and_fails_loudly_without_gpu():
    _build()
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([EXE, "2", "4", "2"], capture_output=True, text=True)
    assert r.returncode == 3                      # the app's runtime-failure exit code (:441)
    assert "no usable CUDA device" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("dim,n,order,tol", [(2, 16, 2, 2e-3), (3, 12, 3, 2e-3), (2, 8, 3, 5e-3)])
def test_steady_driver_converges_to_manufactured_solution(dim, n, order, tol):
    _build()
    r = subprocess.run([EXE, str(dim), str(n), str(order)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    its = int(re.search(r"GMRES iterations: (\d+)", r.stdout).group(1))
    err = float(re.search(r"L2 error: abs [0-9.e+-]+\s+rel ([0-9.e+-]+)", r.stdout).group(1))
    assert 0 < its < 500 and err < tol, r.stdout
