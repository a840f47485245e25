"""Probe of the host-vector (e2e) path on config 2: PCIe copy times, plain vs pipelined
cdm_operator_mult_host (set CDM_PIPE_DEBUG=1 for the library's own stage timings)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import cdm_b200 as cdm  # noqa: E402

ctx = cdm.Context(0)
mesh = cdm.Mesh.cartesian(ctx, 3, 66, perturb=0.1)
sp = cdm.H1Space(mesh, 3)
ess = sp.essential_dofs(np.ones(6, np.int32))
op = cdm.ConvectionDiffusionOperator(sp, kappa=0.1, vel=(1.0, -2.0, 0.5), mass=1.0, ess_dofs=ess)
n = sp.ndof
xh = torch.empty(n, dtype=torch.float64).pin_memory()
yh = torch.empty(n, dtype=torch.float64).pin_memory()
xh.copy_(torch.sin(torch.arange(n, dtype=torch.float64)))
xd = torch.empty(n, dtype=torch.float64, device="cuda")
yd = torch.empty(n, dtype=torch.float64, device="cuda")


def t(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


print("H2D 63MB ms", t(lambda: xd.copy_(xh, non_blocking=True)))
print("D2H 63MB ms", t(lambda: yh.copy_(yd, non_blocking=True)))
x, y = xh.numpy(), yh.numpy()
ref = None
op.set_option("host_pipeline", 0)
print("mult_host un-pipelined ms", t(lambda: op.mult_host(x, y)))
ref = y.copy()
for shape in (0, 1):                       # 0: round-1 schedule (equal chunks, 4 copies per chunk and direction), 1: tapered + merged
    for K in (4, 6, 8, 10, 12, 16):
        opk = cdm.ConvectionDiffusionOperator(sp, kappa=0.1, vel=(1.0, -2.0, 0.5), mass=1.0, ess_dofs=ess)
        opk.set_option("host_pipeline_shape", shape)
        opk.set_option("host_pipeline", K)
        ms = t(lambda: opk.mult_host(x, y))
        print("mult_host shape=%d K=%d ms %.4f  rel diff vs un-pipelined %.2e" % (shape, K, ms, np.linalg.norm(y - ref) / np.linalg.norm(ref)))
        del opk
for cap in (0, 444, 296, 148, 74):            # does the HBM-saturating element kernel starve the copy engines?
    opk = cdm.ConvectionDiffusionOperator(sp, kappa=0.1, vel=(1.0, -2.0, 0.5), mass=1.0, ess_dofs=ess)
    opk.set_option("host_pipeline", 6)
    opk.set_option("grid_cap", cap)
    print("mult_host shape=1 K=6 grid_cap=%d ms %.4f" % (cap, t(lambda: opk.mult_host(x, y))))
    del opk
op.set_option("host_pipeline", 1)
print("mult_host default ms", t(lambda: op.mult_host(x, y)), "rel diff", np.linalg.norm(y - ref) / np.linalg.norm(ref))
