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
for K in (2, 4, 6, 12, 16, 24):          # chunk-count sweep: one operator per K (the plan is built once per operator)
    opk = cdm.ConvectionDiffusionOperator(sp, kappa=0.1, vel=(1.0, -2.0, 0.5), mass=1.0, ess_dofs=ess)
    opk.set_option("host_pipeline", K)
    print("mult_host K=%d ms" % K, t(lambda: opk.mult_host(x, y)))
    del opk
ref = None
for mode in (0, 1):
    op.set_option("host_pipeline", mode)
    print("mult_host pipeline=%d ms" % mode, t(lambda: op.mult_host(x, y)))
    if ref is None:
        ref = y.copy()
    else:
        print("pipelined vs plain rel diff", np.linalg.norm(y - ref) / np.linalg.norm(ref))
