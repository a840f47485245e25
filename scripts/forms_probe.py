"""Timing probe (config 2: 66^3 hex, p=3) of the steps either side of the solve: rule points, linear
form, L2 error, boundary projection and the per-step quadrature-data update with device-resident
per-point coefficients.  CUDA events on the library's stream, inputs resident on the device."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import cdm_b200 as cdm  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 66
p = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = cdm.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
mesh = cdm.Mesh.cartesian(ctx, 3, n, perturb=0.1)
sp = cdm.H1Space(mesh, p)
ess = sp.essential_dofs(np.ones(6, np.int32))
op = cdm.ConvectionDiffusionOperator(sp, kappa=0.1, vel=(1.0, -2.0, 0.5), mass=1.0, ess_dofs=ess)
N, ne = sp.ndof, sp.ne


def timed(fn, reps=10):
    fn()
    ctx.sync()
    with torch.cuda.stream(stream):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
    ctx.sync()
    return e0.elapsed_time(e1) / reps


out = {"n": n, "order": p, "dofs": N, "elements": ne}
q_lf, q_err = p + 1, p + 2
x_lf = torch.zeros((ne, q_lf ** 3, 3), dtype=torch.float64, device="cuda")
x_err = torch.zeros((ne, q_err ** 3, 3), dtype=torch.float64, device="cuda")
out["rule_coords_lf_ms"] = timed(lambda: sp.rule_coords(q_lf, out=x_lf))
out["rule_coords_err_ms"] = timed(lambda: sp.rule_coords(q_err, out=x_err))
f = (1.0 + torch.sin(2.3 * x_lf[..., 0]) * torch.cos(1.7 * x_lf[..., 1])).contiguous()
uex = (torch.sin(3.0 * x_err[..., 0]) * x_err[..., 2]).contiguous()
b = torch.zeros(N, dtype=torch.float64, device="cuda")
u = torch.rand(N, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
out["domain_lf_ms"] = timed(lambda: sp.domain_lf(f, b))
out["domain_lf_gbs"] = (8 * f.numel() + 4 * ne * (p + 1) ** 3 + 16 * N) / out["domain_lf_ms"] / 1e6
out["l2_error_ms"] = timed(lambda: sp.l2_error(u, uex))
out["l2_error_gbs"] = (8 * uex.numel() + 4 * ne * (p + 1) ** 3 + 8 * N) / out["l2_error_ms"] / 1e6
essd = torch.from_numpy(ess).cuda()
g = torch.rand(len(ess), dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
out["project_bdr_ms"] = timed(lambda: sp.project_dofs(essd, g, u))
# per-step D re-setup: constant coefficients vs device-resident per-point arrays (10 values per point)
nq = sp.nq
out["qdata_const_ms"] = timed(lambda: op.update(kappa=0.1, vel=(1.0, -2.0, 0.5), mass=1.0), reps=5)
kap = torch.rand((ne, nq, 6), dtype=torch.float64, device="cuda") + 1.0
vel = torch.rand((ne, nq, 3), dtype=torch.float64, device="cuda")
mas = torch.rand((ne, nq), dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
out["qdata_device_coeff_ms"] = timed(lambda: op.update(kappa=kap, vel=vel, alpha=-1.0, mass=mas), reps=5)
out["qdata_device_coeff_gbs"] = (2 * 8 * 10 * ne * nq + 8 * 24 * ne) / out["qdata_device_coeff_ms"] / 1e6
print(json.dumps(out))
