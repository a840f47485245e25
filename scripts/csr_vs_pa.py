#!/usr/bin/env python
"""The reference's literal algorithm (assembled CSR SpMV) and the matrix-free operator on the same GPU,
same mesh, same quadrature data: kernel time per apply, bytes streamed, assembly time.

  python scripts/csr_vs_pa.py [--n 32] [--order 3]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import cdm_b200 as cdm  # noqa: E402
from bench import algorithmic_bytes, KAPPA, VEL, MASS, PERTURB  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=32)
ap.add_argument("--order", type=int, default=3)
args = ap.parse_args()
ctx = cdm.Context(0)
mesh = cdm.Mesh.cartesian(ctx, 3, args.n, perturb=PERTURB)
sp = cdm.H1Space(mesh, args.order)
ess = sp.essential_dofs(np.ones(6, np.int32))
op = cdm.ConvectionDiffusionOperator(sp, kappa=KAPPA, vel=VEL, mass=MASS, ess_dofs=ess)
x = torch.sin(1.0 + 0.37 * torch.arange(sp.ndof, dtype=torch.float64, device="cuda"))
y = torch.zeros_like(x)
torch.cuda.synchronize()
pa_ms = op.time_kernel(x, y, reps=20)
op.Mult(x, y)
ctx.sync()
ypa = y.clone()
t0 = time.perf_counter()
rowptr, colind, vals = op.assemble_csr()
t_asm = time.perf_counter() - t0
op.set_option("assembly", 1)
csr_ms = op.time_kernel(x, y, reps=20)
op.Mult(x, y)
ctx.sync()
diff = float((y - ypa).norm() / ypa.norm())
nnz = len(colind)
csr_bytes = 12 * nnz + 8 * (sp.ndof + 1) + 16 * sp.ndof
print(json.dumps({"order": args.order, "n": args.n, "dofs": sp.ndof, "nnz": nnz, "nnz_per_row": nnz / sp.ndof,
                  "assemble_s_incl_host_pattern": round(t_asm, 3),
                  "pa_kernel_ms": pa_ms, "pa_GB": algorithmic_bytes(sp.ndof, sp.ne, args.order) / 1e9,
                  "csr_spmv_ms": csr_ms, "csr_GB": csr_bytes / 1e9, "csr_GBs": csr_bytes / csr_ms / 1e6,
                  "spmv_over_pa": csr_ms / pa_ms, "rel_diff_pa_vs_csr": diff}))
