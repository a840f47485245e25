#!/usr/bin/env python
"""Launch every kernel of the path once inside an NVTX range ("prof"), after a warm-up outside it, so that
    ncu --nvtx --nvtx-include "prof/" --set full ... python scripts/profile_targets.py
captures exactly one launch of each: the apply kernels for every order (default kernel per order), the 2-D kernel,
quadrature-data setup, the Jacobi diagonal, the Krylov vector kernels (10 GMRES steps) and the CSR SpMV.
Sizes ~8 M dofs (config-2 scale) unless --small.  Without ncu it just prints the CUDA-event time of each target.
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--only", default="", help="comma-separated target names")
    ap.add_argument("--gmres-steps", type=int, default=4, help="GMRES steps inside the profiled range (0: skip)")
    args = ap.parse_args()
    import torch
    cdm = importlib.import_module("continuum-mechanics-mfem_b200")
    ctx = cdm.Context(0)
    stream = torch.cuda.ExternalStream(ctx.stream)
    nvtx = torch.cuda.nvtx
    only = set(filter(None, args.only.split(",")))
    out = []

    def timed(name, fn, warm=2):
        if only and name not in only:
            return
        for _ in range(warm):
            fn()
        ctx.sync(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nvtx.range_push("prof")
        e0.record(stream)
        fn()
        e1.record(stream)
        ctx.sync(); torch.cuda.synchronize()
        nvtx.range_pop()
        out.append({"target": name, "ms": e0.elapsed_time(e1)})
        print(json.dumps(out[-1]), flush=True)

    def make(dim, p, n, **kw):
        mesh = cdm.Mesh.cartesian(ctx, dim, n, perturb=0.1)
        sp = cdm.H1Space(mesh, p)
        ess = sp.essential_dofs(np.ones(2 * dim, np.int32))
        vel = (1.0, -2.0, 0.5)[:dim]
        op = cdm.ConvectionDiffusionOperator(sp, kappa=0.1, vel=vel, mass=1.0, ess_dofs=ess, **kw)
        x = torch.sin(1.0 + 0.37 * torch.arange(sp.ndof, dtype=torch.float64, device="cuda"))
        y = torch.zeros_like(x)
        torch.cuda.synchronize()
        return mesh, sp, op, x, y

    sc = 0.5 if args.small else 1.0
    sizes3 = {1: int(199 * sc), 2: int(100 * sc), 3: int(66 * sc), 4: int(50 * sc), 5: int(40 * sc), 6: int(33 * sc)}
    for p in (3, 1, 2, 4, 5, 6):
        mesh, sp, op, x, y = make(3, p, sizes3[p])
        timed(f"apply3d_p{p}", lambda: op.Mult(x, y))
        if p == 3:
            op.set_option("scatter", 0)
            timed("apply3d_p3_deterministic", lambda: op.Mult(x, y))
            op.set_option("scatter", 1)
            d = torch.zeros_like(x)
            timed("setup_qdata_p3", lambda: op.update(kappa=0.1, vel=(1.0, -2.0, 0.5), mass=1.0))
            timed("diag_p3", lambda: op.AssembleDiagonal(d))
            b = torch.sin(0.5 + 0.11 * torch.arange(sp.ndof, dtype=torch.float64, device="cuda"))
            xs = torch.zeros_like(b)
            if args.gmres_steps > 0:
                s = cdm.GMRESSolver(cdm.GMRES_PETSC, args.gmres_steps, args.gmres_steps, 0.0, 0.0, jacobi=True)
                s.SetOperator(op)
                timed(f"gmres{args.gmres_steps}_p3", lambda: s.Mult(b, xs), warm=1)
                del s
            del b, xs, d
        del op, sp, mesh, x, y
        torch.cuda.empty_cache()
    for p in (2, 3):
        n2 = int((2800 if p == 1 else 1400 if p == 2 else 930) * sc)
        mesh, sp, op, x, y = make(2, p, n2)
        timed(f"apply2d_p{p}", lambda: op.Mult(x, y))
        del op, sp, mesh, x, y
        torch.cuda.empty_cache()
    mesh, sp, op, x, y = make(3, 3, int(32 * sc) if not args.small else 12)
    op.set_option("assembly", 1)
    timed("csr_spmv_p3", lambda: op.Mult(x, y))
    print(json.dumps({"targets": out}))


if __name__ == "__main__":
    main()
