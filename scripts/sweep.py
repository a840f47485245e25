#!/usr/bin/env python
"""Order sweep (BASELINE config 4): operator-apply GDOF/s and HBM-roofline fraction for
3D hex p = 1..6 at a fixed dof budget, one GPU.  Prints one JSON line per order.

  python scripts/sweep.py [--dofs 5e7] [--steps 10] [--orders 1 2 3 4 5 6] [--dim 3]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import cdm_b200 as cdm  # noqa: E402
from bench import algorithmic_bytes, KAPPA, VEL, MASS, PERTURB  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dofs", type=float, default=5e7)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--orders", type=int, nargs="+", default=[1, 2, 3, 4, 5, 6])
    ap.add_argument("--dim", type=int, default=3)
    ap.add_argument("--scatter", type=int, default=1)
    ap.add_argument("--kernel", type=int, default=-1)
    args = ap.parse_args()
    peak = 6453.1
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    ctx = cdm.Context(0)
    for p in args.orders:
        n = max(1, int(round((args.dofs ** (1.0 / args.dim) - 1) / p)))
        mesh = cdm.Mesh.cartesian(ctx, args.dim, n, perturb=PERTURB)
        sp = cdm.H1Space(mesh, p)
        ess = sp.essential_dofs(np.ones(2 * args.dim, np.int32))
        op = cdm.ConvectionDiffusionOperator(sp, kappa=KAPPA, vel=VEL[:args.dim], mass=MASS, ess_dofs=ess)
        op.set_option("scatter", args.scatter)
        if args.kernel >= 0:
            op.set_option("kernel", args.kernel)
        x = torch.sin(1.0 + 0.37 * torch.arange(sp.ndof, dtype=torch.float64, device="cuda"))
        y = torch.zeros_like(x)
        torch.cuda.synchronize()
        stream = torch.cuda.ExternalStream(ctx.stream)
        for _ in range(3):
            op.Mult(x, y)
        ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            op.Mult(x, y)
        e1.record(stream)
        ctx.sync()
        ms = e0.elapsed_time(e1) / args.steps
        kms = op.time_kernel(x, y, reps=args.steps)
        nbytes = algorithmic_bytes(sp.ndof, sp.ne, p, args.dim)
        print(json.dumps({"dim": args.dim, "order": p, "n": n, "dofs": sp.ndof, "elements": sp.ne, "ms_per_apply": ms,
                          "gdofs": sp.ndof / ms / 1e6, "kernel_ms": kms, "algorithmic_GB": nbytes / 1e9,
                          "achieved_GBs": nbytes / kms / 1e6, "roofline_frac": nbytes / kms / 1e6 / peak,
                          "scatter": args.scatter}), flush=True)
        del op, sp, mesh, x, y
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
