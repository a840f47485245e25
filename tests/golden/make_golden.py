"""Generates tests/golden/oracle_small.json from the CPU oracle.

The reference (MFEM/PETSc application) cannot be built or imported in this
environment and commits no golden vectors, so these fixtures pin the oracle's
*own* output: they guard against drift of the oracle and give the GPU tests a
size-independent anchor.  Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as O  # noqa: E402

cases = []
for dim, p, n, perturb in [(2, 2, 4, 0.0), (2, 3, 3, 0.1), (3, 1, 3, 0.1), (3, 2, 2, 0.1), (3, 3, 2, 0.1), (3, 4, 2, 0.12)]:
    P = O.Problem(dim, p, n, perturb=perturb)
    x = np.sin(1.0 + 0.37 * np.arange(P.ndof))
    y = P.pa_apply(x)
    # linear form with MFEM's default rule (p+1 points) and L2 error with the app's rule (p+2 points) for
    # f(x) = 1 + sin(2.3 x0) cos(1.7 x1) + 0.5 x_last^2
    fn = lambda c: 1.0 + np.sin(2.3 * c[..., 0]) * np.cos(1.7 * c[..., 1]) + 0.5 * c[..., -1] ** 2
    lf = P.domain_lf(fn(P.rule_coords(p + 1)))
    cases.append(dict(dim=dim, p=p, n=n, perturb=perturb, ndof=P.ndof,
                      elem_dof_head=[int(v) for v in P.elem_dof[:2].reshape(-1)],
                      y_norm=float(np.linalg.norm(y)), y_head=[float(v) for v in y[:8]],
                      diag_sum=float(P.pa_diag().sum()),
                      lf_norm=float(np.linalg.norm(lf)), lf_head=[float(v) for v in lf[:8]],
                      l2_error=float(P.l2_error(x, fn(P.rule_coords(p + 2))))))
out = dict(note="config 1 of BASELINE.json is the first case (2D, order 2, 4x4 quads, 81 dofs); "
                "kappa=0.1 c=(1,-2[,0.5]) s=1 (Input/input_2d.yaml:7-10); x_i = sin(1+0.37 i)",
           cases=cases)
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_small.json"), "w") as fh:
    json.dump(out, fh, indent=1)
print("wrote", len(cases), "cases")
