"""B200-native convection-diffusion hot path.

The directory name is not a Python identifier; load it with
``importlib.import_module("continuum-mechanics-mfem_b200")`` (what
``__graft_entry__`` and ``cdm_b200.py`` at the repo root do).
"""
from .capi import (  # noqa: F401
    CdmError, Context, Mesh, H1Space, ConvectionDiffusionOperator, GMRESSolver, CGSolver, Config,
    Coeff, KrylovOpts, KrylovResult, SIGNATURES, LIB_PATH, build, lib,
    GMRES_PETSC, GMRES_MFEM, COEFF_NONE, COEFF_CONST, COEFF_QPT,
    OK, EINVAL, ENOGPU, ECUDA, ENOMEM, ENCCL, EUNSUP,
)
