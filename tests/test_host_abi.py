"""CPU tests of the product's host side: the C-ABI library loads and exports every
symbol include/cdm_b200.h declares, host logic (mesh, numbering, restriction maps,
essential dofs, partition plan) is bit-exact against the oracle, and compute entry
points fail loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import cdm_b200 as cdm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hctx():
    if not os.path.exists(cdm.LIB_PATH):
        cdm.build()
    return cdm.Context(host_only=True)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(cdm.LIB_PATH):
        cdm.build()
    hdr = open(os.path.join(ROOT, "include", "cdm_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(cdm_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 45
    L = C.CDLL(cdm.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert declared == set(cdm.SIGNATURES), declared ^ set(cdm.SIGNATURES)
    L.cdm_version.restype = C.c_char_p
    assert b"sm_100a" in L.cdm_version()


def test_compute_fails_loudly_without_gpu(hctx):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(cdm.CdmError) as ei:
        cdm.Context(0)
    assert ei.value.code == cdm.ENOGPU
    m = cdm.Mesh.cartesian(hctx, 2, 2)
    sp = cdm.H1Space(m, 1)
    with pytest.raises(cdm.CdmError) as ei:
        cdm.ConvectionDiffusionOperator(sp, kappa=1.0)
    assert ei.value.code == cdm.ENOGPU
    # the form kernels (linear form, L2 error, rule points, projection) have no CPU path either
    b = np.zeros(sp.ndof)
    for call in (lambda: sp.rule_coords(2),
                 lambda: sp.domain_lf(np.ones((sp.ne, 4)), b),
                 lambda: sp.l2_error(None, np.ones((sp.ne, 9))),
                 lambda: sp.project_dofs(np.zeros(1, np.int32), np.zeros(1), b)):
        with pytest.raises(cdm.CdmError) as ei:
            call()
        assert ei.value.code == cdm.ENOGPU


def test_rule_points_follow_mfem_intrules():
    """IntRules.Get(geom, order) on segments / squares / cubes: order/2 + 1 Gauss-Legendre points per direction"""
    L = cdm.lib()
    assert [L.cdm_rule_points(o) for o in (0, 1, 2, 3, 4, 5, 8, 9)] == [1, 1, 2, 2, 3, 3, 5, 5]
    for p in range(1, 7):
        assert L.cdm_rule_points(2 * p) == p + 1                      # DomainLFIntegrator default (order 2p)
        assert L.cdm_rule_points(max(2, 2 * p + 3)) == p + 2          # the app's error rule (:383)
        assert L.cdm_rule_points(2 * p + 3 - 1) == p + 2              # Diffusion/Mass/Convection rule in 3D (2p + dim - 1)
    assert L.cdm_rule_points(-1) == cdm.EINVAL


def test_bad_arguments_are_errors_not_crashes(hctx):
    with pytest.raises(cdm.CdmError):
        cdm.Mesh.cartesian(hctx, 4, 2)
    with pytest.raises(cdm.CdmError):
        cdm.Mesh.cartesian(hctx, 2, 0)
    m = cdm.Mesh.cartesian(hctx, 2, 2)
    with pytest.raises(cdm.CdmError) as ei:
        cdm.H1Space(m, 9)
    assert ei.value.code == cdm.EUNSUP
    with pytest.raises(cdm.CdmError):
        cdm.Mesh.from_arrays(hctx, np.zeros((4, 2)), np.array([[0, 1, 2, 7]]), np.zeros((0, 2)), np.zeros(0))


@pytest.mark.parametrize("dim,n", [(2, (3, 5)), (3, (2, 3, 4))])
def test_cartesian_mesh_matches_oracle(hctx, orc, dim, n):
    m = cdm.Mesh.cartesian(hctx, dim, n, perturb=0.1)
    vx, ev, bv, ba = m.arrays()
    ovx, oev, obv, oba = orc.cart_mesh(dim, list(n), None, 0.1)
    assert np.array_equal(ev, oev)
    assert np.abs(vx - ovx).max() < 1e-15
    key = lambda b, a: sorted(map(tuple, np.c_[np.sort(b, 1), a]))
    assert key(bv, ba) == key(obv, oba)


@pytest.mark.parametrize("dim,p,seed", [(2, 1, None), (2, 2, None), (2, 3, 4), (2, 6, 1), (3, 1, None), (3, 2, 8),
                                        (3, 3, None), (3, 3, 21), (3, 4, 5), (3, 6, None)])
def test_index_maps_bit_exact(hctx, orc, dim, p, seed):
    """element->dof table, ElementRestriction offsets/indices and the essential dof list
    must be identical (int32, bit for bit) to the oracle's."""
    P = orc.Problem(dim, p, 3 if p < 6 else 2, perturb=0.1, shuffle_seed=seed)
    m = cdm.Mesh.from_arrays(hctx, P.vx, P.ev, P.bv, P.battr)
    sp = cdm.H1Space(m, p)
    g, o, i = sp.maps()
    assert sp.ndof == P.ndof and sp.q1d == P.q1d
    assert g.dtype == np.int32 and np.array_equal(g, P.elem_dof)
    assert np.array_equal(o, P.offsets) and np.array_equal(i, P.indices)
    assert np.array_equal(sp.essential_dofs(P.marker), P.ess)
    B, G, qw, _ = sp.basis()
    Bo, Go, wo = orc.basis(p, sp.q1d)
    assert np.abs(B - Bo).max() < 1e-14 and np.abs(G - Go).max() < 2e-13 and np.abs(qw - wo).max() < 1e-15
    assert np.abs(sp.dof_coords() - P.coords()).max() < 1e-14


def test_essential_subset(hctx, orc):
    P = orc.Problem(3, 2, 3, ess_attrs=[3, 5])
    m = cdm.Mesh.from_arrays(hctx, P.vx, P.ev, P.bv, P.battr)
    sp = cdm.H1Space(m, 2)
    assert np.array_equal(sp.essential_dofs(P.marker), P.ess)


def test_space_from_table(hctx, orc):
    """the element->dof table may come from the caller (a live MFEM space): maps derive from it"""
    P = orc.Problem(3, 2, 2, shuffle_seed=3)
    rng = np.random.default_rng(0)
    perm = rng.permutation(P.ndof).astype(np.int32)
    table = perm[P.elem_dof]
    m = cdm.Mesh.from_arrays(hctx, P.vx, P.ev, P.bv, P.battr)
    sp = cdm.H1Space(m, 2, elem_dof=table, ndof=P.ndof)
    g, o, i = sp.maps()
    oo, ii = orc.restriction(table, P.ndof)
    assert np.array_equal(g, table) and np.array_equal(o, oo) and np.array_equal(i, ii)
    assert np.array_equal(sp.essential_dofs(P.marker), np.sort(perm[P.ess]))


@pytest.mark.parametrize("dim,p,parts,n", [(2, 2, (2, 2), (5, 4)), (3, 1, (2, 1, 1), (4, 3, 3)), (3, 2, (2, 2, 2), (5, 4, 4)),
                                           (3, 3, (2, 2, 1), (4, 5, 2)), (3, 2, (3, 1, 2), (7, 2, 4))])
def test_partition_plan_is_consistent(hctx, dim, p, parts, n):
    """Box partition: every global dof has exactly one owner, sharers agree on the exchange
    order, and T-dofs sum to the global dof count (ParFiniteElementSpace semantics)."""
    g = cdm.Mesh.cartesian(hctx, dim, n, perturb=0.1)
    gs = cdm.H1Space(g, p)
    nr = int(np.prod(parts))
    import importlib
    capi = importlib.import_module("continuum-mechanics-mfem_b200.capi")
    spaces = []
    for r in range(nr):
        lm = g.partition_box(parts, r)
        spaces.append((lm, cdm.H1Space(lm, p)))
    assert sum(s.ntrue for _, s in spaces) == gs.ndof
    assert sum(s.ne for _, s in spaces) == gs.ne
    # coordinates of owned dofs over all ranks = coordinates of the global dofs
    allx = np.concatenate([s.dof_coords()[:s.ntrue] for _, s in spaces])
    gx = gs.dof_coords()
    assert len(np.unique(np.round(allx, 10), axis=0)) == gs.ndof
    assert {tuple(v) for v in np.round(allx, 10)} == {tuple(v) for v in np.round(gx, 10)}
    # exchange plan: what rank a owns-and-shares with b is exactly what b holds as ghosts of a,
    # in the same order (global lattice keys), and every ghost has exactly one owner peer
    plans = [s.halo() for _, s in spaces]
    keys = [s.dof_global() for _, s in spaces]
    for a in range(nr):
        ghosts_seen = []
        for peer, own, ghost in plans[a]:
            back = [q for q in plans[peer] if q[0] == a]
            assert len(back) == 1
            _, own_b, ghost_b = back[0]
            assert np.array_equal(keys[a][own], keys[peer][ghost_b])
            assert np.array_equal(keys[a][ghost], keys[peer][own_b])
            assert np.all(own < spaces[a][1].ntrue) and np.all(ghost >= spaces[a][1].ntrue)
            ghosts_seen.append(ghost)
        gs_all = np.concatenate(ghosts_seen) if ghosts_seen else np.zeros(0, np.int32)
        assert np.array_equal(np.sort(gs_all), np.arange(spaces[a][1].ntrue, spaces[a][1].ndof))
