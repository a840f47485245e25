"""Data formats either side of the path and the host side of the triangle path (rank 3): Gmsh 2.2
reader, Hilbert-curve element order, YAML / PETSc-options loaders, ParaView writer, triangle H1 numbering against an
independent numpy restatement (oracle/tri_oracle.py), the oracle's ILU(0) against its defining property.  CPU only."""
import os

import numpy as np
import pytest

import cdm_b200 as cdm

GOLD = os.path.join(os.path.dirname(__file__), "golden")
REF = "/root/reference/myapps/convection_diffusion"


@pytest.fixture(scope="module")
def ctx():
    return cdm.Context(host_only=True)


def test_gmsh_fixture_meshes(ctx):
    m = cdm.Mesh.read_gmsh(ctx, os.path.join(GOLD, "square_tri.msh"))
    vx, ev, bv, ba = m.arrays()
    assert m.dim == 2 and m.geometry == (1, 3, 2) and m.nv == 100 and m.nbe == 36 and sorted(set(ba.tolist())) == [1, 2, 3, 4]
    a, b, c = vx[ev[:, 0]], vx[ev[:, 1]], vx[ev[:, 2]]
    area = 0.5 * ((b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0]))
    assert area.min() > 0 and abs(area.sum() - 1.0) < 1e-14                        # counter-clockwise, covers the square
    ln = lambda p, q: np.hypot(*(p - q).T)
    assert np.all(ln(a, b) >= ln(b, c)) and np.all(ln(a, b) >= ln(c, a))           # refine=1: longest edge first
    m0 = cdm.Mesh.read_gmsh(ctx, os.path.join(GOLD, "square_tri.msh"), refine=False)
    assert np.array_equal(np.sort(m0.arrays()[1], axis=1), np.sort(ev, axis=1))    # same triangles, rotated only
    # boundary attributes sit on the sides the physical names say (Mesh/unit_square.geo:18-21)
    for attr, (axis, val) in {1: (1, 0.0), 2: (0, 1.0), 3: (1, 1.0), 4: (0, 0.0)}.items():
        assert np.allclose(vx[bv[ba == attr]][:, :, axis], val)
    d = cdm.Mesh.read_gmsh(ctx, os.path.join(GOLD, "disk_tri.msh"))
    assert np.allclose(np.hypot(*d.arrays()[0][d.arrays()[2]].reshape(-1, 2).T), 1.0)
    with pytest.raises(cdm.CdmError):
        cdm.Mesh.read_gmsh(ctx, os.path.join(GOLD, "oracle_small.json"))


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is only present in the build container")
def test_reference_meshes_and_inputs_load_unmodified(ctx):
    """every mesh and every flat input file the reference ships"""
    want = {"unit_square": (1, 510, 938, 80), "unit_circle": (1, 1593, 3056, 128), "square_0p01": (1, 515, 948, 80),
            "ablation_strip": (0, 2880, 2629, 500), "ablation_strip_tri_uniform": (1, 917, 1676, 156)}
    for name, (geom, nv, ne, nbe) in want.items():
        m = cdm.Mesh.read_gmsh(ctx, f"{REF}/Mesh/{name}.msh")
        assert (m.geometry[0], m.nv, m.ne, m.nbe) == (geom, nv, ne, nbe)
    sp = cdm.H1Space(cdm.Mesh.read_gmsh(ctx, f"{REF}/Mesh/unit_square.msh"), 3)     # Input/input_2d.yaml: order 3
    assert sp.ndof == 4342 and len(sp.essential_dofs(np.ones(4, np.int32))) == 240
    cfg = cdm.Config(f"{REF}/Input/input_2d.yaml")
    assert cfg.get("mesh_file") == "Mesh/unit_square.msh" and cfg.get("order", 1, int) == 3 and cfg.get("kappa", 0.0, float) == 0.1
    assert cfg.get("cy", 0.0, float) == -2.0 and cfg.get("save_paraview", False, bool) is True and cfg.get("absent", "dflt") == "dflt"
    assert cdm.Config(f"{REF}/Input/input.yaml").get("peclet", None, list) == [1.0, 10.0, 100.0]
    o = cdm.GMRESSolver.from_petsc_options(f"{REF}/Input/petsc.opts").opts
    assert (o.rtol, o.atol, o.max_it, o.restart, o.jacobi) == (1e-10, 1e-12, 500, 30, 1)
    o = cdm.GMRESSolver.from_petsc_options(f"{REF}/Input/petsc_circle.opts").opts
    assert (o.rtol, o.atol, o.max_it, o.jacobi) == (1e-10, 1e-12, 2000, 2)


def test_yaml_and_petsc_options_loaders(tmp_path):
    y = tmp_path / "in.yaml"
    y.write_text("# comment\nmesh_file: Mesh/a b.msh   # trailing\norder: 3\nkappa: 1.0e-1\ncy: -2.0\nname: \"q # not a comment\"\n"
                 "peclet: [1.0, 10.0, 100.0]\nsave_paraview: false\n---\n")
    c = cdm.Config(y)
    assert c.get("mesh_file") == "Mesh/a b.msh" and c.get("order", 0, int) == 3 and c.get("kappa", 0.0, float) == 0.1
    assert c.get("name") == "q # not a comment" and c.get("peclet", None, list) == [1.0, 10.0, 100.0]
    assert c.get("save_paraview", True, bool) is False and not c.has("nokey")
    with pytest.raises(cdm.CdmError):
        c.get("mesh_file", 0, int)
    (tmp_path / "nested.yaml").write_text("a:\n  b: 1\n")
    with pytest.raises(cdm.CdmError):
        cdm.Config(tmp_path / "nested.yaml")
    p = tmp_path / "petsc.opts"
    p.write_text("# opts\n-ksp_type gmres\n-ksp_rtol 1.0e-10 -ksp_atol 1.0e-12\n-ksp_max_it 2000\n-pc_type bjacobi\n-sub_ksp_type preonly\n"
                 "-sub_pc_type ilu\n-ksp_monitor\n-ksp_gmres_restart 50\n")
    s = cdm.GMRESSolver.from_petsc_options(p)
    assert s.ksp_type == "gmres" and (s.opts.rtol, s.opts.atol, s.opts.max_it, s.opts.restart, s.opts.jacobi) == (1e-10, 1e-12, 2000, 50, 2)
    p.write_text("-ksp_type cg\n-pc_type none\n")
    s = cdm.GMRESSolver.from_petsc_options(p)
    assert s.ksp_type == "cg" and s.opts.jacobi == 0 and s.opts.rtol == 1e-5            # KSP default rtol
    p.write_text("-pc_type gamg\n")
    with pytest.raises(cdm.CdmError):
        cdm.GMRESSolver.from_petsc_options(p)


@pytest.mark.parametrize("n", [[4, 4], [5, 3], [7, 7], [16, 9], [2, 2, 2], [8, 8, 8], [3, 4, 5], [6, 5, 7]])
def test_hilbert_curve_ordering(ctx, n):
    """NCMesh::GridSfcOrdering: a permutation of the cells whose consecutive cells are neighbours (a few diagonal steps
    appear on odd-sized 3-D grids, as in the generalised Hilbert curve itself); the mesh lists its elements along it"""
    import ctypes as C
    dim = len(n)
    co = np.zeros((int(np.prod(n)), dim), np.int64)
    assert cdm.lib().cdm_grid_sfc_ordering(dim, (C.c_int64 * 3)(*(n + [1] * (3 - dim))), co.ctypes.data_as(C.c_void_p)) == 0
    lin = co[:, 0] + n[0] * (co[:, 1] + (n[1] * co[:, 2] if dim == 3 else 0))
    assert np.array_equal(np.sort(lin), np.arange(len(lin))) and np.all(co[0] == 0)
    step = np.abs(np.diff(co, axis=0))
    assert step.max() == 1 and (step.sum(1) > 1).sum() <= (0 if dim == 2 else len(lin) // 20)
    m, ms = cdm.Mesh.cartesian(ctx, dim, n, perturb=0.1), cdm.Mesh.cartesian(ctx, dim, n, perturb=0.1, sfc_ordering=True)
    (vx, ev, bv, ba), (vxs, evs, bvs, bas) = m.arrays(), ms.arrays()
    assert np.array_equal(vx, vxs) and np.array_equal(bv, bvs) and np.array_equal(evs, ev[lin])
    sp, sps = cdm.H1Space(m, 2), cdm.H1Space(ms, 2)
    assert sp.ndof == sps.ndof                                                           # same space, renumbered


def test_paraview_writer(ctx, tmp_path):
    import xml.etree.ElementTree as ET
    for mesh, p in ((cdm.Mesh.read_gmsh(ctx, os.path.join(GOLD, "square_tri.msh")), 3), (cdm.Mesh.cartesian(ctx, 2, [3, 2]), 2),
                    (cdm.Mesh.cartesian(ctx, 3, [2, 2, 1]), 3)):
        sp = cdm.H1Space(mesh, p)
        X = sp.dof_coords()
        sp.write_paraview(tmp_path, "coll", {"u": X[:, 0] + 2 * X[:, 1], "u_exact": X[:, 0]}, cycle=3, time=0.25)
        assert os.path.exists(tmp_path / "coll" / "coll.pvd") and os.path.exists(tmp_path / "coll" / "Cycle000003" / "data.pvtu")
        piece = ET.parse(tmp_path / "coll" / "Cycle000003" / "proc000000.vtu").getroot().find("UnstructuredGrid/Piece")
        sub = p ** mesh.dim
        assert int(piece.get("NumberOfPoints")) == sp.ne * sp.nd and int(piece.get("NumberOfCells")) == sp.ne * sub
        pts = np.array(piece.find("Points/DataArray").text.split(), float).reshape(-1, 3)
        conn = np.array(piece.find("Cells/DataArray[@Name='connectivity']").text.split(), int)
        vals = np.array(piece.find("PointData/DataArray[@Name='u']").text.split(), float)
        assert np.allclose(vals, pts[:, 0] + 2 * pts[:, 1]) and conn.min() == 0 and conn.max() == sp.ne * sp.nd - 1
        if mesh.dim == 2:                                                                # sub-cells tile the domain
            vpc = 3 if sp.simplex else 4
            c = pts[conn.reshape(-1, vpc)]
            x, y = c[:, :, 0], c[:, :, 1]
            area = 0.5 * np.abs(np.sum(x * np.roll(y, -1, 1) - np.roll(x, -1, 1) * y, axis=1))
            assert abs(area.sum() - 1.0) < 1e-12


@pytest.mark.parametrize("mesh", ["square_tri", "disk_tri"])
@pytest.mark.parametrize("p", [1, 2, 3, 4])
def test_triangle_numbering_matches_independent_restatement(ctx, mesh, p):
    """element-to-dof map (native order = the E-vector order of simplices), essential dofs and dof coordinates of the C++
    host code against the numpy restatement; bit-exact for the integer arrays"""
    from oracle import tri_oracle as T
    m = cdm.Mesh.read_gmsh(ctx, os.path.join(GOLD, mesh + ".msh"))
    vx, ev, bv, ba = m.arrays()
    sp = cdm.H1Space(m, p)
    P = T.TriProblem(p, vx, ev, bv, ba)
    g, o, i = sp.maps()
    assert sp.ndof == P.ndof and sp.nd == P.nd and np.array_equal(g, P.elem_dof)
    assert np.array_equal(sp.essential_dofs(np.ones(int(ba.max()), np.int32)), P.ess)
    assert np.abs(sp.dof_coords() - P.coords()).max() < 1e-15
    if int(ba.max()) == 4:
        mk = np.array([0, 1, 0, 1], np.int32)
        Px = T.TriProblem(p, vx, ev, bv, ba, ess_attrs=(2, 4))
        assert np.array_equal(sp.essential_dofs(mk), Px.ess)


@pytest.mark.parametrize("mesh", ["square_tri", "disk_tri", "quads", "ref_unit_square"])
def test_uniform_refinement_matches_independent_restatement(ctx, mesh):
    """Mesh::UniformRefinement (serial_ref_levels of the reference's drivers; Input/input_diffusion_mms.yaml refines
    Mesh/unit_square.msh once): the C++ host code against the numpy restatement, bit-exact in every array (the midpoint
    coordinates too: same summation order), twice in a row; counts, areas, conformity, inherited boundary attributes"""
    from oracle import refine_oracle as R
    from oracle import tri_oracle as T
    if mesh == "quads":
        m = cdm.Mesh.cartesian(ctx, 2, [5, 3], perturb=0.15)
    elif mesh == "ref_unit_square":
        path = os.path.join(REF, "Mesh", "unit_square.msh")
        if not os.path.exists(path):
            pytest.skip("reference tree not present")
        m = cdm.Mesh.read_gmsh(ctx, path)
    else:
        m = cdm.Mesh.read_gmsh(ctx, os.path.join(GOLD, mesh + ".msh"))
    arrays = m.arrays()
    area0 = None
    for level in range(2):
        vx, ev, bv, ba = arrays
        m = m.uniform_refine()
        got = m.arrays()
        want = R.uniform_refine_2d(vx, ev, bv, ba)
        for g, w in zip(got, want):
            assert g.shape == w.shape and np.array_equal(g, w)
        k = ev.shape[1]
        nedges = len({tuple(sorted((int(e[j]), int(e[(j + 1) % k])))) for e in ev for j in range(k)})
        assert m.nv == len(vx) + nedges + (len(ev) if k == 4 else 0) and m.ne == 4 * len(ev) and m.nbe == 2 * len(bv)
        X = got[0][got[1]]
        x, y = X[:, :, 0], X[:, :, 1]
        area = 0.5 * np.sum(x * np.roll(y, -1, 1) - np.roll(x, -1, 1) * y, axis=1)
        assert area.min() > 0                                     # children keep the orientation
        if area0 is None:
            X0 = vx[ev]
            area0 = 0.5 * np.sum(X0[:, :, 0] * np.roll(X0[:, :, 1], -1, 1) - np.roll(X0[:, :, 0], -1, 1) * X0[:, :, 1], axis=1).sum()
        assert abs(area.sum() - area0) < 1e-13
        # conforming: every interior edge is shared by exactly two children, every boundary edge is a boundary element
        cnt = {}
        for e in got[1]:
            for j in range(k):
                key = tuple(sorted((int(e[j]), int(e[(j + 1) % k]))))
                cnt[key] = cnt.get(key, 0) + 1
        bset = {tuple(sorted((int(a), int(b)))) for a, b in got[2]}
        assert all((c == 2) != (key in bset) for key, c in cnt.items()) and all(cnt[key] == 1 for key in bset)
        # the space on the refined mesh numbers like the restatement (triangles) / builds (quads)
        sp = cdm.H1Space(m, 2)
        if k == 3:
            P = T.TriProblem(2, *got)
            assert sp.ndof == P.ndof and np.array_equal(sp.maps()[0], P.elem_dof)
        else:
            assert sp.ndof == m.nv + len(cnt) + m.ne
        arrays = got


@pytest.mark.parametrize("shuffle", [None, 4])
def test_hexahedral_uniform_refinement(ctx, orc, shuffle):
    """UniformRefinement of hexahedral meshes (perturbed Cartesian, natural and shuffled vertex numbering): bit-exact against the
    numpy restatement, volumes of the children positive and summing to the parents', conforming faces, boundary attributes
    inherited; the H1 numbering and the essential dofs of the C++ host code on the refined mesh equal the oracle's"""
    from oracle import refine_oracle as R
    P0 = orc.Problem(3, 1, [3, 2, 2], perturb=0.12, shuffle_seed=shuffle)
    m = cdm.Mesh.from_arrays(ctx, P0.vx, P0.ev, P0.bv, P0.battr)
    gp = np.array([0.5 - 0.5 / np.sqrt(3.0), 0.5 + 0.5 / np.sqrt(3.0)])

    def volumes(vx, ev):
        X = vx[ev]                                                # (ne, 8, 3), MFEM hex vertex order
        vol = np.zeros(len(ev))
        for x in gp:
            for y in gp:
                for z in gp:
                    dN = np.array([[-(1 - y) * (1 - z), -(1 - x) * (1 - z), -(1 - x) * (1 - y)], [(1 - y) * (1 - z), -x * (1 - z), -x * (1 - y)],
                                   [y * (1 - z), x * (1 - z), -x * y], [-y * (1 - z), (1 - x) * (1 - z), -(1 - x) * y],
                                   [-(1 - y) * z, -(1 - x) * z, (1 - x) * (1 - y)], [(1 - y) * z, -x * z, x * (1 - y)],
                                   [y * z, x * z, x * y], [-y * z, (1 - x) * z, (1 - x) * y]])
                    J = np.einsum("ekr,kc->erc", X, dN)
                    vol += np.linalg.det(J) / 8.0
        return vol

    arrays = m.arrays()
    v0 = volumes(arrays[0], arrays[1]).sum()
    for level in range(2):
        vx, ev, bv, ba = arrays
        m = m.uniform_refine()
        got = m.arrays()
        want = R.uniform_refine_3d(vx, ev, bv, ba)
        for g, w in zip(got, want):
            assert g.shape == w.shape and np.array_equal(g, w)
        assert m.ne == 8 * len(ev) and m.nbe == 4 * len(bv)
        vol = volumes(got[0], got[1])
        assert vol.min() > 0 and abs(vol.sum() - v0) < 1e-13
        cnt = {}
        for e in got[1]:
            for f in R.HEX_F:
                key = tuple(sorted(int(e[i]) for i in f))
                cnt[key] = cnt.get(key, 0) + 1
        bset = {tuple(sorted(int(t) for t in b)) for b in got[2]}
        assert all((c == 2) != (key in bset) for key, c in cnt.items()) and all(cnt[key] == 1 for key in bset) and len(bset) == m.nbe
        for p in (2, 3):
            sp = cdm.H1Space(m, p)
            ndof, elem_dof, _, _ = orc.h1_build(3, p, m.nv, got[1])
            assert sp.ndof == ndof and np.array_equal(sp.maps()[0], elem_dof)
            marker = np.array([1, 0, 1, 1, 0, 1], np.int32)
            mark = orc.h1_bdr_dofs(3, p, m.nv, got[1], got[2], got[3], marker, ndof)
            assert np.array_equal(sp.essential_dofs(marker), np.nonzero(mark)[0].astype(np.int32))
        arrays = got


def test_refined_triangle_meshes_converge_at_the_expected_rate(ctx):
    """the steady MMS problem of linear_convection_diffusion_2D.cpp:159-215 on square_tri.msh and on its uniform refinement
    (CPU oracle solve): the L2 error falls by ~2^(p+1) per level"""
    import scipy.sparse.linalg as spla
    from oracle import tri_oracle as T
    kappa, cx, cy, s, n, mm = 0.1, 1.0, -2.0, 1.0, 1, 1
    uex = lambda x, y: np.sin(n * np.pi * x) * np.sin(mm * np.pi * y)
    f = lambda x, y: (kappa * (n * n + mm * mm) * np.pi ** 2 + s) * uex(x, y) \
        + cx * n * np.pi * np.cos(n * np.pi * x) * np.sin(mm * np.pi * y) + cy * mm * np.pi * np.sin(n * np.pi * x) * np.cos(mm * np.pi * y)
    m = cdm.Mesh.read_gmsh(ctx, os.path.join(GOLD, "square_tri.msh"))
    for p in (1, 2):
        errs = []
        mesh = m
        for level in range(2):
            vx, ev, bv, ba = mesh.arrays()
            P = T.TriProblem(p, vx, ev, bv, ba, kappa=kappa, vel=(cx, cy), mass=s)
            A = P.csr().to_scipy().tocsr()
            nq = p + 3
            xq = P.rule_coords(nq)
            b = P.domain_lf(f(xq[..., 0], xq[..., 1]), nq)
            X = P.coords()
            u = np.zeros(P.ndof)
            u[P.ess] = uex(X[P.ess, 0], X[P.ess, 1])
            free = np.flatnonzero(P.ess_mark == 0)
            rhs = b - A @ u
            u[free] = spla.spsolve(A[free][:, free].tocsc(), rhs[free])
            errs.append(P.l2_error(u, uex(xq[..., 0], xq[..., 1]), nq))
            mesh = mesh.uniform_refine()
        rate = np.log2(errs[0] / errs[1])
        assert p + 0.6 < rate < p + 1.6, (p, errs, rate)


def test_triangle_oracle_known_answers():
    """the numpy triangle oracle itself: K 1 = 0, C 1 = 0, 1^T M 1 = area, symmetry, exactness of the nodal interpolation"""
    from oracle import tri_oracle as T
    ctx = cdm.Context(host_only=True)
    vx, ev, bv, ba = cdm.Mesh.read_gmsh(ctx, os.path.join(GOLD, "square_tri.msh")).arrays()
    for p in (1, 2, 3, 4):
        one = np.ones(T.TriProblem(p, vx, ev, bv, ba).ndof)
        K = T.TriProblem(p, vx, ev, bv, ba, kappa=1.0, vel=None, mass=None).csr().to_scipy()
        Cm = T.TriProblem(p, vx, ev, bv, ba, kappa=None, vel=(1.0, -2.0), mass=None).csr().to_scipy()
        M = T.TriProblem(p, vx, ev, bv, ba, kappa=None, vel=None, mass=1.0).csr().to_scipy()
        assert abs(K @ one).max() < 1e-12 and abs(Cm @ one).max() < 1e-12 and abs(one @ (M @ one) - 1.0) < 1e-13
        assert abs(K - K.T).max() < 1e-12 and abs(M - M.T).max() < 1e-14
        B = T.eval_basis(p, T.tri_nodes(p))
        assert np.abs(B - np.eye(len(B))).max() < 1e-12
        P = T.TriProblem(p, vx, ev, bv, ba)
        X = P.coords()
        u = (1 + X[:, 0]) ** p - X[:, 1] ** p                                             # in P_p: interpolated exactly
        xq = P.rule_coords(p + 2)
        assert P.l2_error(u, (1 + xq[..., 0]) ** p - xq[..., 1] ** p, p + 2) < 1e-13


def test_oracle_ilu0_defining_property(orc):
    import scipy.sparse as sps
    P = orc.Problem(2, 2, 6, perturb=0.1, vel=(1.0, -2.0))
    A = P.csr()
    b = np.random.default_rng(0).uniform(-1, 1, P.ndof)
    A.eliminate(P.ess_mark, np.zeros(P.ndof), b)
    f = A.ilu0()
    F = sps.csr_matrix((f.factors(), A.colind, A.rowptr), shape=(A.n, A.n))
    L, U = sps.tril(F, -1) + sps.eye(A.n), sps.triu(F, 0)
    S = A.to_scipy()
    mask = S.copy(); mask.data[:] = 1.0
    assert abs((L @ U - S).multiply(mask)).max() < 1e-14                                   # (LU)_ij = a_ij on the pattern of A
    z = f.solve(b)
    assert np.linalg.norm(L @ (U @ z) - b) < 1e-13 * np.linalg.norm(b)
    x, info = A.gmres_ilu(b, f)
    _, info_j = A.op().gmres(b, dinv=1 / A.diag())
    assert info["converged"] and info["iters"] < info_j["iters"]
    assert np.linalg.norm(A.spmv(x) - b) < 1e-8 * np.linalg.norm(b)
