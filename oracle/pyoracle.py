"""ctypes front-end of the CPU ORACLE (test infrastructure, not the product).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  PARITY UNPINNED: see cdm_oracle.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "build", "libcdm_oracle.so")

i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if (not force and os.path.exists(_LIB)
            and all(os.path.getmtime(_LIB) >= os.path.getmtime(s) for s in srcs)):
        return _LIB
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB


class KOpts(C.Structure):
    _fields_ = [("variant", C.c_int), ("restart", C.c_int), ("max_it", C.c_int),
                ("rtol", C.c_double), ("atol", C.c_double), ("zero_guess", C.c_int)]


class KRes(C.Structure):
    _fields_ = [("iters", C.c_int), ("converged", C.c_int), ("final_norm", C.c_double),
                ("hist_len", C.c_int)]


_lib = None


def _opt(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    vp, i64, ci, cd = C.c_void_p, C.c_int64, C.c_int, C.c_double
    L.orc_gauss_legendre.argtypes = [ci, f64p, f64p]
    L.orc_gauss_lobatto.argtypes = [ci, f64p]
    L.orc_basis.argtypes = [ci, ci, f64p, f64p, f64p]
    L.orc_q1d.argtypes = [ci, ci]
    L.orc_q1d.restype = ci
    L.orc_cart_sizes.argtypes = [ci, i64p, i64p, i64p, i64p]
    L.orc_cart_mesh.argtypes = [ci, i64p, f64p, cd, f64p, i32p, i32p, i32p]
    L.orc_h1_build.argtypes = [ci, ci, i64, i64, i32p, i32p, i64p, i64p]
    L.orc_h1_build.restype = i64
    L.orc_h1_bdr_dofs.argtypes = [ci, ci, i64, i64, i32p, i64, i32p, i32p, i32p, ci, u8p]
    L.orc_h1_bdr_dofs.restype = ci
    L.orc_restriction.argtypes = [i64, ci, i64, i32p, i32p, i32p]
    L.orc_node_coords.argtypes = [ci, ci, i64, i32p, f64p, f64p]
    L.orc_qdata.argtypes = [ci, ci, i64, i32p, f64p, ci, ci, vp, ci, vp, cd, ci, vp, vp, vp, vp]
    L.orc_pa_apply.argtypes = [ci, ci, i64, i64, i32p, i32p, i32p, vp, vp, vp, f64p, f64p]
    L.orc_pa_apply_fast.argtypes = [ci, ci, i64, i64, i32p, i32p, i32p, vp, vp, vp, f64p, f64p, vp]
    L.orc_pa_diag.argtypes = [ci, ci, i64, i64, i32p, i32p, i32p, vp, vp, vp, f64p]
    L.orc_csr_pattern.argtypes = [i64, ci, i64, i32p, i64p, vp]
    L.orc_csr_pattern.restype = i64
    L.orc_csr_assemble.argtypes = [ci, ci, i64, i64, i32p, f64p, i32p, ci, ci, vp, ci, vp, cd,
                                   ci, vp, i64p, i32p, f64p]
    L.orc_csr_spmv.argtypes = [i64, i64p, i32p, f64p, f64p, f64p]
    L.orc_csr_diag.argtypes = [i64, i64p, i32p, f64p, f64p]
    L.orc_csr_eliminate.argtypes = [i64, i64p, i32p, f64p, u8p, f64p, f64p]
    L.orc_op_csr.argtypes = [i64, i64p, i32p, f64p]
    L.orc_op_csr.restype = vp
    L.orc_op_pa.argtypes = [ci, ci, i64, i64, i32p, i32p, i32p, vp, vp, vp, vp]
    L.orc_op_pa.restype = vp
    L.orc_op_free.argtypes = [vp]
    L.orc_op_mult.argtypes = [vp, f64p, f64p]
    L.orc_op_eliminate_rhs.argtypes = [vp, f64p, f64p]
    L.orc_gmres.argtypes = [vp, vp, f64p, f64p, C.POINTER(KOpts), C.POINTER(KRes), f64p]
    L.orc_cg.argtypes = [vp, vp, f64p, f64p, C.POINTER(KOpts), C.POINTER(KRes), f64p]
    L.orc_ilu0_factor.argtypes = [i64, i64p, i32p, f64p]
    L.orc_ilu0_factor.restype = vp
    L.orc_ilu0_solve.argtypes = [vp, f64p, f64p]
    L.orc_ilu0_get.argtypes = [vp, f64p]
    L.orc_ilu0_free.argtypes = [vp]
    L.orc_gmres_ilu.argtypes = [vp, vp, f64p, f64p, C.POINTER(KOpts), C.POINTER(KRes), f64p]
    L.orc_rule_coords.argtypes = [ci, ci, i64, i32p, f64p, f64p]
    L.orc_domain_lf.argtypes = [ci, ci, ci, i64, i32p, f64p, i32p, f64p, cd, f64p]
    L.orc_l2_error.argtypes = [ci, ci, ci, i64, i32p, f64p, i32p, vp, vp]
    L.orc_l2_error.restype = cd
    L.orc_num_threads.restype = ci
    L.orc_set_num_threads.argtypes = [ci]
    _lib = L
    return L


def gauss_legendre(n):
    x, w = np.zeros(n), np.zeros(n)
    lib().orc_gauss_legendre(n, x, w)
    return x, w


def gauss_lobatto(n):
    x = np.zeros(n)
    lib().orc_gauss_lobatto(n, x)
    return x


def q1d(dim, p):
    return lib().orc_q1d(dim, p)


def basis(p, q):
    B, G, w = np.zeros((q, p + 1)), np.zeros((q, p + 1)), np.zeros(q)
    lib().orc_basis(p, q, B, G, w)
    return B, G, w


def cart_mesh(dim, n, s=None, perturb=0.0):
    n = np.asarray(list(n) + [0] * (3 - len(n)), dtype=np.int64)
    s = np.asarray([1.0] * 3 if s is None else list(s) + [1.0] * (3 - len(s)), dtype=np.float64)
    nv, ne, nbe = (np.zeros(1, np.int64) for _ in range(3))
    lib().orc_cart_sizes(dim, n, nv, ne, nbe)
    nv, ne, nbe = int(nv[0]), int(ne[0]), int(nbe[0])
    vx = np.zeros((nv, dim))
    ev = np.zeros((ne, 2 ** dim), np.int32)
    bv = np.zeros((nbe, 2 ** (dim - 1)), np.int32)
    battr = np.zeros(nbe, np.int32)
    lib().orc_cart_mesh(dim, n, s, float(perturb), vx, ev, bv, battr)
    return vx, ev, bv, battr


def shuffle_vertices(vx, ev, bv, seed):
    """Renumber the vertices with a seeded permutation (exercises edge reversal
    and all quad-face orientations of the H1 numbering)."""
    rng = np.random.default_rng(seed)
    perm = rng.permutation(vx.shape[0]).astype(np.int32)     # old -> new
    vx2 = np.zeros_like(vx)
    vx2[perm] = vx
    return vx2, perm[ev].astype(np.int32), perm[bv].astype(np.int32)


def h1_build(dim, p, nv, ev):
    ne = ev.shape[0]
    nd = (p + 1) ** dim
    ed = np.zeros((ne, nd), np.int32)
    nedges, nfaces = np.zeros(1, np.int64), np.zeros(1, np.int64)
    ndof = lib().orc_h1_build(dim, p, nv, ne, np.ascontiguousarray(ev), ed, nedges, nfaces)
    return int(ndof), ed, int(nedges[0]), int(nfaces[0])


def h1_bdr_dofs(dim, p, nv, ev, bv, battr, marker, ndof):
    mark = np.zeros(ndof, np.uint8)
    marker = np.ascontiguousarray(marker, np.int32)
    rc = lib().orc_h1_bdr_dofs(dim, p, nv, ev.shape[0], np.ascontiguousarray(ev), bv.shape[0],
                               np.ascontiguousarray(bv), battr, marker, len(marker), mark)
    assert rc == 0
    return mark


def restriction(gather, ndof):
    ne, nd = gather.shape
    off = np.zeros(ndof + 1, np.int32)
    ind = np.zeros(ne * nd, np.int32)
    lib().orc_restriction(ne, nd, ndof, gather, off, ind)
    return off, ind


def node_coords(dim, p, ev, vx):
    ne = ev.shape[0]
    out = np.zeros((ne, (p + 1) ** dim, dim))
    lib().orc_node_coords(dim, p, ne, np.ascontiguousarray(ev), np.ascontiguousarray(vx), out)
    return out


def _coef(c):
    """-> (kind, ncomp, array or None)"""
    if c is None:
        return 0, 1, None
    a = np.ascontiguousarray(np.asarray(c, dtype=np.float64))
    if a.ndim <= 1:
        return 1, int(a.size), a.reshape(-1)
    return 2, int(a.shape[-1]) if a.ndim == 3 else 1, a


class Problem:
    """A discretised convection-diffusion-reaction operator on a Cartesian mesh.

    kappa / vel / mass: None (integrator absent), a constant (scalar or vector),
    or a per-quadrature-point array of shape (ne, nq[, ncomp])."""

    def __init__(self, dim, p, n, perturb=0.1, kappa=0.1, vel=(1.0, -2.0, 0.5), alpha=1.0,
                 mass=1.0, ess_attrs="all", shuffle_seed=None, sizes=None):
        self.dim, self.p = dim, p
        n = [n] * dim if np.isscalar(n) else list(n)
        self.n = n
        self.vx, self.ev, self.bv, self.battr = cart_mesh(dim, n, sizes, perturb)
        if shuffle_seed is not None:
            self.vx, self.ev, self.bv = shuffle_vertices(self.vx, self.ev, self.bv, shuffle_seed)
        self.nv, self.ne = self.vx.shape[0], self.ev.shape[0]
        self.ndof, self.elem_dof, self.nedges, self.nfaces = h1_build(dim, p, self.nv, self.ev)
        self.nd = (p + 1) ** dim
        self.q1d = q1d(dim, p)
        self.nq = self.q1d ** dim
        self.offsets, self.indices = restriction(self.elem_dof, self.ndof)
        nattr = int(self.battr.max())
        if ess_attrs == "all":
            marker = np.ones(nattr, np.int32)
        else:
            marker = np.zeros(nattr, np.int32)
            for a in ess_attrs:
                marker[a - 1] = 1
        self.marker = marker
        self.ess_mark = h1_bdr_dofs(dim, p, self.nv, self.ev, self.bv, self.battr, marker, self.ndof)
        self.ess = np.nonzero(self.ess_mark)[0].astype(np.int32)
        if vel is not None and np.ndim(vel) == 1:
            vel = np.asarray(vel, dtype=np.float64)[:dim]
        self.kappa, self.vel, self.alpha, self.mass = kappa, vel, float(alpha), mass
        self.set_coefficients(kappa, vel, alpha, mass)

    def set_coefficients(self, kappa, vel, alpha, mass):
        dim, nsym = self.dim, self.dim * (self.dim + 1) // 2
        kk, kn, ka = _coef(kappa)
        vk, _, va = _coef(vel)
        mk, _, ma = _coef(mass)
        self._coef = (kk, kn, ka, vk, va, float(alpha), mk, ma)
        self.Dd = np.zeros((self.ne, nsym, self.nq)) if kk else None
        self.Dc = np.zeros((self.ne, dim, self.nq)) if vk else None
        self.Dm = np.zeros((self.ne, self.nq)) if mk else None
        lib().orc_qdata(dim, self.p, self.ne, self.ev, self.vx, kk, kn, _opt(ka), vk, _opt(va),
                        float(alpha), mk, _opt(ma), _opt(self.Dd), _opt(self.Dc), _opt(self.Dm))

    # --- partial assembly
    def pa_apply(self, x):
        y = np.zeros(self.ndof)
        lib().orc_pa_apply(self.dim, self.p, self.ne, self.ndof, self.elem_dof, self.offsets,
                           self.indices, _opt(self.Dd), _opt(self.Dc), _opt(self.Dm),
                           np.ascontiguousarray(x, np.float64), y)
        return y

    def pa_apply_fast(self, x, y=None):
        """fused order-specialised CPU apply (cpu_baseline timing path)"""
        if y is None:
            y = np.zeros(self.ndof)
        if getattr(self, "_yE", None) is None:
            self._yE = np.zeros(self.ne * self.nd)
        lib().orc_pa_apply_fast(self.dim, self.p, self.ne, self.ndof, self.elem_dof, self.offsets,
                                self.indices, _opt(self.Dd), _opt(self.Dc), _opt(self.Dm),
                                np.ascontiguousarray(x, np.float64), y, _opt(self._yE))
        return y

    def pa_diag(self):
        d = np.zeros(self.ndof)
        lib().orc_pa_diag(self.dim, self.p, self.ne, self.ndof, self.elem_dof, self.offsets,
                          self.indices, _opt(self.Dd), _opt(self.Dc), _opt(self.Dm), d)
        return d

    def pa_op(self, constrained=True):
        h = lib().orc_op_pa(self.dim, self.p, self.ne, self.ndof, self.elem_dof, self.offsets,
                            self.indices, _opt(self.Dd), _opt(self.Dc), _opt(self.Dm),
                            _opt(self.ess_mark) if constrained else None)
        return Op(h, self.ndof, keep=self)

    # --- full assembly
    def csr(self):
        rowptr = np.zeros(self.ndof + 1, np.int64)
        nnz = lib().orc_csr_pattern(self.ne, self.nd, self.ndof, self.elem_dof, rowptr, None)
        colind = np.zeros(nnz, np.int32)
        lib().orc_csr_pattern(self.ne, self.nd, self.ndof, self.elem_dof, rowptr, _opt(colind))
        vals = np.zeros(nnz)
        kk, kn, ka, vk, va, al, mk, ma = self._coef
        lib().orc_csr_assemble(self.dim, self.p, self.ne, self.ndof, self.ev, self.vx,
                               self.elem_dof, kk, kn, _opt(ka), vk, _opt(va), al, mk, _opt(ma),
                               rowptr, colind, vals)
        return CSR(rowptr, colind, vals)

    def coords(self):
        """physical coordinates of every global dof (ndof, dim)"""
        xc = node_coords(self.dim, self.p, self.ev, self.vx)
        out = np.zeros((self.ndof, self.dim))
        out[self.elem_dof.reshape(-1)] = xc.reshape(-1, self.dim)
        return out

    # --- linear forms / error norms (oracle_forms.c)
    def rule_coords(self, q1d):
        """physical points of the q1d^dim Gauss-Legendre rule: (ne, q1d^dim, dim)"""
        out = np.zeros((self.ne, q1d ** self.dim, self.dim))
        lib().orc_rule_coords(self.dim, q1d, self.ne, self.ev, self.vx, out)
        return out

    def domain_lf(self, f_q, q1d=None, scale=1.0, b=None):
        """DomainLFIntegrator with MFEM's default rule (p+1 points per direction)"""
        q1d = self.p + 1 if q1d is None else q1d
        b = np.zeros(self.ndof) if b is None else b
        lib().orc_domain_lf(self.dim, self.p, q1d, self.ne, self.ev, self.vx, self.elem_dof,
                            np.ascontiguousarray(f_q, np.float64).reshape(-1), float(scale), b)
        return b

    def l2_error(self, u, uex_q, q1d=None):
        """ComputeL2Error with the reference's rule order max(2, 2p+3) -> p+2 points"""
        q1d = max(2, 2 * self.p + 3) // 2 + 1 if q1d is None else q1d
        uu = None if u is None else np.ascontiguousarray(u, np.float64)
        qq = None if uex_q is None else np.ascontiguousarray(uex_q, np.float64).reshape(-1)
        return lib().orc_l2_error(self.dim, self.p, q1d, self.ne, self.ev, self.vx, self.elem_dof,
                                  _opt(uu), _opt(qq))


class CSR:
    def __init__(self, rowptr, colind, vals):
        self.rowptr, self.colind, self.vals = rowptr, colind, vals
        self.n = len(rowptr) - 1

    def spmv(self, x):
        y = np.zeros(self.n)
        lib().orc_csr_spmv(self.n, self.rowptr, self.colind, self.vals,
                           np.ascontiguousarray(x, np.float64), y)
        return y

    def diag(self):
        d = np.zeros(self.n)
        lib().orc_csr_diag(self.n, self.rowptr, self.colind, self.vals, d)
        return d

    def eliminate(self, ess_mark, x, b):
        """FormLinearSystem on the assembled matrix (in place on vals and b)."""
        lib().orc_csr_eliminate(self.n, self.rowptr, self.colind, self.vals, ess_mark,
                                np.ascontiguousarray(x, np.float64), b)

    def op(self):
        return Op(lib().orc_op_csr(self.n, self.rowptr, self.colind, self.vals), self.n, keep=self)

    def ilu0(self):
        """ILU(0) factors of this matrix (natural ordering): -pc_type bjacobi -sub_pc_type ilu on one rank"""
        return ILU0(self)

    def gmres_ilu(self, b, ilu, restart=0, max_it=2000, rtol=1e-10, atol=1e-12):
        o = KOpts(0, restart, max_it, rtol, atol, 1)
        r = KRes()
        x = np.zeros(self.n)
        hist = np.zeros(max_it + 2)
        A = self.op()
        lib().orc_gmres_ilu(A.h, ilu.h, np.ascontiguousarray(b, np.float64), x, C.byref(o), C.byref(r), hist)
        return x, dict(iters=r.iters, converged=bool(r.converged), final_norm=r.final_norm, hist=hist[:r.hist_len].copy())

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self.vals, self.colind, self.rowptr), shape=(self.n, self.n))


class ILU0:
    def __init__(self, A):
        self.A = A
        self.h = lib().orc_ilu0_factor(A.n, A.rowptr, A.colind, A.vals)

    def solve(self, r):
        z = np.zeros(self.A.n)
        lib().orc_ilu0_solve(self.h, np.ascontiguousarray(r, np.float64), z)
        return z

    def factors(self):
        lu = np.zeros(len(self.A.vals))
        lib().orc_ilu0_get(self.h, lu)
        return lu

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_ilu0_free(self.h)
            self.h = None


class Op:
    def __init__(self, h, n, keep=None):
        self.h, self.n, self._keep = h, n, keep

    def __del__(self):
        if self.h:
            lib().orc_op_free(self.h)
            self.h = None

    def mult(self, x):
        y = np.zeros(self.n)
        lib().orc_op_mult(self.h, np.ascontiguousarray(x, np.float64), y)
        return y

    def eliminate_rhs(self, x, b):
        lib().orc_op_eliminate_rhs(self.h, np.ascontiguousarray(x, np.float64), b)

    def _solve(self, fn, b, dinv, x0, variant, restart, max_it, rtol, atol):
        o = KOpts(variant, restart, max_it, rtol, atol, 1 if x0 is None else 0)
        r = KRes()
        x = np.zeros(self.n) if x0 is None else np.array(x0, dtype=np.float64)
        hist = np.zeros(max_it + 2)
        fn(self.h, _opt(dinv), np.ascontiguousarray(b, np.float64), x, C.byref(o), C.byref(r), hist)
        return x, dict(iters=r.iters, converged=bool(r.converged), final_norm=r.final_norm,
                       hist=hist[:r.hist_len].copy())

    def gmres(self, b, dinv=None, x0=None, variant=0, restart=0, max_it=500, rtol=1e-10, atol=1e-12):
        return self._solve(lib().orc_gmres, b, dinv, x0, variant, restart, max_it, rtol, atol)

    def cg(self, b, dinv=None, x0=None, max_it=500, rtol=1e-12, atol=0.0):
        return self._solve(lib().orc_cg, b, dinv, x0, 1, 0, max_it, rtol, atol)


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    """OpenMP threads of the oracle (torchrun exports OMP_NUM_THREADS=1: callers that time the CPU
    path set the count explicitly)"""
    lib().orc_set_num_threads(int(n))
    return num_threads()
