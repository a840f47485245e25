"""ctypes binding of libcdm_b200.so (the C ABI in include/cdm_b200.h) plus thin
host-side classes that mirror the reference's MFEM surface for this path:

    Mesh / H1Space              Mesh, H1_FECollection + ParFiniteElementSpace
                                (linear_convection_diffusion_2D.cpp:290-312)
    ConvectionDiffusionOperator ParBilinearForm + Diffusion/Convection/Mass integrators,
                                Operator::Mult, FormLinearSystem  (:335-351)
    GMRESSolver / CGSolver      PetscLinearSolver (:368-374), mfem::CGSolver
                                (mesh_recession_handler.cpp:270-276)

PyTorch is used only as the owner of device memory / the CUDA stream; there is no
CPU fallback: every compute call raises if the CUDA library or a GPU is missing.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# CDM_B200_LIB: load another build of the same library (kernel-tuning experiments, scripts/build_variant.sh)
LIB_PATH = os.environ.get("CDM_B200_LIB") or os.path.join(_HERE, "libcdm_b200.so")

OK, EINVAL, ENOGPU, ECUDA, ENOMEM, ENCCL, EUNSUP = 0, -1, -2, -3, -4, -5, -6
COEFF_NONE, COEFF_CONST, COEFF_QPT = 0, 1, 2
GMRES_PETSC, GMRES_MFEM = 0, 1


class CdmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"cdm error {code}: {msg}")
        self.code = code


class Coeff(C.Structure):
    _fields_ = [("kind", C.c_int), ("ncomp", C.c_int), ("data", C.c_void_p)]


class KrylovOpts(C.Structure):
    _fields_ = [("variant", C.c_int), ("restart", C.c_int), ("max_it", C.c_int),
                ("rtol", C.c_double), ("atol", C.c_double), ("zero_guess", C.c_int),
                ("jacobi", C.c_int)]


class KrylovResult(C.Structure):
    _fields_ = [("iters", C.c_int), ("converged", C.c_int), ("final_norm", C.c_double),
                ("hist_len", C.c_int), ("seconds", C.c_double)]


def build(force=False, verbose=False):
    """Compile the CUDA extension in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    src = os.path.join(_HERE, "csrc")
    if force:
        subprocess.run(["make", "-C", src, "clean"], check=True, capture_output=not verbose)
    subprocess.run(["make", "-C", src, "-j8"], check=True, capture_output=not verbose)
    return LIB_PATH


_lib = None

# every symbol include/cdm_b200.h declares, with its signature
_vp, _i64, _ci, _cd = C.c_void_p, C.c_int64, C.c_int, C.c_double
_pp = C.POINTER(C.c_void_p)
SIGNATURES = {
    "cdm_init": (_ci, [_ci, _vp, _pp]),
    "cdm_init_host": (_ci, [_pp]),
    "cdm_finalize": (_ci, [_vp]),
    "cdm_last_error": (C.c_char_p, [_vp]),
    "cdm_sync": (_ci, [_vp]),
    "cdm_stream": (_vp, [_vp]),
    "cdm_version": (C.c_char_p, []),
    "cdm_comm_unique_id": (_ci, [_vp]),
    "cdm_comm_init": (_ci, [_vp, _ci, _ci, _vp]),
    "cdm_comm_rank": (_ci, [_vp, C.POINTER(_ci), C.POINTER(_ci)]),
    "cdm_mesh_cartesian": (_ci, [_vp, _ci, C.POINTER(_i64), C.POINTER(_cd), _cd, _pp]),
    "cdm_mesh_from_arrays": (_ci, [_vp, _ci, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _pp]),
    "cdm_mesh_from_arrays_simplex": (_ci, [_vp, _ci, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _pp]),
    "cdm_mesh_read_gmsh": (_ci, [_vp, C.c_char_p, _ci, _pp]),
    "cdm_mesh_cartesian_sfc": (_ci, [_vp, _ci, C.POINTER(_i64), C.POINTER(_cd), _cd, _pp]),
    "cdm_grid_sfc_ordering": (_ci, [_ci, C.POINTER(_i64), _vp]),
    "cdm_mesh_geometry": (_ci, [_vp, C.POINTER(_ci), C.POINTER(_ci), C.POINTER(_ci)]),
    "cdm_mesh_uniform_refine": (_ci, [_vp, _vp, _pp]),
    "cdm_mesh_sizes": (_ci, [_vp, C.POINTER(_ci), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    "cdm_mesh_get": (_ci, [_vp, _vp, _vp, _vp, _vp]),
    "cdm_mesh_destroy": (_ci, [_vp]),
    "cdm_mesh_partition_box": (_ci, [_vp, _vp, C.POINTER(_ci), _ci, _pp]),
    "cdm_mesh_partition_elements": (_ci, [_vp, _vp, C.POINTER(C.c_int32), _ci, _ci, _pp]),
    "cdm_space_create_h1": (_ci, [_vp, _vp, _ci, _pp]),
    "cdm_space_create_from_table": (_ci, [_vp, _vp, _ci, _i64, _vp, _pp]),
    "cdm_space_sizes": (_ci, [_vp, C.POINTER(_ci), C.POINTER(_ci), C.POINTER(_i64), C.POINTER(_i64),
                              C.POINTER(_ci), C.POINTER(_ci), C.POINTER(_i64)]),
    "cdm_space_get_maps": (_ci, [_vp, _vp, _vp, _vp]),
    "cdm_space_essential_dofs": (_ci, [_vp, _vp, _ci, _vp, C.POINTER(_i64)]),
    "cdm_space_get_basis": (_ci, [_vp, _vp, _vp, _vp, _vp]),
    "cdm_space_dof_coords": (_ci, [_vp, _vp]),
    "cdm_space_qpt_coords": (_ci, [_vp, _vp]),
    "cdm_space_halo_peers": (_ci, [_vp, C.POINTER(_ci)]),
    "cdm_space_halo_peer": (_ci, [_vp, _ci, C.POINTER(_ci), C.POINTER(_i64), C.POINTER(_i64), _vp, _vp]),
    "cdm_space_dof_global": (_ci, [_vp, _vp]),
    "cdm_space_sym_peers": (_ci, [_vp, C.POINTER(_ci), C.POINTER(_i64), C.POINTER(_i64)]),
    "cdm_space_sym_peer": (_ci, [_vp, _ci, C.POINTER(_ci), C.POINTER(_i64), C.POINTER(_i64), _vp]),
    "cdm_space_sym_sum_plan": (_ci, [_vp, _vp, _vp, _vp]),
    "cdm_space_elem_perm": (_ci, [_vp, _vp, C.POINTER(_i64)]),
    "cdm_space_destroy": (_ci, [_vp]),
    "cdm_operator_create": (_ci, [_vp, C.POINTER(Coeff), C.POINTER(Coeff), _cd, C.POINTER(Coeff), _vp, _i64, _pp]),
    "cdm_operator_update": (_ci, [_vp, C.POINTER(Coeff), C.POINTER(Coeff), _cd, C.POINTER(Coeff)]),
    "cdm_operator_destroy": (_ci, [_vp]),
    "cdm_operator_size": (_i64, [_vp]),
    "cdm_operator_local_size": (_i64, [_vp]),
    "cdm_operator_apply": (_ci, [_vp, _vp, _vp]),
    "cdm_operator_apply_unconstrained": (_ci, [_vp, _vp, _vp]),
    "cdm_operator_mult_host": (_ci, [_vp, _vp, _vp, _ci]),
    "cdm_operator_diag": (_ci, [_vp, _vp]),
    "cdm_eliminate_rhs": (_ci, [_vp, _vp, _vp]),
    "cdm_operator_get_qdata": (_ci, [_vp, _vp, _vp, _vp]),
    "cdm_operator_assemble_csr": (_ci, [_vp]),
    "cdm_operator_csr_sizes": (_ci, [_vp, C.POINTER(_i64), C.POINTER(_i64)]),
    "cdm_operator_csr_get": (_ci, [_vp, _vp, _vp, _vp]),
    "cdm_operator_set_option": (_ci, [_vp, C.c_char_p, _ci]),
    "cdm_operator_get_option": (_ci, [_vp, C.c_char_p, C.POINTER(_ci)]),
    "cdm_integrator_add_mult_pa": (_ci, [_vp, _vp, _vp]),
    "cdm_integrator_assemble_diagonal_pa": (_ci, [_vp, _vp]),
    "cdm_restriction_mult": (_ci, [_vp, _vp, _vp]),
    "cdm_restriction_mult_transpose": (_ci, [_vp, _vp, _vp]),
    "cdm_prolongate": (_ci, [_vp, _vp, _vp]),
    "cdm_prolongate_transpose": (_ci, [_vp, _vp, _vp]),
    "cdm_space_local_size": (_i64, [_vp]),
    "cdm_space_true_size": (_i64, [_vp]),
    "cdm_operator_time_kernel": (_ci, [_vp, _vp, _vp, _ci, _ci, C.POINTER(_cd)]),
    "cdm_rule_points": (_ci, [_ci]),
    "cdm_space_rule_coords": (_ci, [_vp, _ci, _vp]),
    "cdm_domain_lf": (_ci, [_vp, _ci, _vp, _cd, _ci, _vp]),
    "cdm_l2_error": (_ci, [_vp, _ci, _vp, _vp, C.POINTER(_cd)]),
    "cdm_vec_set_indexed": (_ci, [_vp, _i64, _vp, _vp, _vp]),
    "cdm_config_load": (_ci, [C.c_char_p, _pp]),
    "cdm_config_destroy": (_ci, [_vp]),
    "cdm_config_has": (_ci, [_vp, C.c_char_p]),
    "cdm_config_get_string": (_ci, [_vp, C.c_char_p, C.c_char_p, _ci]),
    "cdm_config_get_double": (_ci, [_vp, C.c_char_p, C.POINTER(_cd)]),
    "cdm_config_get_int": (_ci, [_vp, C.c_char_p, C.POINTER(_ci)]),
    "cdm_config_get_bool": (_ci, [_vp, C.c_char_p, C.POINTER(_ci)]),
    "cdm_config_get_doubles": (_ci, [_vp, C.c_char_p, _vp, _ci, C.POINTER(_ci)]),
    "cdm_write_paraview": (_ci, [_vp, C.c_char_p, C.c_char_p, _ci, _cd, _ci, _vp, _vp]),
    "cdm_petsc_options_load": (_ci, [C.c_char_p, C.POINTER(KrylovOpts), C.POINTER(_ci)]),
    "cdm_operator_ilu_apply": (_ci, [_vp, _vp, _vp]),
    "cdm_operator_ilu_levels": (_ci, [_vp, C.POINTER(_ci), C.POINTER(_ci)]),
    "cdm_launch_count": (_i64, [_vp]),
    "cdm_vec_alloc": (_ci, [_vp, _i64, _pp]),
    "cdm_vec_free": (_ci, [_vp, _vp]),
    "cdm_vec_set": (_ci, [_vp, _i64, _cd, _vp]),
    "cdm_vec_upload": (_ci, [_vp, _i64, _vp, _vp]),
    "cdm_vec_download": (_ci, [_vp, _i64, _vp, _vp]),
    "cdm_axpy": (_ci, [_vp, _i64, _cd, _vp, _vp]),
    "cdm_add": (_ci, [_vp, _i64, _vp, _cd, _vp, _vp]),
    "cdm_pointwise_mult": (_ci, [_vp, _i64, _vp, _vp, _vp]),
    "cdm_dot": (_ci, [_vp, _i64, _vp, _vp, C.POINTER(_cd)]),
    "cdm_mdot": (_ci, [_vp, _i64, _ci, _vp, _vp, _i64, _vp]),
    "cdm_maxpy": (_ci, [_vp, _i64, _ci, _vp, _vp, _i64, _vp]),
    "cdm_norm2": (_ci, [_vp, _i64, _vp, C.POINTER(_cd)]),
    "cdm_gmres": (_ci, [_vp, _vp, _vp, C.POINTER(KrylovOpts), C.POINTER(KrylovResult), _vp]),
    "cdm_cg": (_ci, [_vp, _vp, _vp, C.POINTER(KrylovOpts), C.POINTER(KrylovResult), _vp]),
}


def lib():
    """Load the CUDA extension; fail loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CdmError(ENOGPU, f"{LIB_PATH} is missing: run __graft_entry__.build() "
                               "(make -C continuum-mechanics-mfem_b200/csrc); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


def _ptr(a):
    """host numpy array / torch tensor / int -> void*"""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))


class Context:
    """Device("cuda:i") analogue; host_only=True allows mesh/space construction without a GPU."""

    def __init__(self, device=0, stream=None, host_only=False):
        self.h = C.c_void_p()
        self.host_only = host_only
        L = lib()
        rc = L.cdm_init_host(C.byref(self.h)) if host_only else L.cdm_init(device, stream, C.byref(self.h))
        if rc != OK:
            raise CdmError(rc, "cdm_init failed (no usable CUDA device; there is no CPU fallback)")
        self.device = device

    def check(self, rc):
        if rc != OK:
            raise CdmError(rc, lib().cdm_last_error(self.h).decode())

    def sync(self):
        self.check(lib().cdm_sync(self.h))

    @property
    def stream(self):
        return lib().cdm_stream(self.h)

    @property
    def launches(self):
        return int(lib().cdm_launch_count(self.h))

    def comm_init(self, rank, nranks, unique_id):
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id)) if unique_id is not None else None
        self.check(lib().cdm_comm_init(self.h, rank, nranks, buf))

    @staticmethod
    def comm_unique_id():
        buf = (C.c_char * 128)()
        rc = lib().cdm_comm_unique_id(buf)
        if rc != OK:
            raise CdmError(rc, "ncclGetUniqueId failed")
        return bytes(buf)

    def close(self):
        if self.h:
            lib().cdm_finalize(self.h)
            self.h = C.c_void_p()

    # ---- vector kernels (device pointers: torch tensors)
    def dot(self, x, y):
        r = C.c_double()
        self.check(lib().cdm_dot(self.h, x.numel(), _ptr(x), _ptr(y), C.byref(r)))
        return r.value

    def norm2(self, x):
        r = C.c_double()
        self.check(lib().cdm_norm2(self.h, x.numel(), _ptr(x), C.byref(r)))
        return r.value

    def axpy(self, a, x, y):
        self.check(lib().cdm_axpy(self.h, x.numel(), a, _ptr(x), _ptr(y)))

    def add(self, x, a, y, z):
        self.check(lib().cdm_add(self.h, x.numel(), _ptr(x), a, _ptr(y), _ptr(z)))

    def mdot(self, w, V, k):
        out = np.zeros(k)
        self.check(lib().cdm_mdot(self.h, w.numel(), k, _ptr(w), _ptr(V), V.stride(0), _ptr(out)))
        return out

    def maxpy(self, h, V, w):
        h = np.ascontiguousarray(h, np.float64)
        self.check(lib().cdm_maxpy(self.h, w.numel(), len(h), _ptr(h), _ptr(V), V.stride(0), _ptr(w)))


class Mesh:
    def __init__(self, ctx, handle):
        self.ctx, self.h = ctx, handle
        dim, nv, ne, nbe = C.c_int(), C.c_int64(), C.c_int64(), C.c_int64()
        lib().cdm_mesh_sizes(self.h, C.byref(dim), C.byref(nv), C.byref(ne), C.byref(nbe))
        self.dim, self.nv, self.ne, self.nbe = dim.value, nv.value, ne.value, nbe.value

    @classmethod
    def cartesian(cls, ctx, dim, n, size=None, perturb=0.0, sfc_ordering=False):
        """Mesh::MakeCartesian2D/3D; sfc_ordering=True lists the elements along the Hilbert curve (MFEM's inline meshes)"""
        n = list(n) if not np.isscalar(n) else [n] * dim
        nn = (C.c_int64 * 3)(*((n + [0] * 3)[:3]))
        ss = (C.c_double * 3)(*((list(size) + [1.0] * 3)[:3] if size is not None else [1.0] * 3))
        h = C.c_void_p()
        fn = lib().cdm_mesh_cartesian_sfc if sfc_ordering else lib().cdm_mesh_cartesian
        ctx.check(fn(ctx.h, dim, nn, ss, float(perturb), C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def read_gmsh(cls, ctx, path, refine=True):
        """mfem::Mesh(path, 1, refine) for Gmsh 2.2 ASCII files (linear_convection_diffusion_2D.cpp:290)"""
        h = C.c_void_p()
        ctx.check(lib().cdm_mesh_read_gmsh(ctx.h, os.fspath(path).encode(), 1 if refine else 0, C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def from_arrays_simplex(cls, ctx, vx, ev, bv, battr):
        vx = np.ascontiguousarray(vx, np.float64)
        ev = np.ascontiguousarray(ev, np.int32)
        bv = np.ascontiguousarray(bv, np.int32)
        battr = np.ascontiguousarray(battr, np.int32)
        h = C.c_void_p()
        ctx.check(lib().cdm_mesh_from_arrays_simplex(ctx.h, vx.shape[1], vx.shape[0], _ptr(vx), ev.shape[0], _ptr(ev),
                                                     bv.shape[0], _ptr(bv), _ptr(battr), C.byref(h)))
        return cls(ctx, h)

    @property
    def geometry(self):
        """(geom, vertices per element, vertices per boundary element); geom 0 tensor, 1 simplex"""
        g, a, b = C.c_int(), C.c_int(), C.c_int()
        lib().cdm_mesh_geometry(self.h, C.byref(g), C.byref(a), C.byref(b))
        return g.value, a.value, b.value

    @classmethod
    def from_arrays(cls, ctx, vx, ev, bv, battr):
        vx = np.ascontiguousarray(vx, np.float64)
        ev = np.ascontiguousarray(ev, np.int32)
        bv = np.ascontiguousarray(bv, np.int32)
        battr = np.ascontiguousarray(battr, np.int32)
        h = C.c_void_p()
        ctx.check(lib().cdm_mesh_from_arrays(ctx.h, vx.shape[1], vx.shape[0], _ptr(vx), ev.shape[0], _ptr(ev),
                                             bv.shape[0], _ptr(bv), _ptr(battr), C.byref(h)))
        return cls(ctx, h)

    def uniform_refine(self, levels=1):
        """Mesh::UniformRefinement() `levels` times (triangles, quadrilaterals, hexahedra; serial_ref_levels + par_ref_levels of the
        reference's drivers, linear_convection_diffusion_2D.cpp:295-298)"""
        mesh = self
        for _ in range(int(levels)):
            h = C.c_void_p()
            self.ctx.check(lib().cdm_mesh_uniform_refine(self.ctx.h, mesh.h, C.byref(h)))
            mesh = Mesh(self.ctx, h)
        return mesh

    def partition_box(self, parts, rank):
        pp = (C.c_int * 3)(*(list(parts) + [1] * (3 - len(parts))))
        h = C.c_void_p()
        self.ctx.check(lib().cdm_mesh_partition_box(self.ctx.h, self.h, pp, rank, C.byref(h)))
        return Mesh(self.ctx, h)

    def partition_elements(self, elem_rank, nranks, rank):
        """ParMesh(comm, mesh) with a per-element rank array (METIS-style): this rank's submesh of a quad / hex mesh"""
        er = np.ascontiguousarray(elem_rank, np.int32)
        if er.shape != (self.ne,):
            raise ValueError("elem_rank must have one entry per element")
        h = C.c_void_p()
        self.ctx.check(lib().cdm_mesh_partition_elements(self.ctx.h, self.h, er.ctypes.data_as(C.POINTER(C.c_int32)),
                                                         int(nranks), int(rank), C.byref(h)))
        return Mesh(self.ctx, h)

    def arrays(self):
        _, nve, nvb = self.geometry
        vx = np.zeros((self.nv, self.dim))
        ev = np.zeros((self.ne, nve), np.int32)
        bv = np.zeros((self.nbe, nvb), np.int32)
        battr = np.zeros(self.nbe, np.int32)
        lib().cdm_mesh_get(self.h, _ptr(vx), _ptr(ev), _ptr(bv), _ptr(battr))
        return vx, ev, bv, battr

    def __del__(self):
        if getattr(self, "h", None):
            lib().cdm_mesh_destroy(self.h)
            self.h = None


class H1Space:
    """H1_FECollection(order, dim) + FiniteElementSpace on `mesh`."""

    def __init__(self, mesh, order, elem_dof=None, ndof=None):
        self.mesh, self.ctx = mesh, mesh.ctx
        self.h = C.c_void_p()
        if elem_dof is None:
            self.ctx.check(lib().cdm_space_create_h1(self.ctx.h, mesh.h, order, C.byref(self.h)))
        else:
            elem_dof = np.ascontiguousarray(elem_dof, np.int32)
            self.ctx.check(lib().cdm_space_create_from_table(self.ctx.h, mesh.h, order, int(ndof), _ptr(elem_dof),
                                                             C.byref(self.h)))
        dim, p, d1d, q1d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        ne, ndof_, ntrue = C.c_int64(), C.c_int64(), C.c_int64()
        lib().cdm_space_sizes(self.h, C.byref(dim), C.byref(p), C.byref(ne), C.byref(ndof_), C.byref(d1d),
                              C.byref(q1d), C.byref(ntrue))
        self.dim, self.order, self.ne, self.ndof = dim.value, p.value, ne.value, ndof_.value
        self.d1d, self.q1d, self.ntrue = d1d.value, q1d.value, ntrue.value
        self.nd, self.nq = self.d1d ** self.dim, self.q1d ** self.dim
        self.simplex = mesh.geometry[0] == 1
        if self.simplex:                      # order-p triangle, collapsed Gauss-Legendre rule with q1d^2 points
            self.nd = (self.order + 1) * (self.order + 2) // 2

    def maps(self):
        """ElementRestriction gather_map / offsets / indices (int32)."""
        g = np.zeros((self.ne, self.nd), np.int32)
        o = np.zeros(self.ndof + 1, np.int32)
        i = np.zeros(self.ne * self.nd, np.int32)
        lib().cdm_space_get_maps(self.h, _ptr(g), _ptr(o), _ptr(i))
        return g, o, i

    def essential_dofs(self, marker):
        marker = np.ascontiguousarray(marker, np.int32)
        cnt = C.c_int64()
        self.ctx.check(lib().cdm_space_essential_dofs(self.h, _ptr(marker), len(marker), None, C.byref(cnt)))
        out = np.zeros(cnt.value, np.int32)
        self.ctx.check(lib().cdm_space_essential_dofs(self.h, _ptr(marker), len(marker), _ptr(out), C.byref(cnt)))
        return out

    def basis(self):
        B, G = np.zeros((self.q1d, self.d1d)), np.zeros((self.q1d, self.d1d))
        qw, nodes = np.zeros(self.q1d), np.zeros(self.d1d)
        lib().cdm_space_get_basis(self.h, _ptr(B), _ptr(G), _ptr(qw), _ptr(nodes))
        return B, G, qw, nodes

    def dof_coords(self):
        out = np.zeros((self.ndof, self.dim))
        lib().cdm_space_dof_coords(self.h, _ptr(out))
        return out

    def halo(self):
        """[(peer rank, owned-shared dof ids, ghost dof ids)] of a partitioned space"""
        n = C.c_int()
        lib().cdm_space_halo_peers(self.h, C.byref(n))
        out = []
        for i in range(n.value):
            r, no, ng = C.c_int(), C.c_int64(), C.c_int64()
            lib().cdm_space_halo_peer(self.h, i, C.byref(r), C.byref(no), C.byref(ng), None, None)
            own, ghost = np.zeros(no.value, np.int32), np.zeros(ng.value, np.int32)
            lib().cdm_space_halo_peer(self.h, i, C.byref(r), C.byref(no), C.byref(ng), _ptr(own), _ptr(ghost))
            out.append((r.value, own, ghost))
        return out

    def sym_plan(self):
        """symmetric exchange plan: ([(peer rank, offset, shared dof ids)], sum_dof, sum_off, sum_src)"""
        npe, ns, nc = C.c_int(), C.c_int64(), C.c_int64()
        lib().cdm_space_sym_peers(self.h, C.byref(npe), C.byref(ns), C.byref(nc))
        peers = []
        for i in range(npe.value):
            r, n, off = C.c_int(), C.c_int64(), C.c_int64()
            lib().cdm_space_sym_peer(self.h, i, C.byref(r), C.byref(n), C.byref(off), None)
            idx = np.zeros(n.value, np.int32)
            lib().cdm_space_sym_peer(self.h, i, C.byref(r), C.byref(n), C.byref(off), _ptr(idx))
            peers.append((r.value, off.value, idx))
        dof, off, src = np.zeros(ns.value, np.int32), np.zeros(ns.value + 1, np.int32), np.zeros(nc.value, np.int32)
        lib().cdm_space_sym_sum_plan(self.h, _ptr(dof), _ptr(off), _ptr(src))
        return peers, dof, off, src

    def elem_perm(self):
        """(perm, n_boundary): perm[e] = mesh index of the space's e-th element"""
        perm = np.zeros(self.ne, np.int64)
        nb = C.c_int64()
        lib().cdm_space_elem_perm(self.h, _ptr(perm), C.byref(nb))
        return perm, nb.value

    def dof_global(self):
        k = np.zeros(self.ndof, np.int64)
        lib().cdm_space_dof_global(self.h, _ptr(k))
        return k

    # --- ElementRestriction / prolongation on device vectors
    def restrict(self, xL, xE):
        """ElementRestriction::Mult: xE[e, d] = xL[gather[e, d]]"""
        self.ctx.check(lib().cdm_restriction_mult(self.h, _ptr(xL), _ptr(xE)))

    def restrict_transpose(self, yE, yL):
        """ElementRestriction::MultTranspose (deterministic gather)"""
        self.ctx.check(lib().cdm_restriction_mult_transpose(self.h, _ptr(yE), _ptr(yL)))

    def prolongate(self, xT, uL):
        """RecoverFEMSolution: u_L = P x_T (ghost entries from their owners)"""
        self.ctx.check(lib().cdm_prolongate(self.h, _ptr(xT), _ptr(uL)))

    def prolongate_transpose(self, bL, bT=None):
        """ParallelAssemble: b_T = P^T b_L"""
        self.ctx.check(lib().cdm_prolongate_transpose(self.h, _ptr(bL), _ptr(bT)))

    def qpt_coords(self):
        out = np.zeros((self.ne, self.nq, self.dim))
        lib().cdm_space_qpt_coords(self.h, _ptr(out))
        return out

    # --- linear forms, projection, error norms (the steps either side of the solve)
    def rule_coords(self, q1d=0, out=None):
        """physical points of the q1d^dim Gauss-Legendre rule, (ne, q1d^dim, dim); `out` may be a
        CUDA tensor (the coordinates then stay on the device)"""
        q = self.q1d if q1d == 0 else q1d
        if out is None:
            out = np.zeros((self.ne, q ** self.dim, self.dim))
        self.ctx.check(lib().cdm_space_rule_coords(self.h, q1d, _ptr(out)))
        return out

    def domain_lf(self, f_q, b, q1d=0, scale=1.0, accumulate=False):
        """b = [b +] scale * DomainLFIntegrator(f) assembled on the true dofs (b: device tensor)"""
        if isinstance(f_q, np.ndarray):
            f_q = np.ascontiguousarray(f_q, np.float64)
        self.ctx.check(lib().cdm_domain_lf(self.h, q1d, _ptr(f_q), float(scale), 1 if accumulate else 0, _ptr(b)))
        return b

    def l2_error(self, u, uex_q, q1d=0):
        """ComputeL2Error(u_h, u_exact) with the app's rule; u=None -> ||u_exact||, uex_q=None -> ||u_h||"""
        if isinstance(uex_q, np.ndarray):
            uex_q = np.ascontiguousarray(uex_q, np.float64)
        r = C.c_double(0.0)
        self.ctx.check(lib().cdm_l2_error(self.h, q1d, _ptr(u), _ptr(uex_q), C.byref(r)))
        return r.value

    def write_paraview(self, prefix, collection, fields, cycle=0, time=0.0):
        """ParaViewDataCollection::Save (linear_convection_diffusion_2D.cpp:421-433); fields: {name: host L-vector}"""
        names = list(fields)
        arrs = [np.ascontiguousarray(fields[k], np.float64) for k in names]
        cn = (C.c_char_p * len(names))(*[k.encode() for k in names])
        cp = (C.c_void_p * len(names))(*[a.ctypes.data for a in arrs])
        self.ctx.check(lib().cdm_write_paraview(self.h, os.fspath(prefix).encode(), collection.encode(), int(cycle), float(time),
                                                len(names), cn, cp))

    def project_dofs(self, idx, vals, u):
        """u[idx] = vals (ProjectBdrCoefficient with vals = g at the dof coordinates of idx)"""
        if isinstance(idx, np.ndarray):
            idx = np.ascontiguousarray(idx, np.int32)
        if isinstance(vals, np.ndarray):
            vals = np.ascontiguousarray(vals, np.float64)
        n = idx.numel() if hasattr(idx, "numel") else len(idx)
        self.ctx.check(lib().cdm_vec_set_indexed(self.ctx.h, n, _ptr(idx), _ptr(vals), _ptr(u)))
        return u

    def __del__(self):
        if getattr(self, "h", None):
            lib().cdm_space_destroy(self.h)
            self.h = None


def _coeff(c, dim, which, keep):
    """None | scalar | vector | per-qpt array (ne, nq[, ncomp]) -> Coeff"""
    if c is None:
        return Coeff(COEFF_NONE, 0, None)
    if hasattr(c, "data_ptr"):
        # device-resident per-point values (torch.float64 CUDA tensor, (ne, nq[, ncomp])): used in place
        t = c.contiguous()
        keep.append(t)
        return Coeff(COEFF_QPT, int(t.shape[2]) if t.dim() == 3 else 1, C.c_void_p(t.data_ptr()))
    a = np.ascontiguousarray(np.asarray(c, dtype=np.float64))
    keep.append(a)
    if a.ndim <= 1:
        return Coeff(COEFF_CONST, int(a.size), a.ctypes.data_as(C.c_void_p))
    ncomp = int(a.shape[2]) if a.ndim == 3 else 1
    return Coeff(COEFF_QPT, ncomp, a.ctypes.data_as(C.c_void_p))


class ConvectionDiffusionOperator:
    """a(u,v) = (kappa grad u, grad v) + alpha (vel . grad u, v) + (mass u, v), partially assembled.

    Mult / MultUnconstrained follow mfem::Operator::Mult(x, y) on device vectors
    (torch.float64 CUDA tensors); mult_host takes numpy arrays."""

    def __init__(self, space, kappa=None, vel=None, alpha=1.0, mass=None, ess_dofs=None):
        self.space, self.ctx = space, space.ctx
        keep = []
        if vel is not None and np.ndim(vel) == 1:
            vel = np.asarray(vel, dtype=np.float64)[:space.dim]
        ck, cv, cm = (_coeff(kappa, space.dim, 0, keep), _coeff(vel, space.dim, 1, keep),
                      _coeff(mass, space.dim, 2, keep))
        ess = None if ess_dofs is None else np.ascontiguousarray(ess_dofs, np.int32)
        self.h = C.c_void_p()
        self.ctx.check(lib().cdm_operator_create(space.h, C.byref(ck), C.byref(cv), float(alpha), C.byref(cm),
                                                 _ptr(ess), 0 if ess is None else len(ess), C.byref(self.h)))
        if any(hasattr(k, "data_ptr") for k in keep):
            self.ctx.sync()
        self.height = self.width = int(lib().cdm_operator_size(self.h))
        self.local_size = int(lib().cdm_operator_local_size(self.h))
        self.flags = (kappa is not None, vel is not None, mass is not None)

    def update(self, kappa=None, vel=None, alpha=1.0, mass=None):
        keep = []
        if vel is not None and np.ndim(vel) == 1:
            vel = np.asarray(vel, dtype=np.float64)[:self.space.dim]
        ck, cv, cm = (_coeff(kappa, self.space.dim, 0, keep), _coeff(vel, self.space.dim, 1, keep),
                      _coeff(mass, self.space.dim, 2, keep))
        self.ctx.check(lib().cdm_operator_update(self.h, C.byref(ck), C.byref(cv), float(alpha), C.byref(cm)))
        if any(hasattr(k, "data_ptr") for k in keep):
            self.ctx.sync()          # device-resident coefficient tensors are read asynchronously

    def set_option(self, name, value):
        self.ctx.check(lib().cdm_operator_set_option(self.h, name.encode(), int(value)))

    def get_option(self, name):
        v = C.c_int()
        self.ctx.check(lib().cdm_operator_get_option(self.h, name.encode(), C.byref(v)))
        return v.value

    def assemble_csr(self):
        """full assembly on the device (the reference's a.Assemble()); returns (rowptr, colind, vals) host copies"""
        self.ctx.check(lib().cdm_operator_assemble_csr(self.h))
        n, nnz = _i64(0), _i64(0)
        self.ctx.check(lib().cdm_operator_csr_sizes(self.h, C.byref(n), C.byref(nnz)))
        rowptr, colind, vals = np.zeros(n.value + 1, np.int64), np.zeros(nnz.value, np.int32), np.zeros(nnz.value)
        self.ctx.check(lib().cdm_operator_csr_get(self.h, _ptr(rowptr), _ptr(colind), _ptr(vals)))
        return rowptr, colind, vals

    def Mult(self, x, y):
        self.ctx.check(lib().cdm_operator_apply(self.h, _ptr(x), _ptr(y)))

    def MultUnconstrained(self, x, y):
        self.ctx.check(lib().cdm_operator_apply_unconstrained(self.h, _ptr(x), _ptr(y)))

    def mult_host(self, x, y=None, constrained=True):
        x = np.ascontiguousarray(x, np.float64)
        if y is None:
            y = np.zeros(self.height)
        self.ctx.check(lib().cdm_operator_mult_host(self.h, _ptr(x), _ptr(y), 1 if constrained else 0))
        return y

    def time_kernel(self, x, y, reps=10, constrained=True):
        """mean device time (ms) of the element kernel alone (CUDA events around each launch)"""
        ms = C.c_double()
        self.ctx.check(lib().cdm_operator_time_kernel(self.h, _ptr(x), _ptr(y), reps, 1 if constrained else 0,
                                                      C.byref(ms)))
        return ms.value

    def AddMultPA(self, xE, yE):
        """BilinearFormIntegrator::AddMultPA on E-vectors: yE += B^T D B xE"""
        self.ctx.check(lib().cdm_integrator_add_mult_pa(self.h, _ptr(xE), _ptr(yE)))

    def AssembleDiagonalPA(self, dE):
        """BilinearFormIntegrator::AssembleDiagonalPA: dE += element-wise diagonal"""
        self.ctx.check(lib().cdm_integrator_assemble_diagonal_pa(self.h, _ptr(dE)))

    def ilu_apply(self, r, z):
        """z = (LU)^{-1} r, ILU(0) of the assembled constrained matrix (-pc_type bjacobi -sub_pc_type ilu on one rank)"""
        self.ctx.check(lib().cdm_operator_ilu_apply(self.h, _ptr(r), _ptr(z)))

    def ilu_levels(self):
        a, b = C.c_int(), C.c_int()
        self.ctx.check(lib().cdm_operator_ilu_levels(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def AssembleDiagonal(self, d):
        self.ctx.check(lib().cdm_operator_diag(self.h, _ptr(d)))

    def EliminateRHS(self, x, b):
        self.ctx.check(lib().cdm_eliminate_rhs(self.h, _ptr(x), _ptr(b)))

    def qdata(self):
        sp = self.space
        nsym = sp.dim * (sp.dim + 1) // 2
        Dd = np.zeros((sp.ne, nsym, sp.nq)) if self.flags[0] else None
        Dc = np.zeros((sp.ne, sp.dim, sp.nq)) if self.flags[1] else None
        Dm = np.zeros((sp.ne, sp.nq)) if self.flags[2] else None
        self.ctx.check(lib().cdm_operator_get_qdata(self.h, _ptr(Dd), _ptr(Dc), _ptr(Dm)))
        return Dd, Dc, Dm

    def __del__(self):
        if getattr(self, "h", None):
            lib().cdm_operator_destroy(self.h)
            self.h = None


class _IterativeSolver:
    """mfem::IterativeSolver surface: SetRelTol/SetAbsTol/SetMaxIter/SetOperator/Mult/Get*."""
    _fn = None

    def __init__(self, variant=GMRES_PETSC, restart=0, max_it=500, rtol=1e-10, atol=1e-12, jacobi=True, pc=None):
        # pc: None -> jacobi flag; "none" | "jacobi" | "ilu" (-pc_type bjacobi -sub_pc_type ilu, Input/petsc_circle.opts:6-8)
        pcv = {None: 1 if jacobi else 0, "none": 0, "jacobi": 1, "ilu": 2, "bjacobi": 2}[pc]
        self.opts = KrylovOpts(variant, restart, max_it, rtol, atol, 1, pcv)
        self.res = KrylovResult()
        self.op = None
        self.iterative_mode = False
        self.history = np.zeros(0)

    def SetRelTol(self, v): self.opts.rtol = v
    def SetAbsTol(self, v): self.opts.atol = v
    def SetMaxIter(self, v): self.opts.max_it = int(v)
    def SetKDim(self, v): self.opts.restart = int(v)
    def SetPrintLevel(self, v): pass
    def SetOperator(self, op): self.op = op
    def GetNumIterations(self): return self.res.iters
    def GetConverged(self): return bool(self.res.converged)
    def GetFinalNorm(self): return self.res.final_norm

    def Mult(self, b, x):
        if self.op is None:
            raise CdmError(EINVAL, "SetOperator has not been called")
        self.opts.zero_guess = 0 if self.iterative_mode else 1
        hist = np.zeros(self.opts.max_it + 2)
        fn = getattr(lib(), self._fn)
        self.op.ctx.check(fn(self.op.h, _ptr(b), _ptr(x), C.byref(self.opts), C.byref(self.res), _ptr(hist)))
        self.history = hist[:self.res.hist_len].copy()


class GMRESSolver(_IterativeSolver):
    _fn = "cdm_gmres"

    @classmethod
    def from_petsc_options(cls, path):
        """MFEMInitializePetsc(..., petsc_options_file, ...) + PetscLinearSolver (linear_convection_diffusion_2D.cpp:268-282,368)"""
        s = cls()
        kt = C.c_int()
        rc = lib().cdm_petsc_options_load(os.fspath(path).encode(), C.byref(s.opts), C.byref(kt))
        if rc != OK:
            raise CdmError(rc, f"cannot load PETSc options from {path}")
        s.ksp_type = "cg" if kt.value == 1 else "gmres"
        return s


class CGSolver(_IterativeSolver):
    _fn = "cdm_cg"

    def __init__(self, max_it=500, rtol=1e-12, atol=0.0, jacobi=False):
        super().__init__(GMRES_MFEM, 0, max_it, rtol, atol, jacobi)


class Config:
    """LoadParams (linear_convection_diffusion_2D.cpp:62-127): flat `key: value` YAML; get() keeps the caller's default
    when the key is absent, like `if (n["key"]) p.key = n["key"].as<T>()`."""

    def __init__(self, path):
        self.h = C.c_void_p()
        rc = lib().cdm_config_load(os.fspath(path).encode(), C.byref(self.h))
        if rc != OK:
            raise CdmError(rc, f"cannot load {path}" + (" (nested YAML is not supported)" if rc == EUNSUP else ""))
        self.path = os.fspath(path)

    def has(self, key):
        return bool(lib().cdm_config_has(self.h, key.encode()))

    def get(self, key, default=None, kind=str):
        if not self.has(key):
            return default
        if kind is str:
            buf = C.create_string_buffer(4096)
            lib().cdm_config_get_string(self.h, key.encode(), buf, 4096)
            return buf.value.decode()
        if kind is list:
            n = C.c_int()
            if lib().cdm_config_get_doubles(self.h, key.encode(), None, 0, C.byref(n)) != OK:
                raise CdmError(EINVAL, f"YAML key {key} is not a sequence")
            v = np.zeros(n.value)
            lib().cdm_config_get_doubles(self.h, key.encode(), _ptr(v), n.value, C.byref(n))
            return list(v)
        v = {float: C.c_double, int: C.c_int, bool: C.c_int}[kind]()
        fn = {float: lib().cdm_config_get_double, int: lib().cdm_config_get_int, bool: lib().cdm_config_get_bool}[kind]
        if fn(self.h, key.encode(), C.byref(v)) != OK:
            raise CdmError(EINVAL, f"YAML key {key} has no {kind.__name__} value")
        return kind(v.value)

    def __del__(self):
        if getattr(self, "h", None):
            lib().cdm_config_destroy(self.h)
            self.h = None
