/* cdm_b200.h -- C ABI of the B200-native convection-diffusion hot path.
 *
 * This is the drop-in boundary of the library: plain pointers and sizes, no
 * C++/torch types.  Each entry point names the reference interface it stands
 * behind (file:line relative to
 * /root/reference/myapps/convection_diffusion/; "MFEM:" = the upstream MFEM
 * virtual the reference calls through, MFEM itself is not vendored).
 *
 * Conventions
 *   - every function returns 0 on success or a negative CDM_E* code; it never
 *     throws across the ABI.  cdm_last_error(ctx) returns a message.
 *   - a non-converged Krylov solve is NOT an error: it is reported through
 *     cdm_krylov_result.converged (solver.GetConverged(),
 *     linear_convection_diffusion_2D.cpp:371).
 *   - host arrays passed in are copied; the library owns all device memory
 *     behind the opaque handles.  Pointers named *_dev are device pointers
 *     owned by the caller (or by cdm_vec_alloc); pointers named *_host are
 *     host pointers.
 *   - a context is bound to one CUDA device and one stream; calls on one
 *     context are not thread-safe, different contexts are independent.
 *   - there is no CPU fallback: every compute entry point fails with
 *     CDM_ENOGPU when no CUDA device is usable.
 */
#ifndef CDM_B200_H
#define CDM_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CDM_OK        0
#define CDM_EINVAL   -1   /* bad argument */
#define CDM_ENOGPU   -2   /* no usable CUDA device / CUDA runtime failure at init */
#define CDM_ECUDA    -3   /* CUDA error during a call */
#define CDM_ENOMEM   -4
#define CDM_ENCCL    -5   /* NCCL error */
#define CDM_EUNSUP   -6   /* unsupported order / dimension */

typedef struct cdm_ctx   cdm_ctx;
typedef struct cdm_mesh  cdm_mesh;
typedef struct cdm_space cdm_space;
typedef struct cdm_op    cdm_op;

/* ---------------------------------------------------------------- context */
/* Device("cpu") at linear_convection_diffusion_2D.cpp:287 becomes one of these.
   stream: a cudaStream_t to run on, or NULL to create a private one. */
int  cdm_init(int device, void *stream, cdm_ctx **ctx);
int  cdm_finalize(cdm_ctx *ctx);
const char *cdm_last_error(const cdm_ctx *ctx);
int  cdm_sync(cdm_ctx *ctx);                       /* cudaStreamSynchronize */
void *cdm_stream(cdm_ctx *ctx);
const char *cdm_version(void);
/* host-only context (mesh / space construction, no device); compute calls fail */
int  cdm_init_host(cdm_ctx **ctx);

/* multi-GPU: one context per rank; unique_id = 128-byte ncclUniqueId produced
   by cdm_comm_unique_id on rank 0 and distributed by the caller
   (MPI_COMM_WORLD at linear_convection_diffusion_2D.cpp:240,300). */
int  cdm_comm_unique_id(void *unique_id_128);
int  cdm_comm_init(cdm_ctx *ctx, int rank, int nranks, const void *unique_id_128);
int  cdm_comm_rank(const cdm_ctx *ctx, int *rank, int *nranks);

/* ------------------------------------------------------------------- mesh */
/* Mesh(file)+UniformRefinement (linear_convection_diffusion_2D.cpp:290-298)
   for the Cartesian configs: MFEM Mesh::MakeCartesian2D/3D numbering
   (sfc_ordering=false), optional smooth interior perturbation (units of h). */
int  cdm_mesh_cartesian(cdm_ctx *ctx, int dim, const int64_t n[3], const double size[3],
                        double perturb, cdm_mesh **mesh);
/* any conforming quad/hex mesh: MFEM vertex order within elements */
int  cdm_mesh_from_arrays(cdm_ctx *ctx, int dim, int64_t nv, const double *vertices,
                          int64_t ne, const int32_t *elem_vtx,
                          int64_t nbe, const int32_t *bdr_vtx, const int32_t *bdr_attr,
                          cdm_mesh **mesh);
/* triangle meshes (the meshes the reference ships: Mesh/unit_square.msh, unit_circle.msh): 3 vertices per element,
   2 per boundary segment, counter-clockwise.  Spaces on them are order-p Lagrange triangles; the operator is applied as
   the assembled sparse matrix (there is no tensor structure to sum-factorise). */
int  cdm_mesh_from_arrays_simplex(cdm_ctx *ctx, int dim, int64_t nv, const double *vertices,
                                  int64_t ne, const int32_t *elem_vtx,
                                  int64_t nbe, const int32_t *bdr_vtx, const int32_t *bdr_attr, cdm_mesh **mesh);
/* Mesh(mesh_file, 1, 1) for Gmsh 2.2 ASCII files (linear_convection_diffusion_2D.cpp:290, Input/input_2d.yaml:1):
   triangles, quadrilaterals or hexahedra as elements, lines / quadrilaterals as boundary elements, physical tags as
   attributes; vertices in file order (unused ones dropped), elements in file order.  mark_for_refinement = the
   `refine` constructor argument: triangles are rotated so that their longest edge comes first. */
int  cdm_mesh_read_gmsh(cdm_ctx *ctx, const char *path, int mark_for_refinement, cdm_mesh **mesh);
/* Mesh::MakeCartesian2D/3D with sfc_ordering = true (MFEM's default for inline-quad / inline-hex meshes): elements
   listed along the generalised Hilbert curve; cdm_grid_sfc_ordering returns the curve itself, coords[k*dim + c]. */
int  cdm_mesh_cartesian_sfc(cdm_ctx *ctx, int dim, const int64_t n[3], const double size[3], double perturb, cdm_mesh **mesh);
int  cdm_grid_sfc_ordering(int dim, const int64_t n[3], int64_t *coords);
/* Mesh::UniformRefinement() (linear_convection_diffusion_2D.cpp:295-298: serial_ref_levels / par_ref_levels) of a triangle,
   quadrilateral or hexahedral mesh: edge midpoints, then face / element centres appended to the vertices, four (2D) or eight
   (3D) children per element in MFEM's order, boundary elements split the same way.  The result is a plain mesh without
   Cartesian provenance: partition it with cdm_mesh_partition_elements (a part cannot be refined). */
int  cdm_mesh_uniform_refine(cdm_ctx *ctx, const cdm_mesh *mesh, cdm_mesh **refined);
/* geom: 0 tensor-product elements, 1 simplices */
int  cdm_mesh_geometry(const cdm_mesh *mesh, int *geom, int *verts_per_elem, int *verts_per_bdr);
int  cdm_mesh_sizes(const cdm_mesh *mesh, int *dim, int64_t *nv, int64_t *ne, int64_t *nbe);
int  cdm_mesh_get(const cdm_mesh *mesh, double *vertices, int32_t *elem_vtx,
                  int32_t *bdr_vtx, int32_t *bdr_attr);       /* any pointer may be NULL */
int  cdm_mesh_destroy(cdm_mesh *mesh);
/* ParMesh(MPI_COMM_WORLD, *mesh) (linear_convection_diffusion_2D.cpp:300): element
   partition of a Cartesian mesh into px*py*pz boxes; returns this rank's submesh.
   Shared-dof bookkeeping happens in cdm_space_create_h1 on the returned mesh. */
int  cdm_mesh_partition_box(cdm_ctx *ctx, const cdm_mesh *global, const int parts[3], int rank,
                            cdm_mesh **local);
/* ParMesh(comm, mesh) with a per-element rank array (what METIS hands MFEM; linear_convection_diffusion_2D.cpp:300), for
   quadrilateral / hexahedral meshes that are not Cartesian boxes (Gmsh input, refined meshes): every rank holds the parent
   mesh and the same array elem_rank[parent elements] (values 0 .. nranks-1, nranks <= 64) and gets its submesh; a space
   created on it carries the same shared-dof plans as a box part (owner = lowest rank of the sharing group). */
int  cdm_mesh_partition_elements(cdm_ctx *ctx, const cdm_mesh *global, const int32_t *elem_rank, int nranks, int rank,
                                 cdm_mesh **local);

/* ------------------------------------------------------------------ space */
/* H1_FECollection(order,dim) + (Par)FiniteElementSpace
   (linear_convection_diffusion_2D.cpp:311-312): MFEM entity-major numbering,
   GLL nodes, lexicographic ElementRestriction. */
int  cdm_space_create_h1(cdm_ctx *ctx, const cdm_mesh *mesh, int order, cdm_space **space);
/* same, but with the element->dof table supplied by the caller (e.g. taken from a
   live mfem::FiniteElementSpace): elem_dof[ne * (order+1)^dim], lexicographic in
   element; bit-exact index maps are then derived from MFEM's own numbering. */
int  cdm_space_create_from_table(cdm_ctx *ctx, const cdm_mesh *mesh, int order, int64_t ndof,
                                 const int32_t *elem_dof, cdm_space **space);
int  cdm_space_sizes(const cdm_space *space, int *dim, int *order, int64_t *ne, int64_t *ndof,
                     int *d1d, int *q1d, int64_t *ntrue);
/* ElementRestriction index arrays (MFEM fem/restriction.cpp): gather_map[ne*nd],
   offsets[ndof+1], indices[ne*nd] -- host copies, any pointer may be NULL */
int  cdm_space_get_maps(const cdm_space *space, int32_t *gather_map, int32_t *offsets,
                        int32_t *indices);
/* fespace.GetEssentialTrueDofs(ess_bdr, list) (linear_convection_diffusion_2D.cpp:319-322).
   bdr_marker[nattr]; call with list=NULL to get the count. Sorted ascending. */
int  cdm_space_essential_dofs(const cdm_space *space, const int32_t *bdr_marker, int nattr,
                              int32_t *list, int64_t *count);
/* 1-D tables used by the kernels: B,G are q1d x d1d row-major, qw[q1d], nodes[d1d] */
int  cdm_space_get_basis(const cdm_space *space, double *B, double *G, double *qw, double *nodes);
/* physical coordinates of the dofs (ndof x dim), e.g. to project coefficients
   (ProjectBdrCoefficient, linear_convection_diffusion_2D.cpp:347) */
int  cdm_space_dof_coords(const cdm_space *space, double *xyz_host);
/* physical coordinates of the quadrature points (ne x nq x dim): what
   Coefficient::Eval(T, ip) sees (linear_convection_diffusion_2D.cpp:165) */
int  cdm_space_qpt_coords(const cdm_space *space, double *xyz_host);
/* shared-dof exchange plan of a partitioned space (ParFiniteElementSpace group
   communicator): number of neighbour ranks; for neighbour i its rank, the owned dofs
   it shares (sent by P, summed into by P^T) and the ghost dofs it owns.  Index
   arrays may be NULL to query the counts.  dof_global: lattice key of each local dof. */
int  cdm_space_halo_peers(const cdm_space *space, int *npeers);
int  cdm_space_halo_peer(const cdm_space *space, int i, int *rank, int64_t *n_own, int64_t *n_ghost,
                         int32_t *own_idx, int32_t *ghost_idx);
int  cdm_space_dof_global(const cdm_space *space, int64_t *keys);
/* symmetric exchange plan of a partitioned space (the default multi-GPU protocol, one exchange per apply): every
   rank of a dof's sharing group sends its partial sum to every other rank of the group; the contributions are
   added in ascending rank order, so all sharers hold the bitwise identical sum and the L-vector stays consistent
   on its ghost entries.  sym_peers: number of neighbour ranks, of distinct shared dofs and of CSR contributions.
   sym_peer i: its rank, the local dofs shared with it (global-key order: the same order on both sides) and the
   offset of its segment in the receive buffer.  sum_plan: dof[n_shared], off[n_shared+1], src[n_contrib] --
   src >= 0: index into the receive buffer, src < 0: this rank's own value. */
int  cdm_space_sym_peers(const cdm_space *space, int *npeers, int64_t *n_shared, int64_t *n_contrib);
int  cdm_space_sym_peer(const cdm_space *space, int i, int *rank, int64_t *n, int64_t *offset, int32_t *idx);
int  cdm_space_sym_sum_plan(const cdm_space *space, int32_t *dof, int32_t *off, int32_t *src);
/* element order of the space: perm[e] = index in the mesh of the space's e-th element (identity
   unless the space is partitioned: then the n_boundary elements touching a shared dof come first,
   so that the halo exchange overlaps the interior elements).  Per-element inputs (gather map,
   CDM_COEFF_QPT arrays, cdm_space_qpt_coords) are in the space's element order. */
int  cdm_space_elem_perm(const cdm_space *space, int64_t *perm, int64_t *n_boundary);
int  cdm_space_destroy(cdm_space *space);

/* --------------------------------------------------------------- operator */
#define CDM_COEFF_NONE  0   /* integrator absent */
#define CDM_COEFF_CONST 1   /* ConstantCoefficient / VectorConstantCoefficient */
#define CDM_COEFF_QPT   2   /* values at quadrature points: data[(e*nq+q)*ncomp + c] (host) */
typedef struct
{
   int kind;             /* CDM_COEFF_* */
   int ncomp;            /* kappa: 1 or dim*(dim+1)/2 (sym. matrix 11,21,31,22,32,33);
                            velocity: dim; mass: 1 */
   const double *data;   /* host pointer */
} cdm_coeff;

/* ParBilinearForm a; a.AddDomainIntegrator(new DiffusionIntegrator(kappa));
   a.AddDomainIntegrator(new ConvectionIntegrator(vel, conv_alpha));
   a.AddDomainIntegrator(new MassIntegrator(mass)); a.Assemble();
   (linear_convection_diffusion_2D.cpp:335-339, linear_convection_diffusion_1D.cpp:391-400)
   = BilinearFormIntegrator::AssemblePA: geometric factors + D-tensors on the device.
   ess_dofs: essential dof list used by the constrained apply (may be NULL/0). */
int  cdm_operator_create(cdm_space *space, const cdm_coeff *kappa, const cdm_coeff *vel,
                         double conv_alpha, const cdm_coeff *mass,
                         const int32_t *ess_dofs, int64_t n_ess, cdm_op **op);
/* re-run the quadrature-data setup with new coefficients (per-step re-assembly,
   diffusion_mms_ale.cpp:1017-1023) */
int  cdm_operator_update(cdm_op *op, const cdm_coeff *kappa, const cdm_coeff *vel,
                         double conv_alpha, const cdm_coeff *mass);
int  cdm_operator_destroy(cdm_op *op);
int64_t cdm_operator_size(const cdm_op *op);        /* height = width = true dofs on this rank */
/* local (L-vector) size = true dofs + ghost dofs of this rank.  Vectors allocated with this
   length may be passed to cdm_operator_apply after cdm_operator_set_option(op, "tail", 1):
   the entries past the true size are scratch for the halo exchange and no T<->L copy is made. */
int64_t cdm_operator_local_size(const cdm_op *op);
/* MFEM: Operator::Mult of FormLinearSystem's constrained operator
   (ConstrainedOperator, DIAG_ONE): y = A z, z = x with z[ess]=0, y[ess]=x[ess].
   In multi-GPU mode x,y are this rank's T-vectors and the call includes the
   P / P^T halo exchange. */
int  cdm_operator_apply(cdm_op *op, const double *x_dev, double *y_dev);
/* MFEM: BilinearForm::Mult on L-vectors, no constraints
   (mass_form.Mult(c, rhs), linear_convection_diffusion_1D.cpp:544) */
int  cdm_operator_apply_unconstrained(cdm_op *op, const double *x_dev, double *y_dev);
/* same two calls with host vectors (mfem::Vector under Device("cpu")): the
   host<->device copies are inside the call */
int  cdm_operator_mult_host(cdm_op *op, const double *x_host, double *y_host, int constrained);
/* MFEM: BilinearFormIntegrator::AssembleDiagonalPA + OperatorJacobiSmoother
   (-pc_type jacobi, Input/petsc.opts:6): d = diag(A), d[ess] = 1 */
int  cdm_operator_diag(cdm_op *op, double *d_dev);
/* MFEM: ConstrainedOperator::EliminateRHS (a.FormLinearSystem,
   linear_convection_diffusion_2D.cpp:351): w=0; w[ess]=x[ess]; b -= A w; b[ess]=x[ess] */
int  cdm_eliminate_rhs(cdm_op *op, const double *x_dev, double *b_dev);
/* raw quadrature data (tests): host copy in MFEM layout, q fastest:
   Ddiff[(e*nsym+c)*nq+q], Dconv[(e*dim+c)*nq+q], Dmass[e*nq+q]; pointers may be NULL */
int  cdm_operator_get_qdata(const cdm_op *op, double *Ddiff, double *Dconv, double *Dmass);
/* The reference's literal path on the device, for comparison with the matrix-free operator: a.Assemble()
   into a sparse matrix (linear_convection_diffusion_2D.cpp:339) from the same quadrature data, and MatMult
   inside the Krylov solve (:368) as a CSR SpMV; essential rows / columns are treated like the matrix
   FormLinearSystem hands to the solver (:351).  Single rank.  cdm_operator_set_option(op, "assembly", 1)
   assembles on demand and switches cdm_operator_apply / cdm_gmres / cdm_cg to the SpMV; cdm_operator_update
   refills the values.  cdm_operator_csr_get: host copies, any pointer may be NULL (pattern: rows sorted). */
int  cdm_operator_assemble_csr(cdm_op *op);
int  cdm_operator_csr_sizes(const cdm_op *op, int64_t *nrows, int64_t *nnz);
int  cdm_operator_csr_get(const cdm_op *op, int64_t *rowptr, int32_t *colind, double *vals);
/* tuning knobs (benchmarks): "assembly" (0 partial assembly, default; 1 full assembly + SpMV, see above), "scatter" (0 E-vector + deterministic gather transpose, 1 fp64 red.add, default),
   "kernel" (0 block kernel, 1-3 order-3 warp kernels, 4 group kernel, 5 sub-warp kernel, 6 2D thread-per-element kernel;
   default by dimension and order),
   "ilu_sweep" (ILU(0) apply: 1 (default) one launch per triangular sweep, rows wait for their dependencies on the
   device; 0 one launch per dependency level),
   "tail" (1: caller vectors have the local size, see cdm_operator_local_size),
   "overlap" (multi-GPU: 0 serial halo exchange, 1 overlapped with interior elements up to 3 neighbours, 2 always),
   "halo" (multi-GPU shared-dof protocol: 2 (default) ONE symmetric peer-memory exchange per apply, the result is
   consistent on the ghost entries too; 0 P and P^T over NCCL send/recv; 1 P and P^T over peer-memory stores;
   setting it to 1 is a collective call, mode 2 is set up collectively by the first apply / solve),
   "ghost_in" (with "tail": 1 = the caller guarantees that the ghost entries of x are consistent, e.g. x was produced
   by a previous apply in halo mode 2 or by cdm_prolongate: the P exchange is skipped),
   "allreduce" (1 (default) Krylov scalars all-reduced through peer memory, 0 ncclAllReduce), "host_pipeline" (0/1, or a chunk count for cdm_operator_mult_host),
   "grid_cap" (> 0: upper bound on the persistent grids, so that a small test mesh runs many elements per warp) */
int  cdm_operator_set_option(cdm_op *op, const char *name, int value);
/* current value of an option; "halo" and "allreduce" report the protocol actually in use (the peer-memory
   protocols fall back to NCCL on every rank when their setup fails on any rank) */
int  cdm_operator_get_option(const cdm_op *op, const char *name, int *value);
/* ------------------- integrator-level (E-vector) and prolongation entry points */
/* MFEM: BilinearFormIntegrator::AddMultPA(const Vector &x_E, Vector &y_E) -- what a ParBilinearForm under
   AssemblyLevel::PARTIAL dispatches to after its own ElementRestriction (a.Assemble(),
   linear_convection_diffusion_2D.cpp:339; the Mult of :368 / linear_convection_diffusion_1D.cpp:544):
     y_E += B^T D B x_E   for all integrators of `op` at once.
   E-vectors are MFEM's: [ne][nd], dof fastest (lexicographic in the element), device pointers. */
int  cdm_integrator_add_mult_pa(cdm_op *op, const double *xE_dev, double *yE_dev);
/* MFEM: BilinearFormIntegrator::AssembleDiagonalPA(Vector &diag_E): diag_E += diag(B^T D B), element by element */
int  cdm_integrator_assemble_diagonal_pa(cdm_op *op, double *diagE_dev);
/* MFEM: ElementRestriction::Mult / MultTranspose (fem/restriction.cpp) with the space's index arrays:
     xE[i] = xL[gather_map[i]]  ;  yL[g] = sum_{j in [offsets[g], offsets[g+1])} yE[indices[j]]   (fixed order) */
int  cdm_restriction_mult(cdm_space *space, const double *xL_dev, double *xE_dev);
int  cdm_restriction_mult_transpose(cdm_space *space, const double *yE_dev, double *yL_dev);
/* MFEM: a.RecoverFEMSolution(X, b, u) (linear_convection_diffusion_2D.cpp:377; per step and block at
   linear_convection_diffusion_1D.cpp:569-572) = u_L = P X: the true dofs are the first
   cdm_space_true_size entries of the local vector, the ghost entries are fetched from their owners.
   uL_dev has cdm_space_local_size entries; xT_dev may alias it.  Collective over the ranks. */
int  cdm_prolongate(cdm_space *space, const double *xT_dev, double *uL_dev);
/* MFEM: ParLinearForm::ParallelAssemble (linear_convection_diffusion_2D.cpp:343) = b_T = P^T b_L: partial sums
   held in ghost entries are added into their owners (fixed peer order).  bL_dev is updated in place; bT_dev
   (may be NULL or alias bL_dev) receives the first cdm_space_true_size entries.  Collective. */
int  cdm_prolongate_transpose(cdm_space *space, double *bL_dev, double *bT_dev);
int64_t cdm_space_local_size(const cdm_space *space);
int64_t cdm_space_true_size(const cdm_space *space);

/* measurement hook: run the element kernel of the apply `reps` times on this rank's
   L-vectors and return its mean device time (CUDA events recorded on the context's
   stream directly around each launch; no halo, no essential fix-up) */
int  cdm_operator_time_kernel(cdm_op *op, const double *x_dev, double *y_dev, int reps, int constrained,
                              double *mean_ms);
/* ------------------------------ linear forms, projection and error norms */
/* Points per direction of MFEM's IntRules.Get(SEGMENT/SQUARE/CUBE, order): order/2 + 1. */
int  cdm_rule_points(int order);
/* Physical coordinates of the tensor Gauss-Legendre rule with q1d points per direction in every
   element, xyz[(e*q1d^dim + q)*dim + c], x fastest: what Coefficient::Eval(T, ip) sees for that rule
   (forcing_coeff / exact_coeff, linear_convection_diffusion_2D.cpp:159-215).  q1d = 0 selects the
   operator's rule.  xyz may be a host or a device pointer (computed on the device either way). */
int  cdm_space_rule_coords(const cdm_space *space, int q1d, double *xyz);
/* ParLinearForm b; b.AddDomainIntegrator(new DomainLFIntegrator(f)); b.Assemble(); (+ ParallelAssemble)
   (linear_convection_diffusion_2D.cpp:341-343; once per step at diffusion_mms.cpp:433-437):
     b_dev = (accumulate ? b_dev : 0) + scale * P^T sum_e B^T (w |J| f)
   f_q[e*q1d^dim + q]: f at cdm_space_rule_coords(space, q1d) (host or device pointer).
   q1d = 0 selects DomainLFIntegrator's default rule (order 2p -> p+1 points). b_dev: true-dof vector. */
int  cdm_domain_lf(cdm_space *space, int q1d, const double *f_q, double scale, int accumulate,
                   double *b_dev);
/* u.ComputeL2Error(exact, irs) / ComputeGlobalLpNorm(2, exact, mesh, irs)
   (linear_convection_diffusion_2D.cpp:383-392, diffusion_mms.cpp:466-470):
     result = sqrt( sum_e sum_q w |J| (u_h(x_q) - uex_q)^2 ), all-reduced over the ranks.
   u_dev NULL -> ||uex||, uex_q NULL -> ||u_h||.  q1d = 0 selects the app's rule (order max(2, 2p+3)).
   uex_q: host or device pointer.  Fixed launch shape: the result is bit-reproducible run to run. */
int  cdm_l2_error(cdm_space *space, int q1d, const double *u_dev, const double *uex_q,
                  double *result_host);
/* u_dev[idx[i]] = vals[i]: u.ProjectBdrCoefficient(g, ess_bdr) with idx = the essential dofs and
   vals = g at their cdm_space_dof_coords (linear_convection_diffusion_2D.cpp:347; per step at
   linear_convection_diffusion_1D.cpp:545-546).  idx / vals: host or device pointers. */
int  cdm_vec_set_indexed(cdm_ctx *ctx, int64_t n, const int32_t *idx, const double *vals,
                         double *u_dev);

/* -------------------------------------------- driver input and output files */
/* LoadParams (linear_convection_diffusion_2D.cpp:62-127): the flat `key: value` YAML the drivers read (scalars,
   quoted strings, [a, b] flow sequences, # comments).  Getters return CDM_EINVAL for a missing / malformed key, so the
   caller keeps its default exactly like `if (n["key"]) p.key = n["key"].as<T>()`. */
typedef struct cdm_config cdm_config;
int  cdm_config_load(const char *path, cdm_config **cfg);
int  cdm_config_destroy(cdm_config *cfg);
int  cdm_config_has(const cdm_config *cfg, const char *key);
int  cdm_config_get_string(const cdm_config *cfg, const char *key, char *buf, int buflen);
int  cdm_config_get_double(const cdm_config *cfg, const char *key, double *value);
int  cdm_config_get_int(const cdm_config *cfg, const char *key, int *value);
int  cdm_config_get_bool(const cdm_config *cfg, const char *key, int *value);
int  cdm_config_get_doubles(const cdm_config *cfg, const char *key, double *values, int maxn, int *n);
/* MFEMInitializePetsc(&argc, &argv, petsc_options_file, NULL) (:268-282) for the options the drivers use
   (Input/petsc.opts:2-6, Input/petsc_circle.opts:2-8): -ksp_type gmres|cg, -ksp_rtol, -ksp_atol, -ksp_max_it,
   -ksp_gmres_restart, -pc_type none|jacobi|ilu|bjacobi (+ -sub_pc_type ilu).  Starts from PETSc's defaults.
   ksp_type: 0 gmres, 1 cg; the preconditioner lands in opts->jacobi (0 none, 1 Jacobi, 2 block-Jacobi + ILU(0)). */
/* ParaViewDataCollection ... Save() (:421-433): <prefix>/<collection>/<collection>.pvd + Cycle%06d/{data.pvtu,
   proc%06d.vtu}; every element refined into order^dim linear cells on its node lattice; fields are host L-vectors. */
int  cdm_write_paraview(const cdm_space *space, const char *prefix_path, const char *collection, int cycle, double time,
                        int nfields, const char *const *names, const double *const *fields_host);

/* number of kernel launches issued by this context so far */
int64_t cdm_launch_count(const cdm_ctx *ctx);

/* ---------------------------------------------------------------- vectors */
int  cdm_vec_alloc(cdm_ctx *ctx, int64_t n, double **v_dev);
int  cdm_vec_free(cdm_ctx *ctx, double *v_dev);
int  cdm_vec_set(cdm_ctx *ctx, int64_t n, double value, double *v_dev);
int  cdm_vec_upload(cdm_ctx *ctx, int64_t n, const double *host, double *v_dev);
int  cdm_vec_download(cdm_ctx *ctx, int64_t n, const double *v_dev, double *host);
/* MFEM linalg/vector.cpp kernels used by CGSolver/GMRESSolver
   (mesh_recession_handler.cpp:270-276): y = a*x + y ; z = x + a*y ; dot */
int  cdm_axpy(cdm_ctx *ctx, int64_t n, double a, const double *x_dev, double *y_dev);
int  cdm_add(cdm_ctx *ctx, int64_t n, const double *x_dev, double a, const double *y_dev,
             double *z_dev);
int  cdm_pointwise_mult(cdm_ctx *ctx, int64_t n, const double *d_dev, const double *x_dev,
                        double *y_dev);
/* InnerProduct(comm, x, y) (newton_petsc_solver.hpp:82-85): fixed-order local
   reduction + all-reduce when a communicator is attached */
int  cdm_dot(cdm_ctx *ctx, int64_t n, const double *x_dev, const double *y_dev, double *result_host);
/* k dot products of w against V_0..V_{k-1} (column i at V + i*ldv) in one pass */
int  cdm_mdot(cdm_ctx *ctx, int64_t n, int k, const double *w_dev, const double *V_dev, int64_t ldv,
              double *result_host);
/* w -= sum_i h[i] V_i */
int  cdm_maxpy(cdm_ctx *ctx, int64_t n, int k, const double *h_host, const double *V_dev,
               int64_t ldv, double *w_dev);
int  cdm_norm2(cdm_ctx *ctx, int64_t n, const double *x_dev, double *result_host);

/* ----------------------------------------------------------------- Krylov */
#define CDM_GMRES_PETSC 0   /* classical Gram-Schmidt, default restart 30 (PETSc KSPGMRES) */
#define CDM_GMRES_MFEM  1   /* modified Gram-Schmidt, default restart 50 (mfem::GMRESSolver) */
typedef struct
{
   int    variant;     /* CDM_GMRES_* (ignored by CG) */
   int    restart;     /* 0 = variant default */
   int    max_it;      /* -ksp_max_it 500 */
   double rtol, atol;  /* -ksp_rtol 1e-10 -ksp_atol 1e-12 ; CG: rel 1e-12 abs 0 */
   int    zero_guess;  /* 1: iterative_mode=false (PetscLinearSolver default) */
   int    jacobi;      /* -pc_type: 0 none, 1 jacobi (diag from cdm_operator_diag), 2 bjacobi + ilu: ILU(0) of the
                          assembled matrix, one block per rank (Input/petsc_circle.opts:6-8; cdm_gmres, single rank) */
} cdm_krylov_opts;
typedef struct
{
   int    iters, converged;   /* GetNumIterations / GetConverged */
   double final_norm;         /* GetFinalNorm */
   int    hist_len;           /* entries written to hist */
   double seconds;            /* device time of the solve (CUDA events) */
} cdm_krylov_result;
/* PetscLinearSolver(A).Mult(B,X) with Input/petsc.opts (linear_convection_diffusion_2D.cpp:368-370):
   left-preconditioned restarted GMRES on the constrained operator.
   hist_host: max_it+2 slots for the preconditioned residual history, or NULL. */
int  cdm_gmres(cdm_op *op, const double *b_dev, double *x_dev, const cdm_krylov_opts *opts,
               cdm_krylov_result *result, double *hist_host);
/* mfem::CGSolver::Mult (mesh_recession_handler.cpp:270-276); hist = (r,z) per iteration */
int  cdm_cg(cdm_op *op, const double *b_dev, double *x_dev, const cdm_krylov_opts *opts,
            cdm_krylov_result *result, double *hist_host);
int  cdm_petsc_options_load(const char *path, cdm_krylov_opts *opts, int *ksp_type);
/* PCApply of -pc_type ilu / bjacobi + ilu: z = (LU)^{-1} r with the ILU(0) factors of the matrix the solver sees
   (assembles the matrix and factorises on first use); cdm_operator_ilu_levels: dependency levels of the two sweeps */
int  cdm_operator_ilu_apply(cdm_op *op, const double *r_dev, double *z_dev);
int  cdm_operator_ilu_levels(cdm_op *op, int *forward, int *backward);

#ifdef __cplusplus
}
#endif
#endif /* CDM_B200_H */
