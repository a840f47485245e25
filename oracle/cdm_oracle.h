/* cdm_oracle.h -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * Plain-C restatement of the convection-diffusion hot path of
 * quinnchr-personal/Continuum-Mechanics-MFEM (myapps/convection_diffusion).
 * PARITY UNPINNED: the arithmetic of the reference lives in MFEM / hypre /
 * PETSc, which are un-vendored, unpinned (makefile:4-19 only searches for a
 * config.mk) and absent from this image; the reference commits no golden
 * vectors.  The oracle therefore restates the published MFEM/PETSc
 * definitions twice, independently (assembled CSR and
 * sum-factorised partial assembly), and is pinned only by known-answer tests
 * (tests/test_oracle_*.py) and the app's own manufactured solution.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 */
#ifndef CDM_ORACLE_H
#define CDM_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- 1-D rules and basis (MFEM fem/intrules.cpp, fem/fe/fe_base.cpp) ---- */
void orc_gauss_legendre(int n, double *x, double *w);      /* on [0,1], ascending */
void orc_gauss_lobatto(int n, double *x);                  /* on [0,1], ascending */
/* B,G are Q1D x D1D row-major: B[q*D1D+d] = l_d(x_q), G = l_d'(x_q); D1D=p+1 */
void orc_basis(int p, int q1d, double *B, double *G, double *qw);
int  orc_q1d(int dim, int p);                               /* p+2 (3D), p+1 (2D) */

/* ---- Cartesian mesh (MFEM Mesh::MakeCartesian2D/3D, sfc_ordering=false) ---- */
void orc_cart_sizes(int dim, const int64_t *n, int64_t *nv, int64_t *ne, int64_t *nbe);
void orc_cart_mesh(int dim, const int64_t *n, const double *s, double perturb,
                   double *vx, int32_t *ev, int32_t *bv, int32_t *battr);

/* ---- H1 space: global numbering + lexicographic element->dof table ---- */
/* returns ndof; fills elem_dof[ne * D1D^dim] (lexicographic in element) */
int64_t orc_h1_build(int dim, int p, int64_t nv, int64_t ne, const int32_t *ev,
                     int32_t *elem_dof, int64_t *nedges, int64_t *nfaces);
/* marks (1) every dof lying on a boundary element whose attribute is marked */
int orc_h1_bdr_dofs(int dim, int p, int64_t nv, int64_t ne, const int32_t *ev,
                    int64_t nbe, const int32_t *bv, const int32_t *battr,
                    const int32_t *marker, int nattr, uint8_t *dof_mark);
/* MFEM ElementRestriction index arrays */
void orc_restriction(int64_t ne, int nd, int64_t ndof, const int32_t *gather,
                     int32_t *offsets, int32_t *indices);
/* physical coordinates of every lexicographic node of every element
   (order-1 mesh, GLL nodes): out[ne*nd*dim] -- used to check conformity */
void orc_node_coords(int dim, int p, int64_t ne, const int32_t *ev, const double *vx,
                     double *out);

/* ---- linear forms and error norms (oracle_forms.c) ----
 * orc_rule_coords: physical coordinates of the q1d^dim tensor Gauss-Legendre rule,
 *   out[(e*nq+q)*dim + c] -- where the caller evaluates f / u_exact (Coefficient::Eval).
 * orc_domain_lf: DomainLFIntegrator, b[elem_dof] += scale * w|J| f_q phi_i (MFEM default rule: q1d = p+1).
 * orc_l2_error: GridFunction::ComputeL2Error (u, uex_q), ComputeGlobalLpNorm(2) (u = NULL). */
void orc_rule_coords(int dim, int q1d, int64_t ne, const int32_t *ev, const double *vx, double *out);
void orc_domain_lf(int dim, int p, int q1d, int64_t ne, const int32_t *ev, const double *vx,
                   const int32_t *elem_dof, const double *f_q, double scale, double *b);
double orc_l2_error(int dim, int p, int q1d, int64_t ne, const int32_t *ev, const double *vx,
                    const int32_t *elem_dof, const double *u, const double *uex_q);

/* ---- quadrature data (MFEM bilininteg_{diffusion,convection,mass}_pa) ----
 * coefficient kinds: 0 absent, 1 constant, 2 per-quadrature-point array
 * (point-major: value index = (e*NQ + q)*ncomp + c).
 * kappa: ncomp 1 (scalar) or dim*(dim+1)/2 (symmetric matrix, order 11,21,31,22,32,33)
 * vel:   ncomp dim ; mass: ncomp 1.
 * Outputs (MFEM layout, q fastest): Ddiff[(e*nsym + c)*NQ + q], Dconv[(e*dim+c)*NQ+q],
 * Dmass[e*NQ+q].  Any output may be NULL. */
void orc_qdata(int dim, int p, int64_t ne, const int32_t *ev, const double *vx,
               int kappa_kind, int kappa_ncomp, const double *kappa,
               int vel_kind, const double *vel, double alpha,
               int mass_kind, const double *mass,
               double *Ddiff, double *Dconv, double *Dmass);

/* ---- partial-assembly apply / diagonal (L-vector -> L-vector, unconstrained) ---- */
void orc_pa_apply(int dim, int p, int64_t ne, int64_t ndof,
                  const int32_t *gather, const int32_t *offsets, const int32_t *indices,
                  const double *Ddiff, const double *Dconv, const double *Dmass,
                  const double *x, double *y);
/* fused, order-specialised CPU apply (3D) used for the CPU baseline timing; same result as
   orc_pa_apply up to summation order; yE_work: optional ne*nd scratch */
void orc_pa_apply_fast(int dim, int p, int64_t ne, int64_t ndof,
                       const int32_t *gather, const int32_t *offsets, const int32_t *indices,
                       const double *Ddiff, const double *Dconv, const double *Dmass,
                       const double *x, double *y, double *yE_work);
void orc_pa_diag(int dim, int p, int64_t ne, int64_t ndof,
                 const int32_t *gather, const int32_t *offsets, const int32_t *indices,
                 const double *Ddiff, const double *Dconv, const double *Dmass,
                 double *diag);

/* ---- full assembly (what the reference app executes) ---- */
/* pattern: returns nnz, fills rowptr[ndof+1]; colind may be NULL on the count pass */
int64_t orc_csr_pattern(int64_t ne, int nd, int64_t ndof, const int32_t *elem_dof,
                        int64_t *rowptr, int32_t *colind);
void orc_csr_assemble(int dim, int p, int64_t ne, int64_t ndof, const int32_t *ev,
                      const double *vx, const int32_t *elem_dof,
                      int kappa_kind, int kappa_ncomp, const double *kappa,
                      int vel_kind, const double *vel, double alpha,
                      int mass_kind, const double *mass,
                      const int64_t *rowptr, const int32_t *colind, double *vals);
void orc_csr_spmv(int64_t n, const int64_t *rowptr, const int32_t *colind,
                  const double *vals, const double *x, double *y);
void orc_csr_diag(int64_t n, const int64_t *rowptr, const int32_t *colind,
                  const double *vals, double *d);
/* FormLinearSystem on the assembled matrix: b -= A_e x ; rows+cols of ess
   zeroed, diagonal kept ; b[ess] = A_ii x[ess]  (hypre EliminateBC semantics) */
void orc_csr_eliminate(int64_t n, const int64_t *rowptr, const int32_t *colind,
                       double *vals, const uint8_t *ess_mark, const double *x, double *b);

/* ---- operator handle used by the Krylov oracles ---- */
typedef struct orc_op orc_op;
orc_op *orc_op_csr(int64_t n, const int64_t *rowptr, const int32_t *colind, const double *vals);
/* constrained PA operator (ConstrainedOperator, DIAG_ONE); ess_mark may be NULL */
orc_op *orc_op_pa(int dim, int p, int64_t ne, int64_t ndof,
                  const int32_t *gather, const int32_t *offsets, const int32_t *indices,
                  const double *Ddiff, const double *Dconv, const double *Dmass,
                  const uint8_t *ess_mark);
void orc_op_free(orc_op *op);
void orc_op_mult(const orc_op *op, const double *x, double *y);
int64_t orc_op_size(const orc_op *op);
/* EliminateRHS of ConstrainedOperator: w=0; w[ess]=x[ess]; b -= A w; b[ess]=x[ess] */
void orc_op_eliminate_rhs(const orc_op *op, const double *x, double *b);

typedef struct {
   int    variant;      /* 0: PETSc-like (CGS, restart default 30); 1: mfem-like (MGS, restart default 50) */
   int    restart;
   int    max_it;
   double rtol, atol;
   int    zero_guess;   /* 1: x0 = 0 (PetscLinearSolver iterative_mode=false) */
} orc_krylov_opts;
typedef struct {
   int    iters, converged;
   double final_norm;
   int    hist_len;     /* number of entries written to hist (<= max_it+1) */
} orc_krylov_result;

/* left-preconditioned GMRES(m) with Jacobi dinv (NULL = identity); hist has max_it+1 slots */
void orc_gmres(const orc_op *A, const double *dinv, const double *b, double *x,
               const orc_krylov_opts *o, orc_krylov_result *r, double *hist);
/* mfem::CGSolver restatement; hist = (r,z) values */
/* ILU(0) (natural ordering, PETSc PCILU defaults) and GMRES left-preconditioned with it: -pc_type bjacobi
   -sub_pc_type ilu on one rank (Input/petsc_circle.opts:6-8) */
typedef struct orc_ilu orc_ilu;
orc_ilu *orc_ilu0_factor(int64_t n, const int64_t *rowptr, const int32_t *colind, const double *vals);
void orc_ilu0_solve(const orc_ilu *f, const double *r, double *z);
void orc_ilu0_get(const orc_ilu *f, double *lu);
void orc_ilu0_free(orc_ilu *f);
void orc_gmres_ilu(const orc_op *A, const orc_ilu *ilu, const double *b, double *x,
                   const orc_krylov_opts *o, orc_krylov_result *res, double *hist);
void orc_cg(const orc_op *A, const double *dinv, const double *b, double *x,
            const orc_krylov_opts *o, orc_krylov_result *r, double *hist);

int orc_num_threads(void);
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
