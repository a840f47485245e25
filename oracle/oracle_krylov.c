/* oracle_krylov.c -- CPU ORACLE (test infrastructure): operator handles and the
 * Krylov methods the reference drives.
 *
 * Reference call sites restated:
 *   linear_convection_diffusion_2D.cpp:368-374 + Input/petsc.opts:2-6
 *        PetscLinearSolver: KSP GMRES, rtol 1e-10, atol 1e-12, max_it 500, PCJACOBI
 *        (PETSc defaults: restart 30, classical Gram-Schmidt without refinement,
 *         left preconditioning, preconditioned residual norm, zero initial guess)
 *   mesh_recession_handler.cpp:270-276   mfem::CGSolver rel 1e-12 abs 0 max 500
 *   newton_petsc_solver.hpp:82-85        InnerProduct -> global dot
 * PETSc / MFEM sources are not vendored: SURVEY.md Appendix C.5-C.7.  PARITY UNPINNED.
 */
#include "cdm_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

struct orc_op
{
   int kind;                 /* 0 csr, 1 constrained pa */
   int64_t n;
   const int64_t *rowptr; const int32_t *colind; const double *vals;
   int dim, p; int64_t ne;
   const int32_t *gather, *offsets, *indices;
   const double *Dd, *Dc, *Dm;
   const uint8_t *ess;
   double *work;
};

orc_op *orc_op_csr(int64_t n, const int64_t *rowptr, const int32_t *colind, const double *vals)
{
   orc_op *op = calloc(1, sizeof(*op));
   op->kind = 0; op->n = n; op->rowptr = rowptr; op->colind = colind; op->vals = vals;
   return op;
}

orc_op *orc_op_pa(int dim, int p, int64_t ne, int64_t ndof,
                  const int32_t *gather, const int32_t *offsets, const int32_t *indices,
                  const double *Ddiff, const double *Dconv, const double *Dmass,
                  const uint8_t *ess_mark)
{
   orc_op *op = calloc(1, sizeof(*op));
   op->kind = 1; op->n = ndof; op->dim = dim; op->p = p; op->ne = ne;
   op->gather = gather; op->offsets = offsets; op->indices = indices;
   op->Dd = Ddiff; op->Dc = Dconv; op->Dm = Dmass; op->ess = ess_mark;
   op->work = malloc(sizeof(double) * (ndof ? ndof : 1));
   return op;
}

void orc_op_free(orc_op *op) { if (op) { free(op->work); free(op); } }
int64_t orc_op_size(const orc_op *op) { return op->n; }

/* ConstrainedOperator::Mult (DIAG_ONE): z = x; z[ess] = 0; y = A z; y[ess] = x[ess] */
void orc_op_mult(const orc_op *op, const double *x, double *y)
{
   if (op->kind == 0) { orc_csr_spmv(op->n, op->rowptr, op->colind, op->vals, x, y); return; }
   const double *in = x;
   if (op->ess)
   {
      for (int64_t i = 0; i < op->n; i++) { op->work[i] = op->ess[i] ? 0.0 : x[i]; }
      in = op->work;
   }
   orc_pa_apply(op->dim, op->p, op->ne, op->n, op->gather, op->offsets, op->indices,
                op->Dd, op->Dc, op->Dm, in, y);
   if (op->ess)
      for (int64_t i = 0; i < op->n; i++) if (op->ess[i]) { y[i] = x[i]; }
}

void orc_op_eliminate_rhs(const orc_op *op, const double *x, double *b)
{
   if (op->kind != 1 || !op->ess) { return; }
   double *w = malloc(sizeof(double) * op->n), *t = malloc(sizeof(double) * op->n);
   for (int64_t i = 0; i < op->n; i++) { w[i] = op->ess[i] ? x[i] : 0.0; }
   orc_pa_apply(op->dim, op->p, op->ne, op->n, op->gather, op->offsets, op->indices,
                op->Dd, op->Dc, op->Dm, w, t);
   for (int64_t i = 0; i < op->n; i++) { b[i] = op->ess[i] ? x[i] : b[i] - t[i]; }
   free(w); free(t);
}

static double dot(int64_t n, const double *a, const double *b)
{
   double s = 0.0;
   #pragma omp parallel for reduction(+:s) schedule(static)
   for (int64_t i = 0; i < n; i++) { s += a[i] * b[i]; }
   return s;
}

/* ---- ILU(0), natural ordering, sequential IKJ: the reference's -pc_type bjacobi -sub_pc_type ilu on one rank
   (Input/petsc_circle.opts:6-8; PETSc PCILU defaults: 0 levels of fill).  [PETSc-upstream: not vendored] */
struct orc_ilu { int64_t n; const int64_t *rowptr; const int32_t *colind; double *lu; int64_t *diag; };

orc_ilu *orc_ilu0_factor(int64_t n, const int64_t *rowptr, const int32_t *colind, const double *vals)
{
   orc_ilu *f = calloc(1, sizeof(*f));
   f->n = n; f->rowptr = rowptr; f->colind = colind;
   f->lu = malloc(sizeof(double) * (size_t)rowptr[n]);
   f->diag = malloc(sizeof(int64_t) * (size_t)n);
   memcpy(f->lu, vals, sizeof(double) * (size_t)rowptr[n]);
   for (int64_t i = 0; i < n; i++)
      for (int64_t p = rowptr[i]; p < rowptr[i + 1]; p++) if (colind[p] == i) { f->diag[i] = p; }
   for (int64_t i = 0; i < n; i++)
   {
      for (int64_t p = rowptr[i]; p < f->diag[i]; p++)
      {
         const int32_t k = colind[p];
         const double lik = f->lu[p] / f->lu[f->diag[k]];
         f->lu[p] = lik;
         int64_t s = f->diag[k] + 1;                       /* merge the tails of rows i and k (both sorted) */
         for (int64_t q = p + 1; q < rowptr[i + 1]; q++)
         {
            while (s < rowptr[k + 1] && colind[s] < colind[q]) { s++; }
            if (s < rowptr[k + 1] && colind[s] == colind[q]) { f->lu[q] -= lik * f->lu[s]; }
         }
      }
   }
   return f;
}

void orc_ilu0_solve(const orc_ilu *f, const double *r, double *z)
{
   for (int64_t i = 0; i < f->n; i++)
   {
      double s = r[i];
      for (int64_t p = f->rowptr[i]; p < f->diag[i]; p++) { s -= f->lu[p] * z[f->colind[p]]; }
      z[i] = s;
   }
   for (int64_t i = f->n - 1; i >= 0; i--)
   {
      double s = z[i];
      for (int64_t p = f->diag[i] + 1; p < f->rowptr[i + 1]; p++) { s -= f->lu[p] * z[f->colind[p]]; }
      z[i] = s / f->lu[f->diag[i]];
   }
}

void orc_ilu0_get(const orc_ilu *f, double *lu) { memcpy(lu, f->lu, sizeof(double) * (size_t)f->rowptr[f->n]); }
void orc_ilu0_free(orc_ilu *f) { if (f) { free(f->lu); free(f->diag); free(f); } }

static const orc_ilu *g_ilu = NULL;      /* preconditioner of the solve in progress (orc_gmres_ilu) */

static void pc(int64_t n, const double *dinv, const double *r, double *z)
{
   if (g_ilu) { orc_ilu0_solve(g_ilu, r, z); return; }
   if (!dinv) { memcpy(z, r, sizeof(double) * n); return; }
   #pragma omp parallel for schedule(static)
   for (int64_t i = 0; i < n; i++) { z[i] = dinv[i] * r[i]; }
}

/* Left-preconditioned restarted GMRES on M^{-1}A x = M^{-1}b; the monitored
   quantity is the preconditioned residual norm |s_{j+1}| (KSP_NORM_PRECONDITIONED).
   Stop when rnorm <= max(rtol * rnorm0, atol) (KSPConvergedDefault). */
void orc_gmres(const orc_op *A, const double *dinv, const double *b, double *x,
               const orc_krylov_opts *o, orc_krylov_result *res, double *hist)
{
   const int64_t n = A->n;
   const int m = o->restart > 0 ? o->restart : (o->variant == 0 ? 30 : 50);
   double *V = malloc(sizeof(double) * (size_t)(m + 1) * n);
   double *w = malloc(sizeof(double) * n), *t = malloc(sizeof(double) * n);
   double *H = calloc((size_t)(m + 1) * m, sizeof(double));   /* column-major, ld = m+1 */
   double *cs = malloc(sizeof(double) * m), *sn = malloc(sizeof(double) * m);
   double *s = malloc(sizeof(double) * (m + 1)), *yv = malloc(sizeof(double) * m);
   double *hc = malloc(sizeof(double) * (m + 1));
   int it = 0, conv = 0, hl = 0;
   double rnorm = 0.0, ttol = 0.0;
   if (o->zero_guess) { memset(x, 0, sizeof(double) * n); }
   int first = 1;
   while (1)
   {
      /* r = M^{-1}(b - A x) */
      if (first && o->zero_guess) { pc(n, dinv, b, V); }
      else
      {
         orc_op_mult(A, x, t);
         for (int64_t i = 0; i < n; i++) { t[i] = b[i] - t[i]; }
         pc(n, dinv, t, V);
      }
      double beta = sqrt(dot(n, V, V));
      rnorm = beta;
      if (first)
      {
         /* KSPConvergedDefault: with a non-zero initial guess PETSc scales rtol by the preconditioned norm
            of the right-hand side, ||M^{-1} b||; mfem::GMRESSolver always uses the initial residual */
         double ref = beta;
         if (!o->zero_guess && o->variant == 0) { pc(n, dinv, b, w); ref = sqrt(dot(n, w, w)); }
         ttol = fmax(o->rtol * ref, o->atol);
         hist[hl++] = beta;
         first = 0;
      }
      if (rnorm <= ttol) { conv = 1; break; }
      if (it >= o->max_it) { break; }
      for (int64_t i = 0; i < n; i++) { V[i] /= beta; }
      memset(s, 0, sizeof(double) * (m + 1));
      s[0] = beta;
      int j;
      for (j = 0; j < m && it < o->max_it; )
      {
         double *vj = V + (size_t)j * n, *vn = V + (size_t)(j + 1) * n;
         orc_op_mult(A, vj, t);
         pc(n, dinv, t, w);
         if (o->variant == 0)
         {  /* classical Gram-Schmidt, no refinement: all dots against the same w */
            for (int i = 0; i <= j; i++) { hc[i] = dot(n, w, V + (size_t)i * n); }
            for (int i = 0; i <= j; i++)
            {
               const double *vi = V + (size_t)i * n; double h = hc[i];
               for (int64_t k = 0; k < n; k++) { w[k] -= h * vi[k]; }
            }
         }
         else
         {  /* modified Gram-Schmidt */
            for (int i = 0; i <= j; i++)
            {
               const double *vi = V + (size_t)i * n;
               double h = dot(n, w, vi); hc[i] = h;
               for (int64_t k = 0; k < n; k++) { w[k] -= h * vi[k]; }
            }
         }
         double hn = sqrt(dot(n, w, w));
         hc[j + 1] = hn;
         if (hn != 0.0) { for (int64_t k = 0; k < n; k++) { vn[k] = w[k] / hn; } }
         /* previous rotations, then the new one */
         for (int i = 0; i < j; i++)
         {
            double a = cs[i] * hc[i] + sn[i] * hc[i + 1];
            hc[i + 1] = -sn[i] * hc[i] + cs[i] * hc[i + 1];
            hc[i] = a;
         }
         double den = hypot(hc[j], hc[j + 1]);
         cs[j] = hc[j] / den; sn[j] = hc[j + 1] / den;
         hc[j] = den; hc[j + 1] = 0.0;
         s[j + 1] = -sn[j] * s[j];
         s[j] = cs[j] * s[j];
         for (int i = 0; i <= j; i++) { H[(size_t)j * (m + 1) + i] = hc[i]; }
         j++; it++;
         rnorm = fabs(s[j]);
         hist[hl++] = rnorm;
         if (rnorm <= ttol) { conv = 1; break; }
         if (hn == 0.0) { break; }
      }
      /* back substitution and update x += V y */
      for (int i = j - 1; i >= 0; i--)
      {
         double a = s[i];
         for (int k = i + 1; k < j; k++) { a -= H[(size_t)k * (m + 1) + i] * yv[k]; }
         yv[i] = a / H[(size_t)i * (m + 1) + i];
      }
      for (int i = 0; i < j; i++)
      {
         const double *vi = V + (size_t)i * n; double a = yv[i];
         for (int64_t k = 0; k < n; k++) { x[k] += a * vi[k]; }
      }
      if (conv || it >= o->max_it) { break; }
   }
   res->iters = it; res->converged = conv; res->final_norm = rnorm; res->hist_len = hl;
   free(V); free(w); free(t); free(H); free(cs); free(sn); free(s); free(yv); free(hc);
}

/* the same GMRES with z = (LU)^{-1} r as the left preconditioner */
void orc_gmres_ilu(const orc_op *A, const orc_ilu *ilu, const double *b, double *x,
                   const orc_krylov_opts *o, orc_krylov_result *res, double *hist)
{
   g_ilu = ilu;
   orc_gmres(A, NULL, b, x, o, res, hist);
   g_ilu = NULL;
}

/* mfem::CGSolver::Mult (linalg/solvers.cpp), SURVEY.md C.6 */
void orc_cg(const orc_op *A, const double *dinv, const double *b, double *x,
            const orc_krylov_opts *o, orc_krylov_result *res, double *hist)
{
   const int64_t n = A->n;
   double *r = malloc(sizeof(double) * n), *d = malloc(sizeof(double) * n), *z = malloc(sizeof(double) * n);
   int hl = 0, conv = 0, it = 0;
   if (o->zero_guess) { memset(x, 0, sizeof(double) * n); memcpy(r, b, sizeof(double) * n); }
   else
   {
      orc_op_mult(A, x, r);
      for (int64_t i = 0; i < n; i++) { r[i] = b[i] - r[i]; }
   }
   pc(n, dinv, r, z);
   memcpy(d, z, sizeof(double) * n);
   double nom = dot(n, d, r), nom0 = nom;
   double r0 = fmax(nom * o->rtol * o->rtol, o->atol * o->atol);
   hist[hl++] = nom;
   double betanom = nom;
   if (nom <= r0) { conv = 1; }
   else
   {
      orc_op_mult(A, d, z);
      double den = dot(n, z, d);
      if (den > 0.0)
      {
         for (it = 1; ; it++)
         {
            double alpha = nom / den;
            for (int64_t i = 0; i < n; i++) { x[i] += alpha * d[i]; r[i] -= alpha * z[i]; }
            if (dinv) { pc(n, dinv, r, z); betanom = dot(n, r, z); }
            else { betanom = dot(n, r, r); }
            hist[hl++] = betanom;
            if (betanom <= r0) { conv = 1; break; }
            if (it >= o->max_it) { break; }
            double beta = betanom / nom;
            if (dinv) { for (int64_t i = 0; i < n; i++) { d[i] = z[i] + beta * d[i]; } }
            else { for (int64_t i = 0; i < n; i++) { d[i] = r[i] + beta * d[i]; } }
            orc_op_mult(A, d, z);
            den = dot(n, d, z);
            if (den <= 0.0) { break; }
            nom = betanom;
         }
      }
   }
   (void)nom0;
   res->iters = it; res->converged = conv; res->final_norm = sqrt(fabs(betanom)); res->hist_len = hl;
   free(r); free(d); free(z);
}
