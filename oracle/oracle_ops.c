/* oracle_ops.c -- CPU ORACLE (test infrastructure): quadrature data, the
 * sum-factorised partial-assembly apply (MFEM three-pass structure), its
 * diagonal, and the fully assembled CSR path the reference app executes.
 *
 * Reference call sites restated (file:line relative to
 * /root/reference/myapps/convection_diffusion/):
 *   linear_convection_diffusion_2D.cpp:335-339  Diffusion+Convection+Mass, Assemble
 *   linear_convection_diffusion_1D.cpp:391-400  Mass + Convection(beta,dt) + Diffusion(dt/Pe)
 *   linear_convection_diffusion_1D.cpp:544      mass_form.Mult
 *   linear_convection_diffusion_2D.cpp:351      FormLinearSystem (elimination)
 * Upstream MFEM definitions: SURVEY.md Appendix C.4/C.5.  PARITY UNPINNED.
 */
#include "cdm_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXD 8   /* max D1D */
#define MAXQ 9   /* max Q1D */

static int ipow(int b, int e) { int r = 1; while (e-- > 0) { r *= b; } return r; }

int orc_num_threads(void)
{
#ifdef _OPENMP
   return omp_get_max_threads();
#else
   return 1;
#endif
}
void orc_set_num_threads(int n)
{
#ifdef _OPENMP
   omp_set_num_threads(n);
#else
   (void)n;
#endif
}

/* ----------------------------------------------------- geometric factors */

/* Jacobian of the (bi/tri)linear map at reference point xi; J[a*dim+b] = dX_a/dxi_b */
static void jacobian(int dim, const double *X /* nvpe x dim */, const double *xi, double *J)
{
   if (dim == 2)
   {
      double x = xi[0], y = xi[1];
      double dN[4][2] = { { -(1 - y), -(1 - x) }, { (1 - y), -x }, { y, x }, { -y, (1 - x) } };
      for (int a = 0; a < 2; a++)
         for (int b = 0; b < 2; b++)
         {
            double s = 0.0;
            for (int k = 0; k < 4; k++) { s += X[2 * k + a] * dN[k][b]; }
            J[2 * a + b] = s;
         }
      return;
   }
   double x = xi[0], y = xi[1], z = xi[2];
   double dN[8][3] =
   {
      { -(1 - y) * (1 - z), -(1 - x) * (1 - z), -(1 - x) * (1 - y) },
      {  (1 - y) * (1 - z), -x * (1 - z),       -x * (1 - y) },
      {  y * (1 - z),        x * (1 - z),       -x * y },
      { -y * (1 - z),        (1 - x) * (1 - z), -(1 - x) * y },
      { -(1 - y) * z,       -(1 - x) * z,        (1 - x) * (1 - y) },
      {  (1 - y) * z,       -x * z,              x * (1 - y) },
      {  y * z,              x * z,              x * y },
      { -y * z,              (1 - x) * z,        (1 - x) * y }
   };
   for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++)
      {
         double s = 0.0;
         for (int k = 0; k < 8; k++) { s += X[3 * k + a] * dN[k][b]; }
         J[3 * a + b] = s;
      }
}

/* adjugate A (J*A = det*I) and determinant */
static double adjugate(int dim, const double *J, double *A)
{
   if (dim == 2)
   {
      A[0] = J[3]; A[1] = -J[1]; A[2] = -J[2]; A[3] = J[0];
      return J[0] * J[3] - J[1] * J[2];
   }
   A[0] = J[4] * J[8] - J[5] * J[7];
   A[1] = J[7] * J[2] - J[1] * J[8];
   A[2] = J[1] * J[5] - J[4] * J[2];
   A[3] = J[6] * J[5] - J[3] * J[8];
   A[4] = J[0] * J[8] - J[2] * J[6];
   A[5] = J[3] * J[2] - J[0] * J[5];
   A[6] = J[3] * J[7] - J[6] * J[4];
   A[7] = J[6] * J[1] - J[0] * J[7];
   A[8] = J[0] * J[4] - J[1] * J[3];
   return J[0] * A[0] + J[3] * A[1] + J[6] * A[2];
}

/* symmetric-matrix component (r,c) from packed storage: 2D (11,21,22), 3D (11,21,31,22,32,33) */
static int symidx(int dim, int r, int c)
{
   if (r < c) { int t = r; r = c; c = t; }
   if (dim == 2) { return (c == 0) ? r : 2; }
   if (c == 0) { return r; }
   if (c == 1) { return 2 + r; }
   return 5;
}

static void elem_vertices(int dim, const int32_t *ev, const double *vx, int64_t e, double *X)
{
   int nvpe = (dim == 2) ? 4 : 8;
   for (int k = 0; k < nvpe; k++)
      for (int c = 0; c < dim; c++) { X[dim * k + c] = vx[(int64_t)ev[nvpe * e + k] * dim + c]; }
}

void orc_qdata(int dim, int p, int64_t ne, const int32_t *ev, const double *vx,
               int kappa_kind, int kappa_ncomp, const double *kappa,
               int vel_kind, const double *vel, double alpha,
               int mass_kind, const double *mass,
               double *Ddiff, double *Dconv, double *Dmass)
{
   const int q1d = orc_q1d(dim, p), nq = ipow(q1d, dim), nsym = dim * (dim + 1) / 2;
   double xq[MAXQ], wq[MAXQ];
   orc_gauss_legendre(q1d, xq, wq);
   #pragma omp parallel for schedule(static)
   for (int64_t e = 0; e < ne; e++)
   {
      double X[24];
      elem_vertices(dim, ev, vx, e, X);
      for (int q = 0; q < nq; q++)
      {
         int qx = q % q1d, qy = (q / q1d) % q1d, qz = (dim == 3) ? q / (q1d * q1d) : 0;
         double xi[3] = { xq[qx], xq[qy], (dim == 3) ? xq[qz] : 0.0 };
         double w = wq[qx] * wq[qy] * ((dim == 3) ? wq[qz] : 1.0);
         double J[9], A[9];
         jacobian(dim, X, xi, J);
         double det = adjugate(dim, J, A);
         if (Ddiff && kappa_kind)
         {
            double M[9];
            const double *kp = (kappa_kind == 1) ? kappa : kappa + (e * nq + q) * kappa_ncomp;
            for (int r = 0; r < dim; r++)
               for (int c = 0; c < dim; c++)
                  M[dim * r + c] = (kappa_ncomp == 1) ? ((r == c) ? kp[0] : 0.0) : kp[symidx(dim, r, c)];
            /* (w/det) A M A^T */
            for (int r = 0; r < dim; r++)
               for (int c = 0; c <= r; c++)
               {
                  double s = 0.0;
                  for (int i = 0; i < dim; i++)
                     for (int j = 0; j < dim; j++) { s += A[dim * r + i] * M[dim * i + j] * A[dim * c + j]; }
                  Ddiff[(e * nsym + symidx(dim, r, c)) * nq + q] = w / det * s;
               }
         }
         if (Dconv && vel_kind)
         {
            const double *vp = (vel_kind == 1) ? vel : vel + (e * nq + q) * dim;
            for (int r = 0; r < dim; r++)
            {
               double s = 0.0;
               for (int j = 0; j < dim; j++) { s += A[dim * r + j] * vp[j]; }
               Dconv[(e * dim + r) * nq + q] = alpha * w * s;
            }
         }
         if (Dmass && mass_kind)
         {
            double m = (mass_kind == 1) ? mass[0] : mass[e * nq + q];
            Dmass[e * nq + q] = w * m * det;
         }
      }
   }
}

/* ------------------------------------------------ sum-factorised PA apply */

/* contract direction `dir` of a tensor with extents n[0..dim) (x fastest):
   out[.., o, ..] = sum_i M[o*ni + i] in[.., i, ..]   (or M^T when trans) */
static void contract(int dim, const int *n, int dir, int no, const double *M, int trans,
                     const double *in, double *out)
{
   int ni = n[dir];
   int inner = 1, outer = 1;
   for (int d = 0; d < dir; d++) { inner *= n[d]; }
   for (int d = dir + 1; d < dim; d++) { outer *= n[d]; }
   for (int a = 0; a < outer; a++)
      for (int o = 0; o < no; o++)
         for (int b = 0; b < inner; b++)
         {
            double s = 0.0;
            for (int i = 0; i < ni; i++)
            {
               double m = trans ? M[i * no + o] : M[o * ni + i];
               s += m * in[(a * ni + i) * inner + b];
            }
            out[(a * no + o) * inner + b] = s;
         }
}

/* apply a chain of 1-D operators: op[d] in {B,G}; forward (D1D->Q1D) or transposed */
static void tensor_apply(int dim, int d1d, int q1d, const double *B, const double *G,
                         int gdir /* -1: all B */, int transpose,
                         const double *in, double *out, double *t0, double *t1)
{
   int n[3];
   for (int d = 0; d < dim; d++) { n[d] = transpose ? q1d : d1d; }
   const double *src = in;
   for (int d = 0; d < dim; d++)
   {
      double *dst = (d == dim - 1) ? out : ((d & 1) ? t1 : t0);
      const double *M = (d == gdir) ? G : B;
      contract(dim, n, d, transpose ? d1d : q1d, M, transpose, src, dst);
      n[d] = transpose ? d1d : q1d;
      src = dst;
   }
}

/* one element: yE += (B^T D B) xE, the three integrators as separate passes
   (DiffusionIntegrator::AddMultPA, ConvectionIntegrator::AddMultPA,
   MassIntegrator::AddMultPA in upstream MFEM). */
static void elem_apply(int dim, int d1d, int q1d, const double *B, const double *G,
                       const double *Dd, const double *Dc, const double *Dm,
                       const double *xE, double *yE)
{
   const int nd = ipow(d1d, dim), nq = ipow(q1d, dim);
   double u[MAXQ * MAXQ * MAXQ], g[3][MAXQ * MAXQ * MAXQ], f[MAXQ * MAXQ * MAXQ];
   double t0[MAXQ * MAXQ * MAXQ], t1[MAXQ * MAXQ * MAXQ], out[MAXD * MAXD * MAXD];
   for (int i = 0; i < nd; i++) { yE[i] = 0.0; }
   if (Dd || Dc)
      for (int c = 0; c < dim; c++) { tensor_apply(dim, d1d, q1d, B, G, c, 0, xE, g[c], t0, t1); }
   if (Dd)
   {
      for (int r = 0; r < dim; r++)
      {
         for (int q = 0; q < nq; q++)
         {
            double s = 0.0;
            for (int c = 0; c < dim; c++) { s += Dd[symidx(dim, r, c) * nq + q] * g[c][q]; }
            f[q] = s;
         }
         tensor_apply(dim, d1d, q1d, B, G, r, 1, f, out, t0, t1);
         for (int i = 0; i < nd; i++) { yE[i] += out[i]; }
      }
   }
   if (Dc)
   {
      for (int q = 0; q < nq; q++)
      {
         double s = 0.0;
         for (int c = 0; c < dim; c++) { s += Dc[c * nq + q] * g[c][q]; }
         f[q] = s;
      }
      tensor_apply(dim, d1d, q1d, B, G, -1, 1, f, out, t0, t1);
      for (int i = 0; i < nd; i++) { yE[i] += out[i]; }
   }
   if (Dm)
   {
      tensor_apply(dim, d1d, q1d, B, G, -1, 0, xE, u, t0, t1);
      for (int q = 0; q < nq; q++) { f[q] = Dm[q] * u[q]; }
      tensor_apply(dim, d1d, q1d, B, G, -1, 1, f, out, t0, t1);
      for (int i = 0; i < nd; i++) { yE[i] += out[i]; }
   }
}

void orc_pa_apply(int dim, int p, int64_t ne, int64_t ndof,
                  const int32_t *gather, const int32_t *offsets, const int32_t *indices,
                  const double *Ddiff, const double *Dconv, const double *Dmass,
                  const double *x, double *y)
{
   const int d1d = p + 1, q1d = orc_q1d(dim, p), nd = ipow(d1d, dim), nq = ipow(q1d, dim);
   const int nsym = dim * (dim + 1) / 2;
   double B[MAXQ * MAXD], G[MAXQ * MAXD], qw[MAXQ];
   orc_basis(p, q1d, B, G, qw);
   double *yE = malloc(sizeof(double) * (size_t)ne * nd);
   #pragma omp parallel for schedule(static)
   for (int64_t e = 0; e < ne; e++)
   {
      double xE[MAXD * MAXD * MAXD];
      for (int i = 0; i < nd; i++) { xE[i] = x[gather[e * nd + i]]; }   /* G */
      elem_apply(dim, d1d, q1d, B, G,
                 Ddiff ? Ddiff + e * nsym * nq : NULL,
                 Dconv ? Dconv + e * dim * nq : NULL,
                 Dmass ? Dmass + e * nq : NULL, xE, yE + e * nd);
   }
   #pragma omp parallel for schedule(static)
   for (int64_t g = 0; g < ndof; g++)                                   /* G^T */
   {
      double s = 0.0;
      for (int32_t j = offsets[g]; j < offsets[g + 1]; j++) { s += yE[indices[j]]; }
      y[g] = s;
   }
   free(yE);
}

void orc_pa_diag(int dim, int p, int64_t ne, int64_t ndof,
                 const int32_t *gather, const int32_t *offsets, const int32_t *indices,
                 const double *Ddiff, const double *Dconv, const double *Dmass,
                 double *diag)
{
   (void)gather;
   const int d1d = p + 1, q1d = orc_q1d(dim, p), nd = ipow(d1d, dim), nq = ipow(q1d, dim);
   const int nsym = dim * (dim + 1) / 2;
   double B[MAXQ * MAXD], G[MAXQ * MAXD], qw[MAXQ];
   orc_basis(p, q1d, B, G, qw);
   double *dE = malloc(sizeof(double) * (size_t)ne * nd);
   #pragma omp parallel for schedule(static)
   for (int64_t e = 0; e < ne; e++)
      for (int l = 0; l < nd; l++)
      {
         int lx = l % d1d, ly = (l / d1d) % d1d, lz = (dim == 3) ? l / (d1d * d1d) : 0;
         double s = 0.0;
         for (int q = 0; q < nq; q++)
         {
            int qx = q % q1d, qy = (q / q1d) % q1d, qz = (dim == 3) ? q / (q1d * q1d) : 0;
            double bx = B[qx * d1d + lx], by = B[qy * d1d + ly], bz = (dim == 3) ? B[qz * d1d + lz] : 1.0;
            double gx = G[qx * d1d + lx], gy = G[qy * d1d + ly], gz = (dim == 3) ? G[qz * d1d + lz] : 0.0;
            double phi = bx * by * bz;
            double gr[3] = { gx * by * bz, bx * gy * bz, bx * by * gz };
            if (Ddiff)
               for (int r = 0; r < dim; r++)
                  for (int c = 0; c < dim; c++)
                     s += Ddiff[(e * nsym + symidx(dim, r, c)) * nq + q] * gr[r] * gr[c];
            if (Dconv)
               for (int c = 0; c < dim; c++) { s += phi * Dconv[(e * dim + c) * nq + q] * gr[c]; }
            if (Dmass) { s += Dmass[e * nq + q] * phi * phi; }
         }
         dE[e * nd + l] = s;
      }
   for (int64_t g = 0; g < ndof; g++)
   {
      double s = 0.0;
      for (int32_t j = offsets[g]; j < offsets[g + 1]; j++) { s += dE[indices[j]]; }
      diag[g] = s;
   }
   free(dE);
}

/* -------------------------------------------------------- full assembly */

static int cmp_i32(const void *a, const void *b)
{
   int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
   return (x > y) - (x < y);
}

int64_t orc_csr_pattern(int64_t ne, int nd, int64_t ndof, const int32_t *elem_dof,
                        int64_t *rowptr, int32_t *colind)
{
   int32_t *off = malloc(sizeof(int32_t) * (ndof + 1));
   int32_t *ind = malloc(sizeof(int32_t) * (size_t)ne * nd);
   orc_restriction(ne, nd, ndof, elem_dof, off, ind);
   int64_t nnz = 0;
   int cap = 64 * nd;
   int32_t *buf = malloc(sizeof(int32_t) * cap);
   rowptr[0] = 0;
   for (int64_t g = 0; g < ndof; g++)
   {
      int cnt = 0;
      int need = (off[g + 1] - off[g]) * nd;
      if (need > cap) { cap = need; buf = realloc(buf, sizeof(int32_t) * cap); }
      for (int32_t j = off[g]; j < off[g + 1]; j++)
      {
         int64_t e = ind[j] / nd;
         for (int i = 0; i < nd; i++) { buf[cnt++] = elem_dof[e * nd + i]; }
      }
      qsort(buf, cnt, sizeof(int32_t), cmp_i32);
      int u = 0;
      for (int i = 0; i < cnt; i++)
         if (i == 0 || buf[i] != buf[i - 1])
         {
            if (colind) { colind[nnz + u] = buf[i]; }
            u++;
         }
      nnz += u;
      rowptr[g + 1] = nnz;
   }
   free(buf); free(off); free(ind);
   return nnz;
}

/* element matrix in the style of {Diffusion,Convection,Mass}Integrator::
   AssembleElementMatrix: full shape / physical-gradient vectors per point. */
static void elem_matrix(int dim, int d1d, int q1d, const double *B, const double *G, const double *qw,
                        const double *xq, const double *X, int64_t e, int nq,
                        int kappa_kind, int kappa_ncomp, const double *kappa,
                        int vel_kind, const double *vel, double alpha,
                        int mass_kind, const double *mass, double *elmat)
{
   const int nd = ipow(d1d, dim);
   double shape[MAXD * MAXD * MAXD], dshape[MAXD * MAXD * MAXD][3], dphys[MAXD * MAXD * MAXD][3];
   for (int i = 0; i < nd * nd; i++) { elmat[i] = 0.0; }
   for (int q = 0; q < nq; q++)
   {
      int qx = q % q1d, qy = (q / q1d) % q1d, qz = (dim == 3) ? q / (q1d * q1d) : 0;
      double xi[3] = { xq[qx], xq[qy], (dim == 3) ? xq[qz] : 0.0 };
      double w = qw[qx] * qw[qy] * ((dim == 3) ? qw[qz] : 1.0);
      double J[9], A[9];
      jacobian(dim, X, xi, J);
      double det = adjugate(dim, J, A);
      for (int l = 0; l < nd; l++)
      {
         int lx = l % d1d, ly = (l / d1d) % d1d, lz = (dim == 3) ? l / (d1d * d1d) : 0;
         double bx = B[qx * d1d + lx], by = B[qy * d1d + ly], bz = (dim == 3) ? B[qz * d1d + lz] : 1.0;
         double gx = G[qx * d1d + lx], gy = G[qy * d1d + ly], gz = (dim == 3) ? G[qz * d1d + lz] : 0.0;
         shape[l] = bx * by * bz;
         dshape[l][0] = gx * by * bz; dshape[l][1] = bx * gy * bz; dshape[l][2] = bx * by * gz;
         /* physical gradient = J^{-T} dshape = A^T dshape / det */
         for (int a = 0; a < dim; a++)
         {
            double s = 0.0;
            for (int b = 0; b < dim; b++) { s += A[dim * b + a] * dshape[l][b]; }
            dphys[l][a] = s / det;
         }
      }
      double M[9] = { 0 }, v[3] = { 0, 0, 0 }, ms = 0.0;
      if (kappa_kind)
      {
         const double *kp = (kappa_kind == 1) ? kappa : kappa + (e * nq + q) * kappa_ncomp;
         for (int r = 0; r < dim; r++)
            for (int c = 0; c < dim; c++)
               M[dim * r + c] = (kappa_ncomp == 1) ? ((r == c) ? kp[0] : 0.0) : kp[symidx(dim, r, c)];
      }
      if (vel_kind)
      {
         const double *vp = (vel_kind == 1) ? vel : vel + (e * nq + q) * dim;
         for (int c = 0; c < dim; c++) { v[c] = vp[c]; }
      }
      if (mass_kind) { ms = (mass_kind == 1) ? mass[0] : mass[e * nq + q]; }
      const double wd = w * det;
      for (int i = 0; i < nd; i++)
      {
         double Mgi[3] = { 0, 0, 0 };   /* grad(v_i)^T M */
         for (int c = 0; c < dim; c++)
            for (int r = 0; r < dim; r++) { Mgi[c] += dphys[i][r] * M[dim * r + c]; }
         for (int j = 0; j < nd; j++)
         {
            double s = 0.0;
            if (kappa_kind) { for (int c = 0; c < dim; c++) { s += Mgi[c] * dphys[j][c]; } }
            if (vel_kind)
            {
               double cv = 0.0;
               for (int c = 0; c < dim; c++) { cv += v[c] * dphys[j][c]; }
               s += alpha * cv * shape[i];
            }
            if (mass_kind) { s += ms * shape[i] * shape[j]; }
            elmat[i * nd + j] += wd * s;
         }
      }
   }
}

void orc_csr_assemble(int dim, int p, int64_t ne, int64_t ndof, const int32_t *ev,
                      const double *vx, const int32_t *elem_dof,
                      int kappa_kind, int kappa_ncomp, const double *kappa,
                      int vel_kind, const double *vel, double alpha,
                      int mass_kind, const double *mass,
                      const int64_t *rowptr, const int32_t *colind, double *vals)
{
   const int d1d = p + 1, q1d = orc_q1d(dim, p), nd = ipow(d1d, dim), nq = ipow(q1d, dim);
   double B[MAXQ * MAXD], G[MAXQ * MAXD], qw[MAXQ], xq[MAXQ], wtmp[MAXQ];
   orc_basis(p, q1d, B, G, qw);
   orc_gauss_legendre(q1d, xq, wtmp);
   for (int64_t i = 0; i < rowptr[ndof]; i++) { vals[i] = 0.0; }
   const int64_t chunk = 512;
   double *mats = malloc(sizeof(double) * (size_t)chunk * nd * nd);
   for (int64_t e0 = 0; e0 < ne; e0 += chunk)
   {
      int64_t e1 = (e0 + chunk < ne) ? e0 + chunk : ne;
      #pragma omp parallel for schedule(static)
      for (int64_t e = e0; e < e1; e++)
      {
         double X[24];
         elem_vertices(dim, ev, vx, e, X);
         elem_matrix(dim, d1d, q1d, B, G, qw, xq, X, e, nq, kappa_kind, kappa_ncomp, kappa,
                     vel_kind, vel, alpha, mass_kind, mass, mats + (e - e0) * nd * nd);
      }
      /* serial, element-ordered scatter: deterministic summation order */
      for (int64_t e = e0; e < e1; e++)
      {
         const double *m = mats + (e - e0) * nd * nd;
         const int32_t *dofs = elem_dof + e * nd;
         for (int i = 0; i < nd; i++)
         {
            int64_t r0 = rowptr[dofs[i]], r1 = rowptr[dofs[i] + 1];
            for (int j = 0; j < nd; j++)
            {
               int32_t c = dofs[j];
               int64_t lo = r0, hi = r1 - 1;
               while (lo < hi)
               {
                  int64_t mid = (lo + hi) >> 1;
                  if (colind[mid] < c) { lo = mid + 1; } else { hi = mid; }
               }
               vals[lo] += m[i * nd + j];
            }
         }
      }
   }
   free(mats);
}

void orc_csr_spmv(int64_t n, const int64_t *rowptr, const int32_t *colind,
                  const double *vals, const double *x, double *y)
{
   #pragma omp parallel for schedule(static)
   for (int64_t i = 0; i < n; i++)
   {
      double s = 0.0;
      for (int64_t k = rowptr[i]; k < rowptr[i + 1]; k++) { s += vals[k] * x[colind[k]]; }
      y[i] = s;
   }
}

void orc_csr_diag(int64_t n, const int64_t *rowptr, const int32_t *colind,
                  const double *vals, double *d)
{
   for (int64_t i = 0; i < n; i++)
   {
      d[i] = 0.0;
      for (int64_t k = rowptr[i]; k < rowptr[i + 1]; k++)
         if (colind[k] == i) { d[i] = vals[k]; }
   }
}

void orc_csr_eliminate(int64_t n, const int64_t *rowptr, const int32_t *colind,
                       double *vals, const uint8_t *ess, const double *x, double *b)
{
   /* b_free -= A_fe x_e */
   for (int64_t i = 0; i < n; i++)
   {
      if (ess[i]) { continue; }
      double s = 0.0;
      for (int64_t k = rowptr[i]; k < rowptr[i + 1]; k++)
         if (ess[colind[k]]) { s += vals[k] * x[colind[k]]; vals[k] = 0.0; }
      b[i] -= s;
   }
   for (int64_t i = 0; i < n; i++)
   {
      if (!ess[i]) { continue; }
      double d = 0.0;
      for (int64_t k = rowptr[i]; k < rowptr[i + 1]; k++)
      {
         if (colind[k] == i) { d = vals[k]; } else { vals[k] = 0.0; }
      }
      b[i] = d * x[i];
   }
}

/* ------------------------------------------------------------------------------
 * Fast CPU partial-assembly apply (3D): one fused pass per element with compile-time
 * loop bounds (macro-instantiated per order), the shape an optimised CPU PA code has
 * (MFEM's SmemPA*Apply3D kernels fused over the three integrators).  Used for the
 * cpu_baseline timing of bench.py; checked against orc_pa_apply in tests/test_oracle.py.
 * ------------------------------------------------------------------------------ */
#define DEFINE_FAST_APPLY3D(D, Q)                                                                   \
static void fast_elem_apply3d_##D##_##Q(const double *restrict B, const double *restrict G,        \
                                        const double *restrict Dd, const double *restrict Dc,      \
                                        const double *restrict Dm, const double *restrict xE,      \
                                        double *restrict yE)                                       \
{                                                                                                   \
   enum { NQ = Q * Q * Q };                                                                         \
   double tB[D][D][Q], tG[D][D][Q];               /* [dz][dy][qx] */                                \
   double vBB[D][Q][Q], vGB[D][Q][Q], vBG[D][Q][Q];   /* [dz][qy][qx] */                            \
   double u[Q][Q][Q], ux[Q][Q][Q], uy[Q][Q][Q], uz[Q][Q][Q];                                        \
   for (int dz = 0; dz < D; dz++) for (int dy = 0; dy < D; dy++) for (int qx = 0; qx < Q; qx++)    \
   {                                                                                                \
      double a = 0.0, b = 0.0;                                                                      \
      for (int dx = 0; dx < D; dx++) { const double v = xE[dx + D * (dy + D * dz)]; a += B[qx * D + dx] * v; b += G[qx * D + dx] * v; } \
      tB[dz][dy][qx] = a; tG[dz][dy][qx] = b;                                                       \
   }                                                                                                \
   for (int dz = 0; dz < D; dz++) for (int qy = 0; qy < Q; qy++) for (int qx = 0; qx < Q; qx++)    \
   {                                                                                                \
      double a = 0.0, b = 0.0, c = 0.0;                                                             \
      for (int dy = 0; dy < D; dy++)                                                                \
      { a += B[qy * D + dy] * tB[dz][dy][qx]; b += B[qy * D + dy] * tG[dz][dy][qx]; c += G[qy * D + dy] * tB[dz][dy][qx]; } \
      vBB[dz][qy][qx] = a; vGB[dz][qy][qx] = b; vBG[dz][qy][qx] = c;                                \
   }                                                                                                \
   for (int qz = 0; qz < Q; qz++) for (int qy = 0; qy < Q; qy++) for (int qx = 0; qx < Q; qx++)    \
   {                                                                                                \
      double a = 0.0, b = 0.0, c = 0.0, d = 0.0;                                                    \
      for (int dz = 0; dz < D; dz++)                                                                \
      {                                                                                             \
         a += B[qz * D + dz] * vBB[dz][qy][qx]; b += B[qz * D + dz] * vGB[dz][qy][qx];              \
         c += B[qz * D + dz] * vBG[dz][qy][qx]; d += G[qz * D + dz] * vBB[dz][qy][qx];              \
      }                                                                                             \
      const int q = qx + Q * (qy + Q * qz);                                                         \
      double fx = 0.0, fy = 0.0, fz = 0.0, s = 0.0;                                                 \
      if (Dd)                                                                                       \
      {                                                                                             \
         fx = Dd[q] * b + Dd[NQ + q] * c + Dd[2 * NQ + q] * d;                                      \
         fy = Dd[NQ + q] * b + Dd[3 * NQ + q] * c + Dd[4 * NQ + q] * d;                             \
         fz = Dd[2 * NQ + q] * b + Dd[4 * NQ + q] * c + Dd[5 * NQ + q] * d;                         \
      }                                                                                             \
      if (Dc) { s = Dc[q] * b + Dc[NQ + q] * c + Dc[2 * NQ + q] * d; }                              \
      if (Dm) { s += Dm[q] * a; }                                                                   \
      u[qz][qy][qx] = s; ux[qz][qy][qx] = fx; uy[qz][qy][qx] = fy; uz[qz][qy][qx] = fz;             \
   }                                                                                                \
   /* transposed: z, y, x */                                                                        \
   for (int dz = 0; dz < D; dz++) for (int qy = 0; qy < Q; qy++) for (int qx = 0; qx < Q; qx++)    \
   {                                                                                                \
      double a = 0.0, b = 0.0, c = 0.0;                                                             \
      for (int qz = 0; qz < Q; qz++)                                                                \
      {                                                                                             \
         a += B[qz * D + dz] * ux[qz][qy][qx]; b += B[qz * D + dz] * uy[qz][qy][qx];                \
         c += G[qz * D + dz] * uz[qz][qy][qx] + B[qz * D + dz] * u[qz][qy][qx];                     \
      }                                                                                             \
      vBB[dz][qy][qx] = a; vGB[dz][qy][qx] = b; vBG[dz][qy][qx] = c;                                \
   }                                                                                                \
   for (int dz = 0; dz < D; dz++) for (int dy = 0; dy < D; dy++) for (int qx = 0; qx < Q; qx++)    \
   {                                                                                                \
      double a = 0.0, b = 0.0;                                                                      \
      for (int qy = 0; qy < Q; qy++)                                                                \
      { a += B[qy * D + dy] * vBB[dz][qy][qx]; b += G[qy * D + dy] * vGB[dz][qy][qx] + B[qy * D + dy] * vBG[dz][qy][qx]; } \
      tB[dz][dy][qx] = a; tG[dz][dy][qx] = b;                                                       \
   }                                                                                                \
   for (int dz = 0; dz < D; dz++) for (int dy = 0; dy < D; dy++) for (int dx = 0; dx < D; dx++)    \
   {                                                                                                \
      double a = 0.0;                                                                               \
      for (int qx = 0; qx < Q; qx++) { a += G[qx * D + dx] * tB[dz][dy][qx] + B[qx * D + dx] * tG[dz][dy][qx]; } \
      yE[dx + D * (dy + D * dz)] = a;                                                               \
   }                                                                                                \
}
DEFINE_FAST_APPLY3D(2, 3)
DEFINE_FAST_APPLY3D(3, 4)
DEFINE_FAST_APPLY3D(4, 5)
DEFINE_FAST_APPLY3D(5, 6)
DEFINE_FAST_APPLY3D(6, 7)
DEFINE_FAST_APPLY3D(7, 8)

void orc_pa_apply_fast(int dim, int p, int64_t ne, int64_t ndof,
                       const int32_t *gather, const int32_t *offsets, const int32_t *indices,
                       const double *Ddiff, const double *Dconv, const double *Dmass,
                       const double *x, double *y, double *yE_work)
{
   if (dim != 3 || p < 1 || p > 6)
   {
      orc_pa_apply(dim, p, ne, ndof, gather, offsets, indices, Ddiff, Dconv, Dmass, x, y);
      return;
   }
   const int d1d = p + 1, q1d = p + 2, nd = d1d * d1d * d1d, nq = q1d * q1d * q1d;
   double B[MAXQ * MAXD], G[MAXQ * MAXD], qw[MAXQ];
   orc_basis(p, q1d, B, G, qw);
   double *yE = yE_work ? yE_work : malloc(sizeof(double) * (size_t)ne * nd);
   #pragma omp parallel for schedule(static)
   for (int64_t e = 0; e < ne; e++)
   {
      double xE[MAXD * MAXD * MAXD];
      for (int i = 0; i < nd; i++) { xE[i] = x[gather[e * nd + i]]; }
      const double *dd = Ddiff ? Ddiff + e * 6 * nq : NULL, *dc = Dconv ? Dconv + e * 3 * nq : NULL;
      const double *dm = Dmass ? Dmass + e * nq : NULL;
      double *ye = yE + e * nd;
      switch (p)
      {
         case 1: fast_elem_apply3d_2_3(B, G, dd, dc, dm, xE, ye); break;
         case 2: fast_elem_apply3d_3_4(B, G, dd, dc, dm, xE, ye); break;
         case 3: fast_elem_apply3d_4_5(B, G, dd, dc, dm, xE, ye); break;
         case 4: fast_elem_apply3d_5_6(B, G, dd, dc, dm, xE, ye); break;
         case 5: fast_elem_apply3d_6_7(B, G, dd, dc, dm, xE, ye); break;
         default: fast_elem_apply3d_7_8(B, G, dd, dc, dm, xE, ye); break;
      }
   }
   #pragma omp parallel for schedule(static)
   for (int64_t g = 0; g < ndof; g++)
   {
      double s = 0.0;
      for (int32_t j = offsets[g]; j < offsets[g + 1]; j++) { s += yE[indices[j]]; }
      y[g] = s;
   }
   if (!yE_work) { free(yE); }
}
