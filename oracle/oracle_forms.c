/* oracle_forms.c -- CPU ORACLE (test infrastructure): the steps either side of
 * the solve -- linear-form assembly, nodal projection points and L2 error norms.
 *
 * Restates (upstream MFEM, not vendored under /root/reference -- SURVEY.md
 * Appendix C) what the reference calls at
 *   linear_convection_diffusion_2D.cpp:341-343  ParLinearForm + DomainLFIntegrator(f)
 *   diffusion_mms.cpp:433-437                   the same, once per time step
 *   linear_convection_diffusion_2D.cpp:383-392  ComputeL2Error / ComputeGlobalLpNorm
 *                                               with rules of order max(2, 2p+3)
 * MFEM conventions restated here:
 *   DomainLFIntegrator(Q, a=2, b=0) integrates with IntRules.Get(geom, a*p + b),
 *   i.e. order/2+1 = p+1 Gauss-Legendre points per direction;
 *   b_i = sum_q w_q |J_q| f(x_q) phi_i(x_q);
 *   ComputeL2Error: sqrt( sum_e sum_q w_q |J_q| (u_h(x_q) - u(x_q))^2 ).
 * Deliberately written as dense loops over (point, basis function) -- no sum
 * factorisation -- so that it is an independent formulation of what the CUDA
 * kernels compute.  PARITY UNPINNED (see cdm_oracle.h).
 */
#include "cdm_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static int ipow_(int b, int e) { int r = 1; while (e-- > 0) { r *= b; } return r; }

/* 1-D linear shape pair and its derivative */
static void lin1d(double t, double *n, double *dn) { n[0] = 1.0 - t; n[1] = t; dn[0] = -1.0; dn[1] = 1.0; }

/* vertex k of the reference square/cube in MFEM order -> (i,j,k) corner bits */
static const int corner2[4][2] = { {0, 0}, {1, 0}, {1, 1}, {0, 1} };
static const int corner3[8][3] = { {0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1} };

/* physical point and |det J| of the (bi/tri)linear map of element vertices X at xi */
static double map_point(int dim, const double *X, const double *xi, double *xphys)
{
   double n[3][2], dn[3][2], J[9];
   for (int c = 0; c < dim; c++) { lin1d(xi[c], n[c], dn[c]); }
   const int nv = (dim == 2) ? 4 : 8;
   memset(J, 0, sizeof(J));
   for (int c = 0; c < dim; c++) { xphys[c] = 0.0; }
   for (int k = 0; k < nv; k++)
   {
      const int *cb = (dim == 2) ? corner2[k] : corner3[k];
      double N = 1.0;
      for (int c = 0; c < dim; c++) { N *= n[c][cb[c]]; }
      for (int a = 0; a < dim; a++) { xphys[a] += N * X[dim * k + a]; }
      for (int b = 0; b < dim; b++)
      {
         double dN = 1.0;
         for (int c = 0; c < dim; c++) { dN *= (c == b) ? dn[c][cb[c]] : n[c][cb[c]]; }
         for (int a = 0; a < dim; a++) { J[dim * a + b] += dN * X[dim * k + a]; }
      }
   }
   if (dim == 2) { return J[0] * J[3] - J[1] * J[2]; }
   return J[0] * (J[4] * J[8] - J[5] * J[7]) - J[1] * (J[3] * J[8] - J[5] * J[6]) + J[2] * (J[3] * J[7] - J[4] * J[6]);
}

static void elem_X(int dim, const int32_t *ev, const double *vx, int64_t e, double *X)
{
   const int nv = (dim == 2) ? 4 : 8;
   for (int k = 0; k < nv; k++)
      for (int c = 0; c < dim; c++) { X[dim * k + c] = vx[(int64_t)ev[nv * e + k] * dim + c]; }
}

/* out[(e*nq + q)*dim + c]: physical coordinates of the q1d^dim tensor rule, x fastest */
void orc_rule_coords(int dim, int q1d, int64_t ne, const int32_t *ev, const double *vx, double *out)
{
   const int nq = ipow_(q1d, dim);
   double *xq = malloc(sizeof(double) * q1d), *wq = malloc(sizeof(double) * q1d);
   orc_gauss_legendre(q1d, xq, wq);
   #pragma omp parallel for schedule(static)
   for (int64_t e = 0; e < ne; e++)
   {
      double X[24];
      elem_X(dim, ev, vx, e, X);
      for (int q = 0; q < nq; q++)
      {
         double xi[3] = { xq[q % q1d], xq[(q / q1d) % q1d], (dim == 3) ? xq[q / (q1d * q1d)] : 0.0 };
         map_point(dim, X, xi, out + (e * nq + q) * dim);
      }
   }
   free(xq); free(wq);
}

/* b[elem_dof[e][i]] += scale * sum_q w_q |J| f_q phi_i(x_q); b must be initialised by the caller */
void orc_domain_lf(int dim, int p, int q1d, int64_t ne, const int32_t *ev, const double *vx,
                   const int32_t *elem_dof, const double *f_q, double scale, double *b)
{
   const int d1d = p + 1, nd = ipow_(d1d, dim), nq = ipow_(q1d, dim);
   double *B = malloc(sizeof(double) * q1d * d1d), *G = malloc(sizeof(double) * q1d * d1d);
   double *qw = malloc(sizeof(double) * q1d), *xq = malloc(sizeof(double) * q1d);
   orc_basis(p, q1d, B, G, qw);
   orc_gauss_legendre(q1d, xq, qw);
   for (int64_t e = 0; e < ne; e++)       /* serial: fixed summation order */
   {
      double X[24], xp[3];
      elem_X(dim, ev, vx, e, X);
      for (int i = 0; i < nd; i++)
      {
         const int ix = i % d1d, iy = (i / d1d) % d1d, iz = (dim == 3) ? i / (d1d * d1d) : 0;
         double s = 0.0;
         for (int q = 0; q < nq; q++)
         {
            const int qx = q % q1d, qy = (q / q1d) % q1d, qz = (dim == 3) ? q / (q1d * q1d) : 0;
            double xi[3] = { xq[qx], xq[qy], (dim == 3) ? xq[qz] : 0.0 };
            const double det = map_point(dim, X, xi, xp);
            const double w = qw[qx] * qw[qy] * ((dim == 3) ? qw[qz] : 1.0);
            const double phi = B[qx * d1d + ix] * B[qy * d1d + iy] * ((dim == 3) ? B[qz * d1d + iz] : 1.0);
            s += w * det * f_q[e * nq + q] * phi;
         }
         b[elem_dof[e * nd + i]] += scale * s;
      }
   }
   free(B); free(G); free(qw); free(xq);
}

/* sqrt( sum_e sum_q w |J| (u_h - uex)^2 ); u == NULL -> u_h = 0, uex_q == NULL -> uex = 0 */
double orc_l2_error(int dim, int p, int q1d, int64_t ne, const int32_t *ev, const double *vx,
                    const int32_t *elem_dof, const double *u, const double *uex_q)
{
   const int d1d = p + 1, nd = ipow_(d1d, dim), nq = ipow_(q1d, dim);
   double *B = malloc(sizeof(double) * q1d * d1d), *G = malloc(sizeof(double) * q1d * d1d);
   double *qw = malloc(sizeof(double) * q1d), *xq = malloc(sizeof(double) * q1d);
   orc_basis(p, q1d, B, G, qw);
   orc_gauss_legendre(q1d, xq, qw);
   double total = 0.0;
   for (int64_t e = 0; e < ne; e++)
   {
      double X[24], xp[3];
      elem_X(dim, ev, vx, e, X);
      double se = 0.0;
      for (int q = 0; q < nq; q++)
      {
         const int qx = q % q1d, qy = (q / q1d) % q1d, qz = (dim == 3) ? q / (q1d * q1d) : 0;
         double xi[3] = { xq[qx], xq[qy], (dim == 3) ? xq[qz] : 0.0 };
         const double det = map_point(dim, X, xi, xp);
         const double w = qw[qx] * qw[qy] * ((dim == 3) ? qw[qz] : 1.0);
         double uh = 0.0;
         if (u)
            for (int i = 0; i < nd; i++)
            {
               const int ix = i % d1d, iy = (i / d1d) % d1d, iz = (dim == 3) ? i / (d1d * d1d) : 0;
               uh += u[elem_dof[e * nd + i]] * B[qx * d1d + ix] * B[qy * d1d + iy] * ((dim == 3) ? B[qz * d1d + iz] : 1.0);
            }
         const double d = uh - (uex_q ? uex_q[e * nq + q] : 0.0);
         se += w * det * d * d;
      }
      total += se;
   }
   free(B); free(G); free(qw); free(xq);
   return sqrt(total);
}
