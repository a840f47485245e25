/* oracle_space.c -- CPU ORACLE (test infrastructure): rules, basis, Cartesian
 * mesh, H1 numbering and the ElementRestriction index arrays.
 *
 * Restates (upstream MFEM, not vendored under /root/reference -- see
 * SURVEY.md Appendix C.1-C.3) the objects built by the reference at
 *   linear_convection_diffusion_2D.cpp:290-305  (Mesh / ParMesh)
 *   linear_convection_diffusion_2D.cpp:311-312  (H1_FECollection, ParFiniteElementSpace)
 *   linear_convection_diffusion_2D.cpp:319-322  (GetEssentialTrueDofs)
 * PARITY UNPINNED (see cdm_oracle.h).
 */
#include "cdm_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

/* ------------------------------------------------------------------ rules */

/* Legendre P_n and derivative at z in [-1,1] by the three-term recurrence */
static void legendre(int n, double z, double *pn, double *dpn)
{
   double p0 = 1.0, p1 = z;
   if (n == 0) { *pn = 1.0; *dpn = 0.0; return; }
   for (int k = 2; k <= n; k++)
   {
      double p2 = ((2.0 * k - 1.0) * z * p1 - (k - 1.0) * p0) / k;
      p0 = p1; p1 = p2;
   }
   *pn = p1;
   *dpn = n * (z * p1 - p0) / (z * z - 1.0);
}

/* MFEM IntRules.Get(SEGMENT, order): n = order/2+1 Gauss-Legendre points on
   [0,1], ascending, weights summing to 1 (fem/intrules.cpp GaussLegendre). */
void orc_gauss_legendre(int n, double *x, double *w)
{
   for (int i = 0; i < (n + 1) / 2; i++)
   {
      double z = cos(M_PI * (i + 0.75) / (n + 0.5));
      double pn, dpn;
      for (int it = 0; it < 100; it++)
      {
         legendre(n, z, &pn, &dpn);
         double dz = pn / dpn;
         z -= dz;
         if (fabs(dz) < 1e-16) { break; }
      }
      legendre(n, z, &pn, &dpn);
      double wt = 2.0 / ((1.0 - z * z) * dpn * dpn);
      /* z is the i-th largest root; map [-1,1] -> [0,1] */
      x[n - 1 - i] = 0.5 * (1.0 + z);
      x[i] = 0.5 * (1.0 - z);
      w[n - 1 - i] = w[i] = 0.5 * wt;
   }
}

/* Gauss-Lobatto nodes (H1_FECollection default BasisType::GaussLobatto):
   endpoints plus the roots of P'_{n-1}, on [0,1], ascending. */
void orc_gauss_lobatto(int n, double *x)
{
   x[0] = 0.0; x[n - 1] = 1.0;
   if (n == 1) { x[0] = 0.5; return; }
   int m = n - 1; /* roots of P'_m */
   for (int i = 1; i <= (n - 2 + 1) / 2; i++)
   {
      /* Chebyshev-Gauss-Lobatto initial guess, then Newton on q(z)=P'_m(z) */
      double z = cos(M_PI * i / m);
      for (int it = 0; it < 100; it++)
      {
         double pm, dpm;
         legendre(m, z, &pm, &dpm);
         /* (1-z^2) P''_m = 2 z P'_m - m(m+1) P_m */
         double d2 = (2.0 * z * dpm - m * (m + 1.0) * pm) / (1.0 - z * z);
         double dz = dpm / d2;
         z -= dz;
         if (fabs(dz) < 1e-16) { break; }
      }
      x[n - 1 - i] = 0.5 * (1.0 + z);
      x[i] = 0.5 * (1.0 - z);
   }
   if (n % 2 == 1) { x[n / 2] = 0.5; }
}

int orc_q1d(int dim, int p)
{
   /* Diffusion 2p+dim-1, Mass 2p+dim-1, Convection 2p+dim-1 on Q1 meshes:
      IntRules.Get(order) -> n = order/2+1 */
   int order = 2 * p + dim - 1;
   return order / 2 + 1;
}

/* Lagrange basis on the p+1 GLL nodes evaluated at the q1d Gauss points. */
void orc_basis(int p, int q1d, double *B, double *G, double *qw)
{
   int d1d = p + 1;
   double *xn = malloc(sizeof(double) * d1d);
   double *xq = malloc(sizeof(double) * q1d);
   orc_gauss_lobatto(d1d, xn);
   orc_gauss_legendre(q1d, xq, qw);
   for (int q = 0; q < q1d; q++)
   {
      double x = xq[q];
      for (int j = 0; j < d1d; j++)
      {
         double val = 1.0;
         for (int k = 0; k < d1d; k++)
            if (k != j) { val *= (x - xn[k]) / (xn[j] - xn[k]); }
         double der = 0.0;
         for (int m = 0; m < d1d; m++)
         {
            if (m == j) { continue; }
            double t = 1.0 / (xn[j] - xn[m]);
            for (int k = 0; k < d1d; k++)
               if (k != j && k != m) { t *= (x - xn[k]) / (xn[j] - xn[k]); }
            der += t;
         }
         B[q * d1d + j] = val;
         G[q * d1d + j] = der;
      }
   }
   free(xn); free(xq);
}

/* ------------------------------------------------------------------- mesh */

void orc_cart_sizes(int dim, const int64_t *n, int64_t *nv, int64_t *ne, int64_t *nbe)
{
   if (dim == 2)
   {
      *nv = (n[0] + 1) * (n[1] + 1);
      *ne = n[0] * n[1];
      *nbe = 2 * (n[0] + n[1]);
   }
   else
   {
      *nv = (n[0] + 1) * (n[1] + 1) * (n[2] + 1);
      *ne = n[0] * n[1] * n[2];
      *nbe = 2 * (n[0] * n[1] + n[1] * n[2] + n[0] * n[2]);
   }
}

/* Smooth deterministic displacement that vanishes on the boundary of the box
   (SURVEY.md 8(d) "smoothly perturbed interior vertices"). a is in units of h. */
static void perturb_point(int dim, const double *s, const double *h, double a, double *X)
{
   const double tp = 2.0 * M_PI;
   double u = X[0] / s[0], v = X[1] / s[1], w = (dim == 3) ? X[2] / s[2] : 0.0;
   double bump = sin(M_PI * u) * sin(M_PI * v) * ((dim == 3) ? sin(M_PI * w) : 1.0);
   double d0 = sin(tp * (u + v) + 0.3);
   double d1 = cos(tp * (v + w) + 0.5);
   double d2 = sin(tp * (w + u) + 1.0);
   X[0] += a * h[0] * bump * d0;
   X[1] += a * h[1] * bump * d1;
   if (dim == 3) { X[2] += a * h[2] * bump * d2; }
}

void orc_cart_mesh(int dim, const int64_t *n, const double *s, double perturb,
                   double *vx, int32_t *ev, int32_t *bv, int32_t *battr)
{
   const int64_t nx = n[0], ny = n[1], nz = (dim == 3) ? n[2] : 0;
   double h[3] = { s[0] / nx, s[1] / ny, (dim == 3) ? s[2] / nz : 0.0 };
   if (dim == 2)
   {
      for (int64_t j = 0; j <= ny; j++)
         for (int64_t i = 0; i <= nx; i++)
         {
            double X[2] = { i * h[0], j * h[1] };
            if (i == nx) { X[0] = s[0]; }
            if (j == ny) { X[1] = s[1]; }
            if (perturb != 0.0 && i > 0 && i < nx && j > 0 && j < ny)
               perturb_point(2, s, h, perturb, X);
            int64_t v = i + j * (nx + 1);
            vx[2 * v] = X[0]; vx[2 * v + 1] = X[1];
         }
#define V2(i,j) ((int32_t)((i) + (j) * (nx + 1)))
      int64_t e = 0;
      for (int64_t j = 0; j < ny; j++)
         for (int64_t i = 0; i < nx; i++, e++)
         {
            ev[4 * e + 0] = V2(i, j);     ev[4 * e + 1] = V2(i + 1, j);
            ev[4 * e + 2] = V2(i + 1, j + 1); ev[4 * e + 3] = V2(i, j + 1);
         }
      /* boundary: bottom 1, right 2, top 3, left 4 */
      int64_t b = 0;
      for (int64_t i = 0; i < nx; i++, b++) { bv[2*b] = V2(i, 0); bv[2*b+1] = V2(i + 1, 0); battr[b] = 1; }
      for (int64_t j = 0; j < ny; j++, b++) { bv[2*b] = V2(nx, j); bv[2*b+1] = V2(nx, j + 1); battr[b] = 2; }
      for (int64_t i = 0; i < nx; i++, b++) { bv[2*b] = V2(i + 1, ny); bv[2*b+1] = V2(i, ny); battr[b] = 3; }
      for (int64_t j = 0; j < ny; j++, b++) { bv[2*b] = V2(0, j + 1); bv[2*b+1] = V2(0, j); battr[b] = 4; }
#undef V2
      return;
   }
   for (int64_t k = 0; k <= nz; k++)
      for (int64_t j = 0; j <= ny; j++)
         for (int64_t i = 0; i <= nx; i++)
         {
            double X[3] = { i * h[0], j * h[1], k * h[2] };
            if (i == nx) { X[0] = s[0]; }
            if (j == ny) { X[1] = s[1]; }
            if (k == nz) { X[2] = s[2]; }
            if (perturb != 0.0 && i > 0 && i < nx && j > 0 && j < ny && k > 0 && k < nz)
               perturb_point(3, s, h, perturb, X);
            int64_t v = i + (nx + 1) * (j + (ny + 1) * k);
            vx[3 * v] = X[0]; vx[3 * v + 1] = X[1]; vx[3 * v + 2] = X[2];
         }
#define V3(i,j,k) ((int32_t)((i) + (nx + 1) * ((j) + (ny + 1) * (k))))
   int64_t e = 0;
   for (int64_t k = 0; k < nz; k++)
      for (int64_t j = 0; j < ny; j++)
         for (int64_t i = 0; i < nx; i++, e++)
         {
            int32_t *v = ev + 8 * e;
            v[0] = V3(i, j, k);         v[1] = V3(i + 1, j, k);
            v[2] = V3(i + 1, j + 1, k); v[3] = V3(i, j + 1, k);
            v[4] = V3(i, j, k + 1);         v[5] = V3(i + 1, j, k + 1);
            v[6] = V3(i + 1, j + 1, k + 1); v[7] = V3(i, j + 1, k + 1);
         }
   /* boundary attrs: bottom(z=0)=1 front(y=0)=2 right(x=max)=3 back(y=max)=4
      left(x=0)=5 top(z=max)=6 ; vertex order = outward-facing local hex face */
   int64_t b = 0;
   for (int64_t j = 0; j < ny; j++) for (int64_t i = 0; i < nx; i++, b++)
   { int32_t *q = bv + 4*b; q[0]=V3(i,j+1,0); q[1]=V3(i+1,j+1,0); q[2]=V3(i+1,j,0); q[3]=V3(i,j,0); battr[b]=1; }
   for (int64_t j = 0; j < ny; j++) for (int64_t i = 0; i < nx; i++, b++)
   { int32_t *q = bv + 4*b; q[0]=V3(i,j,nz); q[1]=V3(i+1,j,nz); q[2]=V3(i+1,j+1,nz); q[3]=V3(i,j+1,nz); battr[b]=6; }
   for (int64_t k = 0; k < nz; k++) for (int64_t j = 0; j < ny; j++, b++)
   { int32_t *q = bv + 4*b; q[0]=V3(0,j+1,k); q[1]=V3(0,j,k); q[2]=V3(0,j,k+1); q[3]=V3(0,j+1,k+1); battr[b]=5; }
   for (int64_t k = 0; k < nz; k++) for (int64_t j = 0; j < ny; j++, b++)
   { int32_t *q = bv + 4*b; q[0]=V3(nx,j,k); q[1]=V3(nx,j+1,k); q[2]=V3(nx,j+1,k+1); q[3]=V3(nx,j,k+1); battr[b]=3; }
   for (int64_t k = 0; k < nz; k++) for (int64_t i = 0; i < nx; i++, b++)
   { int32_t *q = bv + 4*b; q[0]=V3(i,0,k); q[1]=V3(i+1,0,k); q[2]=V3(i+1,0,k+1); q[3]=V3(i,0,k+1); battr[b]=2; }
   for (int64_t k = 0; k < nz; k++) for (int64_t i = 0; i < nx; i++, b++)
   { int32_t *q = bv + 4*b; q[0]=V3(i+1,ny,k); q[1]=V3(i,ny,k); q[2]=V3(i,ny,k+1); q[3]=V3(i+1,ny,k+1); battr[b]=4; }
#undef V3
}

/* ------------------------------------------------------------ entity tables */

/* MFEM mesh/hexahedron.cpp, mesh/quadrilateral.cpp local topology */
static const int HEX_EDGES[12][2] = { {0,1},{1,2},{3,2},{0,3},{4,5},{5,6},{7,6},{4,7},{0,4},{1,5},{2,6},{3,7} };
static const int HEX_FACES[6][4] = { {3,2,1,0},{0,1,5,4},{1,2,6,5},{2,3,7,6},{3,0,4,7},{4,5,6,7} };
static const int QUAD_EDGES[4][2] = { {0,1},{1,2},{2,3},{3,0} };

/* DSTable-like: edges keyed by (lo,hi), ids handed out in first-encounter order */
typedef struct { int64_t nv; int32_t *head; int32_t *next, *hi; int64_t cnt, cap; } edge_tab;

static void et_init(edge_tab *t, int64_t nv)
{
   t->nv = nv; t->head = malloc(sizeof(int32_t) * nv);
   for (int64_t i = 0; i < nv; i++) { t->head[i] = -1; }
   t->cap = 4 * nv + 16; t->cnt = 0;
   t->next = malloc(sizeof(int32_t) * t->cap); t->hi = malloc(sizeof(int32_t) * t->cap);
}
static void et_free(edge_tab *t) { free(t->head); free(t->next); free(t->hi); }
static int32_t et_get(edge_tab *t, int32_t a, int32_t b, int insert)
{
   int32_t lo = a < b ? a : b, hi = a < b ? b : a;
   for (int32_t k = t->head[lo]; k >= 0; k = t->next[k])
      if (t->hi[k] == hi) { return k; }
   if (!insert) { return -1; }
   if (t->cnt == t->cap)
   {
      t->cap *= 2;
      t->next = realloc(t->next, sizeof(int32_t) * t->cap);
      t->hi = realloc(t->hi, sizeof(int32_t) * t->cap);
   }
   int32_t id = (int32_t)t->cnt++;
   t->hi[id] = hi; t->next[id] = t->head[lo]; t->head[lo] = id;
   return id;
}

/* STable3D-like: quads keyed by their sorted vertex 4-tuple; stores the vertex
   order of the first element that created the face */
typedef struct { int64_t cap, cnt; int32_t *slot; int32_t *key; int32_t *base; int64_t fcap; } face_tab;

static void sort4(int32_t *k)
{
   for (int i = 1; i < 4; i++)
   {
      int32_t v = k[i]; int j = i - 1;
      while (j >= 0 && k[j] > v) { k[j + 1] = k[j]; j--; }
      k[j + 1] = v;
   }
}
static void ft_init(face_tab *t, int64_t ne)
{
   t->cap = 1; while (t->cap < 16 * ne + 64) { t->cap <<= 1; }
   t->slot = malloc(sizeof(int32_t) * t->cap);
   for (int64_t i = 0; i < t->cap; i++) { t->slot[i] = -1; }
   t->fcap = 4 * ne + 16; t->cnt = 0;
   t->key = malloc(sizeof(int32_t) * 4 * t->fcap);
   t->base = malloc(sizeof(int32_t) * 4 * t->fcap);
}
static void ft_free(face_tab *t) { free(t->slot); free(t->key); free(t->base); }
static int32_t ft_get(face_tab *t, const int32_t *v, int insert)
{
   int32_t k[4] = { v[0], v[1], v[2], v[3] };
   sort4(k);
   uint64_t h = 1469598103934665603ULL;
   for (int i = 0; i < 4; i++) { h ^= (uint64_t)(uint32_t)k[i]; h *= 1099511628211ULL; }
   uint64_t m = (uint64_t)t->cap - 1;
   for (uint64_t s = h & m;; s = (s + 1) & m)
   {
      int32_t id = t->slot[s];
      if (id < 0)
      {
         if (!insert) { return -1; }
         if (t->cnt == t->fcap)
         {
            t->fcap *= 2;
            t->key = realloc(t->key, sizeof(int32_t) * 4 * t->fcap);
            t->base = realloc(t->base, sizeof(int32_t) * 4 * t->fcap);
         }
         id = (int32_t)t->cnt++;
         memcpy(t->key + 4 * id, k, sizeof(k));
         memcpy(t->base + 4 * id, v, sizeof(int32_t) * 4);
         t->slot[s] = id;
         return id;
      }
      if (!memcmp(t->key + 4 * id, k, sizeof(k))) { return id; }
   }
}

/* Mesh::GetQuadOrientation(base, test) [MFEM-upstream, from memory] */
static int quad_orientation(const int32_t *base, const int32_t *test)
{
   int i;
   for (i = 0; i < 4; i++) if (test[i] == base[0]) { break; }
   if (test[(i + 1) % 4] == base[1]) { return 2 * i; }
   return 2 * i + 1;
}

/* H1_FECollection QuadDofOrd[ori][o], o = i + j*(p-1)  [MFEM-upstream] */
static int quad_dof_ord(int ori, int pm1, int i, int j)
{
   int pm2 = pm1 - 1;
   switch (ori)
   {
      case 0: return i + j * pm1;
      case 1: return j + i * pm1;
      case 2: return j + (pm2 - i) * pm1;
      case 3: return (pm2 - i) + j * pm1;
      case 4: return (pm2 - i) + (pm2 - j) * pm1;
      case 5: return (pm2 - j) + (pm2 - i) * pm1;
      case 6: return (pm2 - j) + i * pm1;
      default: return i + (pm2 - j) * pm1;
   }
}

/* lexicographic -> native dof map of H1_QuadrilateralElement / H1_HexahedronElement */
static void h1_dof_map(int dim, int p, int *dof_map)
{
   int p1 = p + 1, o = 0;
   if (dim == 2)
   {
      dof_map[0 + 0 * p1] = o++; dof_map[p + 0 * p1] = o++;
      dof_map[p + p * p1] = o++; dof_map[0 + p * p1] = o++;
      for (int i = 1; i < p; i++) { dof_map[i + 0 * p1] = o++; }
      for (int i = 1; i < p; i++) { dof_map[p + i * p1] = o++; }
      for (int i = 1; i < p; i++) { dof_map[(p - i) + p * p1] = o++; }
      for (int i = 1; i < p; i++) { dof_map[0 + (p - i) * p1] = o++; }
      for (int j = 1; j < p; j++) for (int i = 1; i < p; i++) { dof_map[i + j * p1] = o++; }
      return;
   }
#define L(i,j,k) ((i) + p1 * ((j) + p1 * (k)))
   dof_map[L(0,0,0)] = o++; dof_map[L(p,0,0)] = o++; dof_map[L(p,p,0)] = o++; dof_map[L(0,p,0)] = o++;
   dof_map[L(0,0,p)] = o++; dof_map[L(p,0,p)] = o++; dof_map[L(p,p,p)] = o++; dof_map[L(0,p,p)] = o++;
   for (int i = 1; i < p; i++) { dof_map[L(i,0,0)] = o++; }   /* (0,1) */
   for (int i = 1; i < p; i++) { dof_map[L(p,i,0)] = o++; }   /* (1,2) */
   for (int i = 1; i < p; i++) { dof_map[L(i,p,0)] = o++; }   /* (3,2) */
   for (int i = 1; i < p; i++) { dof_map[L(0,i,0)] = o++; }   /* (0,3) */
   for (int i = 1; i < p; i++) { dof_map[L(i,0,p)] = o++; }   /* (4,5) */
   for (int i = 1; i < p; i++) { dof_map[L(p,i,p)] = o++; }   /* (5,6) */
   for (int i = 1; i < p; i++) { dof_map[L(i,p,p)] = o++; }   /* (7,6) */
   for (int i = 1; i < p; i++) { dof_map[L(0,i,p)] = o++; }   /* (4,7) */
   for (int i = 1; i < p; i++) { dof_map[L(0,0,i)] = o++; }   /* (0,4) */
   for (int i = 1; i < p; i++) { dof_map[L(p,0,i)] = o++; }   /* (1,5) */
   for (int i = 1; i < p; i++) { dof_map[L(p,p,i)] = o++; }   /* (2,6) */
   for (int i = 1; i < p; i++) { dof_map[L(0,p,i)] = o++; }   /* (3,7) */
   for (int j = 1; j < p; j++) for (int i = 1; i < p; i++) { dof_map[L(i,p-j,0)] = o++; }   /* (3,2,1,0) */
   for (int j = 1; j < p; j++) for (int i = 1; i < p; i++) { dof_map[L(i,0,j)] = o++; }     /* (0,1,5,4) */
   for (int j = 1; j < p; j++) for (int i = 1; i < p; i++) { dof_map[L(p,i,j)] = o++; }     /* (1,2,6,5) */
   for (int j = 1; j < p; j++) for (int i = 1; i < p; i++) { dof_map[L(p-i,p,j)] = o++; }   /* (2,3,7,6) */
   for (int j = 1; j < p; j++) for (int i = 1; i < p; i++) { dof_map[L(0,p-i,j)] = o++; }   /* (3,0,4,7) */
   for (int j = 1; j < p; j++) for (int i = 1; i < p; i++) { dof_map[L(i,j,p)] = o++; }     /* (4,5,6,7) */
   for (int k = 1; k < p; k++) for (int j = 1; j < p; j++) for (int i = 1; i < p; i++) { dof_map[L(i,j,k)] = o++; }
#undef L
}

static int ipow(int b, int e) { int r = 1; while (e-- > 0) { r *= b; } return r; }

/* Build entity tables in first-encounter order (element order, local entity order). */
static void build_tables(int dim, int64_t nv, int64_t ne, const int32_t *ev,
                         edge_tab *et, face_tab *ft)
{
   et_init(et, nv);
   if (dim == 3) { ft_init(ft, ne); }
   for (int64_t e = 0; e < ne; e++)
   {
      if (dim == 2)
      {
         const int32_t *v = ev + 4 * e;
         for (int k = 0; k < 4; k++) { et_get(et, v[QUAD_EDGES[k][0]], v[QUAD_EDGES[k][1]], 1); }
      }
      else
      {
         const int32_t *v = ev + 8 * e;
         for (int k = 0; k < 12; k++) { et_get(et, v[HEX_EDGES[k][0]], v[HEX_EDGES[k][1]], 1); }
         for (int k = 0; k < 6; k++)
         {
            int32_t fv[4];
            for (int i = 0; i < 4; i++) { fv[i] = v[HEX_FACES[k][i]]; }
            ft_get(ft, fv, 1);
         }
      }
   }
}

int64_t orc_h1_build(int dim, int p, int64_t nv, int64_t ne, const int32_t *ev,
                     int32_t *elem_dof, int64_t *nedges_out, int64_t *nfaces_out)
{
   const int p1 = p + 1, pm1 = p - 1, nd = ipow(p1, dim);
   edge_tab et; face_tab ft; memset(&ft, 0, sizeof(ft));
   build_tables(dim, nv, ne, ev, &et, &ft);
   const int64_t nedges = et.cnt, nfaces = (dim == 3) ? ft.cnt : 0;
   const int64_t edge0 = nv, face0 = edge0 + nedges * pm1;
   const int64_t int0 = face0 + nfaces * (int64_t)pm1 * pm1;
   const int nint = ipow(pm1, dim);
   const int64_t ndof = int0 + ne * nint;
   int *dof_map = malloc(sizeof(int) * nd);
   int32_t *native = malloc(sizeof(int32_t) * nd);
   h1_dof_map(dim, p, dof_map);
   const int nvpe = (dim == 2) ? 4 : 8, nepe = (dim == 2) ? 4 : 12;
   for (int64_t e = 0; e < ne; e++)
   {
      const int32_t *v = ev + nvpe * e;
      int o = 0;
      for (int k = 0; k < nvpe; k++) { native[o++] = v[k]; }
      for (int k = 0; k < nepe; k++)
      {
         int32_t a = (dim == 2) ? v[QUAD_EDGES[k][0]] : v[HEX_EDGES[k][0]];
         int32_t b = (dim == 2) ? v[QUAD_EDGES[k][1]] : v[HEX_EDGES[k][1]];
         int32_t id = et_get(&et, a, b, 0);
         for (int i = 0; i < pm1; i++)
         {
            int ii = (a < b) ? i : (pm1 - 1 - i);   /* global orientation: low -> high vertex id */
            native[o++] = (int32_t)(edge0 + (int64_t)id * pm1 + ii);
         }
      }
      if (dim == 3)
      {
         for (int k = 0; k < 6; k++)
         {
            int32_t fv[4];
            for (int i = 0; i < 4; i++) { fv[i] = v[HEX_FACES[k][i]]; }
            int32_t id = ft_get(&ft, fv, 0);
            int ori = quad_orientation(ft.base + 4 * id, fv);
            for (int j = 0; j < pm1; j++)
               for (int i = 0; i < pm1; i++)
                  native[o++] = (int32_t)(face0 + (int64_t)id * pm1 * pm1 + quad_dof_ord(ori, pm1, i, j));
         }
      }
      for (int i = 0; i < nint; i++) { native[o++] = (int32_t)(int0 + e * nint + i); }
      for (int l = 0; l < nd; l++) { elem_dof[e * nd + l] = native[dof_map[l]]; }
   }
   free(dof_map); free(native);
   et_free(&et); if (dim == 3) { ft_free(&ft); }
   if (nedges_out) { *nedges_out = nedges; }
   if (nfaces_out) { *nfaces_out = nfaces; }
   return ndof;
}

int orc_h1_bdr_dofs(int dim, int p, int64_t nv, int64_t ne, const int32_t *ev,
                    int64_t nbe, const int32_t *bv, const int32_t *battr,
                    const int32_t *marker, int nattr, uint8_t *dof_mark)
{
   const int pm1 = p - 1;
   edge_tab et; face_tab ft; memset(&ft, 0, sizeof(ft));
   build_tables(dim, nv, ne, ev, &et, &ft);
   const int64_t edge0 = nv, face0 = edge0 + et.cnt * pm1;
   int rc = 0;
   for (int64_t b = 0; b < nbe; b++)
   {
      int a = battr[b];
      if (a < 1 || a > nattr || !marker[a - 1]) { continue; }
      if (dim == 2)
      {
         const int32_t *v = bv + 2 * b;
         dof_mark[v[0]] = dof_mark[v[1]] = 1;
         int32_t id = et_get(&et, v[0], v[1], 0);
         if (id < 0) { rc = -1; continue; }
         for (int i = 0; i < pm1; i++) { dof_mark[edge0 + (int64_t)id * pm1 + i] = 1; }
      }
      else
      {
         const int32_t *v = bv + 4 * b;
         for (int i = 0; i < 4; i++)
         {
            dof_mark[v[i]] = 1;
            int32_t id = et_get(&et, v[i], v[(i + 1) % 4], 0);
            if (id < 0) { rc = -1; continue; }
            for (int k = 0; k < pm1; k++) { dof_mark[edge0 + (int64_t)id * pm1 + k] = 1; }
         }
         int32_t fid = ft_get(&ft, v, 0);
         if (fid < 0) { rc = -1; continue; }
         for (int k = 0; k < pm1 * pm1; k++) { dof_mark[face0 + (int64_t)fid * pm1 * pm1 + k] = 1; }
      }
   }
   et_free(&et); if (dim == 3) { ft_free(&ft); }
   return rc;
}

/* ElementRestriction (fem/restriction.cpp): offsets = counts prefix-summed,
   indices = the (e,d) pairs of each L-dof in increasing e*nd+d order. */
void orc_restriction(int64_t ne, int nd, int64_t ndof, const int32_t *gather,
                     int32_t *offsets, int32_t *indices)
{
   const int64_t n = ne * nd;
   for (int64_t i = 0; i <= ndof; i++) { offsets[i] = 0; }
   for (int64_t i = 0; i < n; i++) { offsets[gather[i] + 1]++; }
   for (int64_t i = 0; i < ndof; i++) { offsets[i + 1] += offsets[i]; }
   int32_t *fill = malloc(sizeof(int32_t) * (ndof ? ndof : 1));
   for (int64_t i = 0; i < ndof; i++) { fill[i] = offsets[i]; }
   for (int64_t i = 0; i < n; i++) { indices[fill[gather[i]]++] = (int32_t)i; }
   free(fill);
}

void orc_node_coords(int dim, int p, int64_t ne, const int32_t *ev, const double *vx,
                     double *out)
{
   const int p1 = p + 1, nd = ipow(p1, dim), nvpe = (dim == 2) ? 4 : 8;
   double *xn = malloc(sizeof(double) * p1);
   orc_gauss_lobatto(p1, xn);
   for (int64_t e = 0; e < ne; e++)
      for (int l = 0; l < nd; l++)
      {
         double xi = xn[l % p1], eta = xn[(l / p1) % p1], ze = (dim == 3) ? xn[l / (p1 * p1)] : 0.0;
         double N[8];
         if (dim == 2)
         {
            N[0] = (1 - xi) * (1 - eta); N[1] = xi * (1 - eta); N[2] = xi * eta; N[3] = (1 - xi) * eta;
         }
         else
         {
            N[0] = (1 - xi) * (1 - eta) * (1 - ze); N[1] = xi * (1 - eta) * (1 - ze);
            N[2] = xi * eta * (1 - ze);             N[3] = (1 - xi) * eta * (1 - ze);
            N[4] = (1 - xi) * (1 - eta) * ze;       N[5] = xi * (1 - eta) * ze;
            N[6] = xi * eta * ze;                   N[7] = (1 - xi) * eta * ze;
         }
         for (int c = 0; c < dim; c++)
         {
            double s = 0.0;
            for (int k = 0; k < nvpe; k++) { s += N[k] * vx[(int64_t)ev[nvpe * e + k] * dim + c]; }
            out[(e * nd + l) * dim + c] = s;
         }
      }
   free(xn);
}
