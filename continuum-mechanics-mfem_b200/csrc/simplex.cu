// simplex.cu -- device kernels for H1 Lagrange triangles ("needed for triangles / unstructured
// meshes where tensor PA does not apply"): quadrature data, linear form, error norm, rule points.  The operator itself
// is applied as the assembled CSR matrix (csr_path.cu), i.e. the reference's literal algorithm
// (a.Assemble() -> sparse matrix -> MatMult, linear_convection_diffusion_2D.cpp:339,368) -- triangles have no
// sum-factorised form.  Everything is dense-table based: B[q][i], dB/dxi[c][q][i] of the nodal basis at the points of a
// collapsed Gauss-Legendre rule (host_simplex.cpp); the element map is affine, so J is constant per element.
#include "cdm_internal.hpp"
#include "kernels_common.cuh"
#include <cmath>

namespace
{
struct SCoef { int kind; int ncomp; const double *data; double c[4]; };

// one thread per (element, point): D_diff = (w/det) A M A^T (11,21,22), D_conv = alpha w A c, D_mass = w s det,
// A = adj(J), J_ab = d x_a / d xi_b of the affine map; layout per element [component][nq]
__global__ void __launch_bounds__(256)
k_setup_qdata_tri(int64_t ne, int nq, const double *__restrict__ ex, const double *__restrict__ qw, SCoef kap, SCoef vel, double alpha,
                  SCoef mas, int has_diff, int has_conv, int has_mass, int slab, double *__restrict__ D)
{
   const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (t >= ne * nq) { return; }
   const int64_t e = t / nq;
   const int q = (int)(t - e * nq);
   const double *X = ex + e * 6;
   const double J11 = X[2] - X[0], J12 = X[4] - X[0], J21 = X[3] - X[1], J22 = X[5] - X[1];
   const double det = J11 * J22 - J12 * J21;
   const double A11 = J22, A12 = -J12, A21 = -J21, A22 = J11;       // adj(J)
   const double w = qw[q];
   double *dp = D + e * (int64_t)slab + q;
   int c = 0;
   if (has_diff)
   {
      double m11, m21, m22;
      const double *k = (kap.kind == CDM_COEFF_CONST) ? kap.c : kap.data + (e * nq + q) * kap.ncomp;
      if (kap.ncomp == 1) { m11 = m22 = k[0]; m21 = 0.0; } else { m11 = k[0]; m21 = k[1]; m22 = k[2]; }
      // A M A^T
      const double t11 = A11 * m11 + A12 * m21, t12 = A11 * m21 + A12 * m22;
      const double t21 = A21 * m11 + A22 * m21, t22 = A21 * m21 + A22 * m22;
      const double s = w / det;
      dp[0] = s * (t11 * A11 + t12 * A12);
      dp[nq] = s * (t21 * A11 + t22 * A12);
      dp[2 * nq] = s * (t21 * A21 + t22 * A22);
      c = 3;
   }
   if (has_conv)
   {
      const double *v = (vel.kind == CDM_COEFF_CONST) ? vel.c : vel.data + (e * nq + q) * 2;
      dp[c * nq] = alpha * w * (A11 * v[0] + A12 * v[1]);
      dp[(c + 1) * nq] = alpha * w * (A21 * v[0] + A22 * v[1]);
      c += 2;
   }
   if (has_mass)
   {
      const double s = (mas.kind == CDM_COEFF_CONST) ? mas.c[0] : mas.data[e * nq + q];
      dp[c * nq] = w * s * det;
   }
}

// physical points of a rule: x = x0 + J xi
__global__ void __launch_bounds__(256)
k_rule_coords_tri(int64_t ne, int nq, const double *__restrict__ ex, const double *__restrict__ qx, double *__restrict__ out)
{
   const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (t >= ne * nq) { return; }
   const int64_t e = t / nq;
   const int q = (int)(t - e * nq);
   const double *X = ex + e * 6;
   const double xi = qx[2 * q], eta = qx[2 * q + 1];
   out[2 * t] = X[0] + (X[2] - X[0]) * xi + (X[4] - X[0]) * eta;
   out[2 * t + 1] = X[1] + (X[3] - X[1]) * xi + (X[5] - X[1]) * eta;
}

// E-vector of the linear form: bE[e][i] = scale * sum_q B[q][i] w_q |J| f(e,q)
__global__ void __launch_bounds__(128)
k_domain_lf_tri(int64_t ne, int nq, int nd, const double *__restrict__ ex, const double *__restrict__ qw, const double *__restrict__ B,
                const double *__restrict__ fq, double scale, double *__restrict__ bE)
{
   const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (t >= ne * nd) { return; }
   const int64_t e = t / nd;
   const int i = (int)(t - e * nd);
   const double *X = ex + e * 6;
   const double det = fabs((X[2] - X[0]) * (X[5] - X[1]) - (X[4] - X[0]) * (X[3] - X[1]));
   double s = 0.0;
   for (int q = 0; q < nq; q++) { s += B[(size_t)q * nd + i] * qw[q] * fq[e * nq + q]; }
   bE[t] = scale * det * s;
}

// per element: sqrt( sum_q w |J| (u_h(x_q) - uex_q)^2 )
__global__ void __launch_bounds__(128)
k_l2_error_tri(int64_t ne, int nq, int nd, const double *__restrict__ ex, const double *__restrict__ qw, const double *__restrict__ B,
               const int32_t *__restrict__ gather, const double *__restrict__ u, const double *__restrict__ uex, double *__restrict__ out)
{
   const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (e >= ne) { return; }
   const double *X = ex + e * 6;
   const double det = fabs((X[2] - X[0]) * (X[5] - X[1]) - (X[4] - X[0]) * (X[3] - X[1]));
   double s = 0.0;
   for (int q = 0; q < nq; q++)
   {
      double uh = 0.0;
      if (u) { for (int i = 0; i < nd; i++) { uh += B[(size_t)q * nd + i] * u[gather[e * nd + i]]; } }
      const double d = uh - (uex ? uex[e * nq + q] : 0.0);
      s += qw[q] * d * d;
   }
   out[e] = sqrt(det * s);
}

bool on_device_ptr(const void *p)
{
   cudaPointerAttributes a;
   if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
   return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

SCoef make_coef(const cdm_coeff *c)
{
   SCoef s;
   s.kind = (c && c->kind != CDM_COEFF_NONE) ? c->kind : CDM_COEFF_NONE;
   s.ncomp = c ? c->ncomp : 0;
   s.data = nullptr;
   for (int i = 0; i < 4; i++) { s.c[i] = 0.0; }
   if (s.kind == CDM_COEFF_CONST) { for (int i = 0; i < s.ncomp && i < 4; i++) { s.c[i] = c->data[i]; } }
   return s;
}

// tables of a rule with n points per direction, uploaded to one temporary device block: [qw | qx | B]
struct RuleDev
{
   int nq = 0;
   double *block = nullptr, *qw = nullptr, *qx = nullptr, *B = nullptr;
   ~RuleDev() { cudaFree(block); }
};
int make_rule(const cdm_space *sp, int n, RuleDev &r)
{
   cdm_ctx *c = sp->ctx;
   const int nq = n * n, nd = sp->nd;
   std::vector<double> xy(2 * nq), w(nq), B((size_t)nq * nd), G((size_t)2 * nq * nd);
   cdm_host_tri_rule(n, xy.data(), w.data());
   cdm_host_tri_basis(sp->p, nq, xy.data(), B.data(), G.data());
   std::vector<double> h;
   h.insert(h.end(), w.begin(), w.end());
   h.insert(h.end(), xy.begin(), xy.end());
   h.insert(h.end(), B.begin(), B.end());
   CDM_CUDA(c, cudaMalloc(&r.block, h.size() * sizeof(double)));
   CDM_CUDA(c, cudaMemcpy(r.block, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
   r.nq = nq; r.qw = r.block; r.qx = r.block + nq; r.B = r.block + 3 * nq;
   return CDM_OK;
}
}  // namespace

int cdm_k_setup_qdata_simplex(cdm_op *op, const cdm_coeff *kappa, const cdm_coeff *vel, double alpha, const cdm_coeff *mass)
{
   cdm_space *sp = op->sp;
   cdm_ctx *c = sp->ctx;
   SCoef ck = make_coef(kappa), cv = make_coef(vel), cm = make_coef(mass);
   const cdm_coeff *src[3] = {kappa, vel, mass};
   SCoef *dst[3] = {&ck, &cv, &cm};
   for (int i = 0; i < 3; i++)
   {
      if (dst[i]->kind != CDM_COEFF_QPT) { continue; }
      const size_t bytes = sizeof(double) * (size_t)sp->ne * sp->nq * src[i]->ncomp;
      if (on_device_ptr(src[i]->data)) { dst[i]->data = src[i]->data; continue; }
      if (op->coef_bytes[i] < bytes)
      {
         cudaFree(op->coef_dev[i]); op->coef_dev[i] = nullptr; op->coef_bytes[i] = 0;
         CDM_CUDA(c, cudaMalloc(&op->coef_dev[i], bytes));
         op->coef_bytes[i] = bytes;
      }
      CDM_CUDA(c, cudaMemcpyAsync(op->coef_dev[i], src[i]->data, bytes, cudaMemcpyHostToDevice, c->stream));
      CDM_CUDA(c, cudaStreamSynchronize(c->stream));       // the caller's host array may be a temporary
      dst[i]->data = op->coef_dev[i];
   }
   const int64_t n = sp->ne * sp->nq;
   k_setup_qdata_tri<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(sp->ne, sp->nq, sp->elem_x_dev, sp->sqw_dev, ck, cv, alpha, cm,
                                                                        op->has_diff, op->has_conv, op->has_mass, op->slab, op->D_dev);
   c->launches++;
   CDM_CUDA(c, cudaGetLastError());
   return CDM_OK;
}

int cdm_simplex_rule_coords(const cdm_space *sp, int nq1d, double *xyz)
{
   cdm_ctx *c = sp->ctx;
   RuleDev r;
   int rc = make_rule(sp, nq1d, r); if (rc) { return rc; }
   const int64_t n = sp->ne * r.nq;
   const bool dev = on_device_ptr(xyz);
   double *out = xyz;
   if (!dev) { CDM_CUDA(c, cudaMalloc(&out, sizeof(double) * 2 * (size_t)n)); }
   k_rule_coords_tri<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(sp->ne, r.nq, sp->elem_x_dev, r.qx, out);
   c->launches++;
   cudaError_t e = cudaStreamSynchronize(c->stream);
   if (!dev)
   {
      if (e == cudaSuccess) { e = cudaMemcpy(xyz, out, sizeof(double) * 2 * (size_t)n, cudaMemcpyDeviceToHost); }
      cudaFree(out);
   }
   if (e != cudaSuccess) { return cdm_fail(c, CDM_ECUDA, cudaGetErrorString(e)); }
   return CDM_OK;
}

int cdm_simplex_domain_lf(cdm_space *sp, int nq1d, const double *f_q, double scale, int accumulate, double *b_dev)
{
   cdm_ctx *c = sp->ctx;
   RuleDev r;
   int rc = make_rule(sp, nq1d, r); if (rc) { return rc; }
   const int64_t nf = sp->ne * r.nq, nE = sp->ne * sp->nd;
   const double *fd = f_q;
   double *ftmp = nullptr, *bE = nullptr, *bL = nullptr;
   if (!on_device_ptr(f_q))
   {
      CDM_CUDA(c, cudaMalloc(&ftmp, sizeof(double) * (size_t)nf));
      CDM_CUDA(c, cudaMemcpy(ftmp, f_q, sizeof(double) * (size_t)nf, cudaMemcpyHostToDevice));
      fd = ftmp;
   }
   if (cudaMalloc(&bE, sizeof(double) * (size_t)nE) != cudaSuccess || cudaMalloc(&bL, sizeof(double) * (size_t)sp->ndof) != cudaSuccess)
   { cudaGetLastError(); cudaFree(ftmp); cudaFree(bE); return cdm_fail(c, CDM_ENOMEM, "cdm_domain_lf: out of device memory"); }
   k_domain_lf_tri<<<(unsigned)((nE + 127) / 128), 128, 0, c->stream>>>(sp->ne, r.nq, sp->nd, sp->elem_x_dev, r.qw, r.B, fd, scale, bE);
   c->launches++;
   rc = cdm_k_restrict_transpose(sp, bE, bL);                                 // deterministic gather (ElementRestriction^T)
   if (!rc) { rc = accumulate ? cdm_k_axpy(c, sp->ndof, 1.0, bL, b_dev)
                              : (cudaMemcpyAsync(b_dev, bL, sizeof(double) * (size_t)sp->ndof, cudaMemcpyDeviceToDevice, c->stream) == cudaSuccess ? CDM_OK : CDM_ECUDA); }
   cudaStreamSynchronize(c->stream);
   cudaFree(ftmp); cudaFree(bE); cudaFree(bL);
   return rc;
}

int cdm_simplex_l2_error(cdm_space *sp, int nq1d, const double *u_dev, const double *uex_q, double *result_host)
{
   cdm_ctx *c = sp->ctx;
   RuleDev r;
   int rc = make_rule(sp, nq1d, r); if (rc) { return rc; }
   const int64_t nf = sp->ne * r.nq;
   const double *ud = uex_q;
   double *utmp = nullptr, *per = nullptr;
   if (uex_q && !on_device_ptr(uex_q))
   {
      CDM_CUDA(c, cudaMalloc(&utmp, sizeof(double) * (size_t)nf));
      CDM_CUDA(c, cudaMemcpy(utmp, uex_q, sizeof(double) * (size_t)nf, cudaMemcpyHostToDevice));
      ud = utmp;
   }
   const int64_t ld = (sp->ne + 1) & ~(int64_t)1;
   if (cudaMalloc(&per, sizeof(double) * (size_t)ld) != cudaSuccess) { cudaGetLastError(); cudaFree(utmp); return cdm_fail(c, CDM_ENOMEM, "cdm_l2_error: out of device memory"); }
   k_l2_error_tri<<<(unsigned)((sp->ne + 127) / 128), 128, 0, c->stream>>>(sp->ne, r.nq, sp->nd, sp->elem_x_dev, r.qw, r.B, sp->gather_dev, u_dev, ud, per);
   c->launches++;
   double *out = c->red_dev + (size_t)CDM_RED_MAXK * CDM_RED_BLOCKS;
   rc = cdm_k_mdot_dev(c, sp->ne, 1, per, per, ld, out);                     // sum of the squared element norms, fixed order
   double s2 = 0.0;
   if (!rc && cudaMemcpyAsync(&s2, out, sizeof(double), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) { rc = CDM_ECUDA; }
   if (cudaStreamSynchronize(c->stream) != cudaSuccess && !rc) { rc = CDM_ECUDA; }
   cudaFree(utmp); cudaFree(per);
   if (rc) { return cdm_fail(c, rc, "cdm_l2_error (simplex) failed"); }
   *result_host = std::sqrt(s2);
   return CDM_OK;
}
