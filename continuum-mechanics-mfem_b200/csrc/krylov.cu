// krylov.cu -- device-resident GMRES(m) and CG drivers.  Vectors never leave the
// GPU; per iteration the host sees only the (j+2) Hessenberg scalars.
//
// Stands behind
//   PetscLinearSolver(A).Mult(B, X) with Input/petsc.opts:2-6
//       (linear_convection_diffusion_2D.cpp:368-374; per time step and block at
//        linear_convection_diffusion_1D.cpp:553-566)
//   mfem::CGSolver::Mult (mesh_recession_handler.cpp:270-276)
// Algorithm variants: SURVEY.md Appendix C.6 / C.7.
#include "cdm_internal.hpp"
#include "kernels_common.cuh"
#include <cmath>
#include <vector>

extern "C" int cdm_apply_tail(cdm_op *op, double *x_buf, double *y_buf, bool constrained);   // capi.cu

namespace
{
inline double *red_out(cdm_ctx *c) { return c->red_dev + (size_t)CDM_RED_MAXK * CDM_RED_BLOCKS; }

int ensure_ws(cdm_op *op, int64_t doubles)
{
   cdm_ctx *ctx = op->sp->ctx;
   if (op->kry_len >= doubles) { return CDM_OK; }
   if (op->kry_dev) { cudaFree(op->kry_dev); op->kry_dev = nullptr; op->kry_len = 0; }
   if (cudaMalloc(&op->kry_dev, sizeof(double) * (size_t)doubles) != cudaSuccess)
   { cudaGetLastError(); return cdm_fail(ctx, CDM_ENOMEM, "krylov: cannot allocate workspace"); }
   op->kry_len = doubles;
   return CDM_OK;
}

int ensure_dinv(cdm_op *op)
{
   cdm_ctx *ctx = op->sp->ctx;
   if (op->dinv_dev) { return CDM_OK; }
   CDM_CUDA(ctx, cudaMalloc(&op->dinv_dev, sizeof(double) * (size_t)op->sp->ndof));
   int rc = cdm_operator_diag(op, op->dinv_dev);
   if (!rc) { rc = cdm_k_recip(ctx, op->sp->ntrue, op->dinv_dev, op->dinv_dev); }     // a zero diagonal entry -> 1 (PCJacobi)
   if (rc) { cudaFree(op->dinv_dev); op->dinv_dev = nullptr; }                          // never keep a half-built diagonal
   return rc;
}

// fetch k doubles of the device result area to the host (one sync)
int fetch(cdm_ctx *c, const double *dev, int k, double *host)
{
   CDM_CUDA(c, cudaMemcpyAsync(c->red_host, dev, sizeof(double) * k, cudaMemcpyDeviceToHost, c->stream));
   CDM_CUDA(c, cudaStreamSynchronize(c->stream));
   for (int i = 0; i < k; i++) { host[i] = c->red_host[i]; }
   return CDM_OK;
}
}  // namespace

#define RC(call) do { int rc_ = (call); if (rc_) { return rc_; } } while (0)

extern "C" int cdm_gmres(cdm_op *op, const double *b, double *x, const cdm_krylov_opts *o,
                         cdm_krylov_result *res, double *hist)
{
   if (!op || !b || !x || !o || !res) { return CDM_EINVAL; }
   cdm_space *sp = op->sp;
   cdm_ctx *c = sp->ctx;
   CDM_REQUIRE_GPU(c);
   const int64_t n = sp->ntrue;
   const int64_t ld = (sp->ndof + 31) & ~(int64_t)31;     // room for the ghost tail
   const int m = o->restart > 0 ? o->restart : (o->variant == CDM_GMRES_PETSC ? 30 : 50);
   if (m + 1 > CDM_RED_MAXK) { return cdm_fail(c, CDM_EINVAL, "cdm_gmres: restart too large"); }
   RC(ensure_ws(op, (int64_t)(m + 3) * ld));
   double *V = op->kry_dev, *w = V + (int64_t)(m + 1) * ld, *t = w + ld;
   const double *dinv = nullptr;
   if (o->jacobi) { RC(ensure_dinv(op)); dinv = op->dinv_dev; }
   double *h_dev = red_out(c);                 // [m+1] dots, then ||w||^2 at h_dev[m+1]
   double *y_dev = red_out(c) + 2 * CDM_RED_MAXK;
   std::vector<double> H((size_t)(m + 1) * m, 0.0), cs(m), sn(m), s(m + 1), yv(m), hc(m + 2);
   int it = 0, conv = 0, hl = 0;
   double rnorm = 0.0, ttol = 0.0;
   bool first = true;
   CDM_CUDA(c, cudaEventRecord(c->ev0, c->stream));
   if (o->zero_guess) { RC(cdm_k_set(c, n, 0.0, x)); }
   while (true)
   {
      // V0 = M^{-1}(b - A x)
      if (first && o->zero_guess)
      {
         if (dinv) { RC(cdm_k_pmult(c, n, dinv, b, V)); }
         else { CDM_CUDA(c, cudaMemcpyAsync(V, b, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream)); }
      }
      else
      {
         CDM_CUDA(c, cudaMemcpyAsync(w, x, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
         RC(cdm_apply_tail(op, w, t, true));
         RC(cdm_k_add(c, n, b, -1.0, t, t));
         if (dinv) { RC(cdm_k_pmult(c, n, dinv, t, V)); }
         else { CDM_CUDA(c, cudaMemcpyAsync(V, t, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream)); }
      }
      RC(cdm_k_mdot_dev(c, n, 1, V, V, ld, h_dev));
      RC(cdm_allreduce_sum(c, h_dev, 1));
      double b2; RC(fetch(c, h_dev, 1, &b2));
      const double beta = std::sqrt(b2);
      rnorm = beta;
      if (first)
      {
         // KSPConvergedDefault: with a non-zero initial guess PETSc scales rtol by ||M^{-1} b||, the preconditioned
         // norm of the right-hand side; mfem::GMRESSolver always uses the initial residual
         double ref = beta;
         if (!o->zero_guess && o->variant == CDM_GMRES_PETSC)
         {
            if (dinv) { RC(cdm_k_pmult(c, n, dinv, b, w)); }
            RC(cdm_k_mdot_dev(c, n, 1, dinv ? w : b, dinv ? w : b, ld, h_dev));
            RC(cdm_allreduce_sum(c, h_dev, 1));
            double nb2; RC(fetch(c, h_dev, 1, &nb2));
            ref = std::sqrt(nb2);
         }
         ttol = std::fmax(o->rtol * ref, o->atol);
         if (hist) { hist[hl] = beta; } hl++;
         first = false;
      }
      if (rnorm <= ttol) { conv = 1; break; }
      if (it >= o->max_it) { break; }
      RC(cdm_k_scale(c, n, 1.0 / beta, V, V));
      std::fill(s.begin(), s.end(), 0.0);
      s[0] = beta;
      int j = 0;
      while (j < m && it < o->max_it)
      {
         double *vj = V + (int64_t)j * ld, *vn = V + (int64_t)(j + 1) * ld;
         RC(cdm_apply_tail(op, vj, dinv ? t : w, true));
         if (dinv && o->variant != CDM_GMRES_PETSC) { RC(cdm_k_pmult(c, n, dinv, t, w)); }
         if (o->variant == CDM_GMRES_PETSC)
         {
            // classical Gram-Schmidt: all j+1 dots against the same w, one all-reduce; the Jacobi
            // scaling w = dinv .* (A v_j) is fused into the first pass of the multi-dot
            if (dinv) { RC(cdm_k_mdot_pc_dev(c, n, j + 1, t, dinv, w, V, ld, h_dev)); }
            else { RC(cdm_k_mdot_dev(c, n, j + 1, w, V, ld, h_dev)); }
            RC(cdm_allreduce_sum(c, h_dev, j + 1));
            RC(cdm_k_maxpy_dev(c, n, j + 1, h_dev, V, ld, w, h_dev + (j + 1)));
         }
         else
         {
            // modified Gram-Schmidt, scalars stay on the device
            for (int i = 0; i <= j; i++)
            {
               RC(cdm_k_mdot_dev(c, n, 1, w, V + (int64_t)i * ld, ld, h_dev + i));
               RC(cdm_allreduce_sum(c, h_dev + i, 1));
               RC(cdm_k_maxpy_dev(c, n, 1, h_dev + i, V + (int64_t)i * ld, ld, w, i == j ? h_dev + (j + 1) : nullptr));
            }
         }
         RC(cdm_allreduce_sum(c, h_dev + (j + 1), 1));
         RC(cdm_k_scale_by_rnorm(c, n, h_dev + (j + 1), w, vn));
         RC(fetch(c, h_dev, j + 2, hc.data()));
         const double hn = std::sqrt(hc[j + 1]);
         hc[j + 1] = hn;
         for (int i = 0; i < j; i++)
         {
            const double a = cs[i] * hc[i] + sn[i] * hc[i + 1];
            hc[i + 1] = -sn[i] * hc[i] + cs[i] * hc[i + 1];
            hc[i] = a;
         }
         const double den = std::hypot(hc[j], hc[j + 1]);
         cs[j] = hc[j] / den; sn[j] = hc[j + 1] / den;
         hc[j] = den; hc[j + 1] = 0.0;
         s[j + 1] = -sn[j] * s[j];
         s[j] = cs[j] * s[j];
         for (int i = 0; i <= j; i++) { H[(size_t)j * (m + 1) + i] = hc[i]; }
         j++; it++;
         rnorm = std::fabs(s[j]);
         if (hist) { hist[hl] = rnorm; } hl++;
         if (rnorm <= ttol) { conv = 1; break; }
         if (hn == 0.0) { break; }
      }
      for (int i = j - 1; i >= 0; i--)
      {
         double a = s[i];
         for (int k = i + 1; k < j; k++) { a -= H[(size_t)k * (m + 1) + i] * yv[k]; }
         yv[i] = a / H[(size_t)i * (m + 1) + i];
      }
      // x += V y  (maxpy subtracts, so upload -y)
      for (int i = 0; i < j; i++) { c->red_host[2 * CDM_RED_MAXK + i] = -yv[i]; }
      CDM_CUDA(c, cudaMemcpyAsync(y_dev, c->red_host + 2 * CDM_RED_MAXK, sizeof(double) * j, cudaMemcpyHostToDevice, c->stream));
      RC(cdm_k_maxpy_dev(c, n, j, y_dev, V, ld, x, nullptr));
      CDM_CUDA(c, cudaStreamSynchronize(c->stream));   // red_host is reused
      if (conv || it >= o->max_it) { break; }
   }
   CDM_CUDA(c, cudaEventRecord(c->ev1, c->stream));
   CDM_CUDA(c, cudaEventSynchronize(c->ev1));
   float ms = 0.f;
   cudaEventElapsedTime(&ms, c->ev0, c->ev1);
   res->iters = it; res->converged = conv; res->final_norm = rnorm; res->hist_len = hl;
   res->seconds = ms * 1e-3;
   return CDM_OK;
}

extern "C" int cdm_cg(cdm_op *op, const double *b, double *x, const cdm_krylov_opts *o,
                      cdm_krylov_result *res, double *hist)
{
   if (!op || !b || !x || !o || !res) { return CDM_EINVAL; }
   cdm_space *sp = op->sp;
   cdm_ctx *c = sp->ctx;
   CDM_REQUIRE_GPU(c);
   const int64_t n = sp->ntrue;
   const int64_t ld = (sp->ndof + 31) & ~(int64_t)31;
   RC(ensure_ws(op, 4 * ld));
   double *r = op->kry_dev, *d = r + ld, *z = d + ld, *zz = z + ld;
   const double *dinv = nullptr;
   if (o->jacobi) { RC(ensure_dinv(op)); dinv = op->dinv_dev; }
   double *sc = red_out(c);
   int it = 0, conv = 0, hl = 0;
   CDM_CUDA(c, cudaEventRecord(c->ev0, c->stream));
   if (o->zero_guess)
   {
      RC(cdm_k_set(c, n, 0.0, x));
      CDM_CUDA(c, cudaMemcpyAsync(r, b, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
   }
   else
   {
      CDM_CUDA(c, cudaMemcpyAsync(d, x, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
      RC(cdm_apply_tail(op, d, r, true));
      RC(cdm_k_add(c, n, b, -1.0, r, r));
   }
   if (dinv) { RC(cdm_k_pmult(c, n, dinv, r, d)); }
   else { CDM_CUDA(c, cudaMemcpyAsync(d, r, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream)); }
   double nom, den, betanom;
   RC(cdm_k_mdot_dev(c, n, 1, d, r, ld, sc)); RC(cdm_allreduce_sum(c, sc, 1)); RC(fetch(c, sc, 1, &nom));
   const double r0 = std::fmax(nom * o->rtol * o->rtol, o->atol * o->atol);
   if (hist) { hist[hl] = nom; } hl++;
   betanom = nom;
   if (nom <= r0) { conv = 1; }
   else
   {
      RC(cdm_apply_tail(op, d, z, true));
      RC(cdm_k_mdot_dev(c, n, 1, z, d, ld, sc)); RC(cdm_allreduce_sum(c, sc, 1)); RC(fetch(c, sc, 1, &den));
      if (den > 0.0)
      {
         for (it = 1;; it++)
         {
            const double alpha = nom / den;
            RC(cdm_k_cg_update(c, n, alpha, d, z, x, r, sc));       // x += a d ; r -= a z ; sc = (r,r)
            if (dinv)
            {
               RC(cdm_k_pmult(c, n, dinv, r, zz));
               RC(cdm_k_mdot_dev(c, n, 1, r, zz, ld, sc));
            }
            RC(cdm_allreduce_sum(c, sc, 1));
            RC(fetch(c, sc, 1, &betanom));
            if (hist) { hist[hl] = betanom; } hl++;
            if (betanom <= r0) { conv = 1; break; }
            if (it >= o->max_it) { break; }
            const double beta = betanom / nom;
            RC(cdm_k_add(c, n, dinv ? zz : r, beta, d, d));          // d = z + beta d
            RC(cdm_apply_tail(op, d, z, true));
            RC(cdm_k_mdot_dev(c, n, 1, d, z, ld, sc)); RC(cdm_allreduce_sum(c, sc, 1)); RC(fetch(c, sc, 1, &den));
            if (den <= 0.0) { break; }
            nom = betanom;
         }
      }
   }
   CDM_CUDA(c, cudaEventRecord(c->ev1, c->stream));
   CDM_CUDA(c, cudaEventSynchronize(c->ev1));
   float ms = 0.f;
   cudaEventElapsedTime(&ms, c->ev0, c->ev1);
   res->iters = it; res->converged = conv; res->final_norm = std::sqrt(std::fabs(betanom)); res->hist_len = hl;
   res->seconds = ms * 1e-3;
   return CDM_OK;
}
