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

extern "C" int cdm_apply_tail(cdm_op *op, double *x_buf, double *y_buf, bool constrained, bool x_ghost_valid);   // capi.cu
extern "C" bool cdm_apply_keeps_ghosts(const cdm_op *op);

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

// fetch k doubles of the device result area to the host (one sync)
int fetch(cdm_ctx *c, const double *dev, int k, double *host)
{
   CDM_CUDA(c, cudaMemcpyAsync(c->red_host, dev, sizeof(double) * k, cudaMemcpyDeviceToHost, c->stream));
   CDM_CUDA(c, cudaStreamSynchronize(c->stream));
   for (int i = 0; i < k; i++) { host[i] = c->red_host[i]; }
   return cdm_check_p2p(c);          // a peer-memory kernel that timed out invalidates what was just read
}
}  // namespace

#define RC(call) do { int rc_ = (call); if (rc_) { return rc_; } } while (0)

// Jacobi inverse diagonal on the whole local vector: on partitioned spaces the ghost entries are filled from their
// owners so that dinv .* (ghost-consistent vector) stays ghost-consistent
static int ensure_dinv(cdm_op *op)
{
   cdm_space *sp = op->sp;
   cdm_ctx *ctx = sp->ctx;
   if (op->dinv_dev) { return CDM_OK; }
   CDM_CUDA(ctx, cudaMalloc(&op->dinv_dev, sizeof(double) * (size_t)sp->ndof));
   int rc = cdm_operator_diag(op, op->dinv_dev);
   if (!rc && ctx->nranks > 1 && !sp->peers.empty()) { rc = cdm_halo_P_space(sp, op->dinv_dev); }
   if (!rc) { rc = cdm_k_recip(ctx, sp->ndof, op->dinv_dev, op->dinv_dev); }            // a zero diagonal entry -> 1 (PCJacobi)
   if (rc) { cudaFree(op->dinv_dev); op->dinv_dev = nullptr; }                          // never keep a half-built diagonal
   return rc;
}

// Restarted GMRES with a LAZILY NORMALISED basis: V_i holds the un-normalised vector u_i, nrm2[i] = ||u_i||^2 stays on
// the device, v_i = u_i / ||u_i|| is never written.  One iteration is
//     t = A u_j                                          (element kernel + one shared-dof exchange)
//     u_{j+1} <- dinv .* t / ||u_j||,  h_i = (u_{j+1}, u_i) / ||u_i||      (fused multi-dot, Jacobi and scaling in its first pass)
//     u_{j+1} -= sum_i (h_i / ||u_i||) u_i,  nrm2[j+1] = ||u_{j+1}||^2      (fused multi-axpy + norm)
// i.e. classical Gram-Schmidt exactly as KSPGMRES does it, without the separate normalisation pass.  The host needs the
// j+2 scalars of iteration j only for the Givens rotations and the convergence test; they are copied to pinned
// memory behind an event, and the operator apply of iteration j+1 (which does not depend on them) is queued BEFORE the
// host waits for that event: the device never idles on the host.
// On partitioned spaces all vector updates run over the whole local vector (ghost tail included) and the dots over
// the true dofs only, so every basis vector stays ghost-consistent and the apply needs no P exchange.
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
   if (m + 2 > CDM_RED_MAXK) { return cdm_fail(c, CDM_EINVAL, "cdm_gmres: restart too large"); }
   RC(ensure_ws(op, (int64_t)(m + 3) * ld));
   double *V = op->kry_dev, *w = V + (int64_t)(m + 1) * ld, *t = w + ld;
   const bool par = c->nranks > 1 && !sp->peers.empty();
   if (par && op->halo_mode == 2 && op->assembly == 0 && sp->sym.ready == 0) { RC(cdm_halo_sym_setup(sp)); }
   const bool gh = cdm_apply_keeps_ghosts(op);            // applies return ghost-consistent vectors
   const int64_t nall = gh ? sp->ndof : n, ntail = nall - n;
   // preconditioner (-pc_type): 0 none, 1 jacobi, 2 block-Jacobi + ILU(0) on the assembled matrix (one block per rank;
   // single rank only, like the assembled path itself)
   const double *dinv = nullptr;
   const bool ilu = o->jacobi == 2;
   if (o->jacobi == 1) { RC(ensure_dinv(op)); dinv = op->dinv_dev; }
   if (ilu)
   {
      if (par) { return cdm_fail(c, CDM_EUNSUP, "cdm_gmres: the ILU(0) preconditioner is single-rank"); }
      if (!op->csr) { RC(cdm_operator_assemble_csr(op)); }
      op->assembly = 1;                                   // ILU(0) belongs to the assembled matrix: apply = SpMV
      RC(cdm_ilu_setup(op));
   }
   // z = M^{-1} r on the true dofs (initial residuals; the iteration itself fuses Jacobi into the multi-dot)
   auto precond = [&](const double *r, double *z) -> int
   { return ilu ? cdm_ilu_apply(op, r, z) : cdm_k_pmult_scaled(c, n, dinv, nullptr, r, z); };
   double *h_dev = red_out(c);                            // [m+1] dots of the current iteration
   double *nrm2 = red_out(c) + CDM_RED_MAXK;              // [m+2] squared norms of u_0 .. u_m
   double *y_dev = red_out(c) + 2 * CDM_RED_MAXK;
   std::vector<double> H((size_t)(m + 1) * m, 0.0), cs(m), sn(m), s(m + 1), yv(m), hc(m + 2), unorm(m + 2, 1.0);
   int it = 0, conv = 0, hl = 0;
   double rnorm = 0.0, ttol = 0.0;
   bool first = true;
   if (!c->ev_kry) { CDM_CUDA(c, cudaEventCreateWithFlags(&c->ev_kry, cudaEventDisableTiming)); }
   CDM_CUDA(c, cudaEventRecord(c->ev0, c->stream));
   if (o->zero_guess) { RC(cdm_k_set(c, n, 0.0, x)); }
   while (true)
   {
      // u_0 = M^{-1}(b - A x)
      if (first && o->zero_guess) { RC(precond(b, V)); }
      else
      {
         CDM_CUDA(c, cudaMemcpyAsync(w, x, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
         RC(cdm_apply_tail(op, w, t, true, false));
         RC(cdm_k_add(c, n, b, -1.0, t, t));
         RC(precond(t, V));
      }
      if (gh) { RC(cdm_halo_P_space(sp, V)); }             // once per cycle: make u_0 ghost-consistent
      RC(cdm_k_mdot_dev(c, n, 1, V, V, ld, nrm2));
      RC(cdm_allreduce_sum(c, nrm2, 1));
      double b2; RC(fetch(c, nrm2, 1, &b2));
      const double beta = std::sqrt(b2);
      unorm[0] = beta;
      rnorm = beta;
      if (first)
      {
         // KSPConvergedDefault: with a non-zero initial guess PETSc scales rtol by ||M^{-1} b||, the preconditioned
         // norm of the right-hand side; mfem::GMRESSolver always uses the initial residual
         double ref = beta;
         if (!o->zero_guess && o->variant == CDM_GMRES_PETSC)
         {
            RC(precond(b, w));
            RC(cdm_k_mdot_dev(c, n, 1, w, w, ld, h_dev));
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
      std::fill(s.begin(), s.end(), 0.0);
      s[0] = beta;
      int j = 0;
      bool applied = false;                                // t already holds A u_j (queued ahead of the host wait)
      while (j < m && it < o->max_it)
      {
         double *uj = V + (int64_t)j * ld, *un = V + (int64_t)(j + 1) * ld;
         if (!applied) { RC(cdm_apply_tail(op, uj, t, true, gh)); }
         applied = false;
         const double *pre = t;                            // M^{-1} A u_j when the preconditioner is not the fused Jacobi
         if (ilu) { RC(cdm_ilu_apply(op, t, w)); pre = w; }
         if (o->variant == CDM_GMRES_PETSC)
         {
            // classical Gram-Schmidt: all j+1 dots against the same vector, one all-reduce
            RC(cdm_k_mdot_lazy_dev(c, n, j + 1, pre, dinv, nrm2 + j, un, V, ld, nrm2, h_dev));
            if (ntail > 0) { RC(cdm_k_pmult_scaled(c, ntail, dinv ? dinv + n : nullptr, nrm2 + j, t + n, un + n)); }
            RC(cdm_allreduce_sum(c, h_dev, j + 1));
            RC(cdm_k_maxpy_lazy_dev(c, n, j + 1, h_dev, nrm2, V, ld, un, nrm2 + (j + 1)));
            if (ntail > 0) { RC(cdm_k_maxpy_lazy_dev(c, ntail, j + 1, h_dev, nrm2, V + n, ld, un + n, nullptr)); }
         }
         else
         {
            // modified Gram-Schmidt (mfem::GMRESSolver), scalars stay on the device
            RC(cdm_k_pmult_scaled(c, nall, dinv, nrm2 + j, pre, un));
            for (int i = 0; i <= j; i++)
            {
               RC(cdm_k_mdot_dev(c, n, 1, un, V + (int64_t)i * ld, ld, h_dev + i));
               RC(cdm_k_scale_dots(c, 1, nrm2 + i, h_dev + i));
               RC(cdm_allreduce_sum(c, h_dev + i, 1));
               RC(cdm_k_maxpy_lazy_dev(c, n, 1, h_dev + i, nrm2 + i, V + (int64_t)i * ld, ld, un, i == j ? nrm2 + (j + 1) : nullptr));
               if (ntail > 0) { RC(cdm_k_maxpy_lazy_dev(c, ntail, 1, h_dev + i, nrm2 + i, V + (int64_t)i * ld + n, ld, un + n, nullptr)); }
            }
         }
         RC(cdm_allreduce_sum(c, nrm2 + (j + 1), 1));
         CDM_CUDA(c, cudaMemcpyAsync(c->red_host, h_dev, sizeof(double) * (j + 1), cudaMemcpyDeviceToHost, c->stream));
         CDM_CUDA(c, cudaMemcpyAsync(c->red_host + (j + 1), nrm2 + (j + 1), sizeof(double), cudaMemcpyDeviceToHost, c->stream));
         CDM_CUDA(c, cudaEventRecord(c->ev_kry, c->stream));
         // the next apply does not depend on the host: queue it before waiting (not across a restart / the last step)
         if (j + 1 < m && it + 1 < o->max_it) { RC(cdm_apply_tail(op, un, t, true, gh)); applied = true; }
         CDM_CUDA(c, cudaEventSynchronize(c->ev_kry));
         RC(cdm_check_p2p(c));
         for (int i = 0; i < j + 2; i++) { hc[i] = c->red_host[i]; }
         const double hn = std::sqrt(hc[j + 1]);
         unorm[j + 1] = hn;
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
      // x += V y = sum_i (y_i / ||u_i||) u_i  (maxpy subtracts, so upload the negated coefficients)
      for (int i = 0; i < j; i++) { c->red_host[2 * CDM_RED_MAXK + i] = -yv[i] / unorm[i]; }
      CDM_CUDA(c, cudaMemcpyAsync(y_dev, c->red_host + 2 * CDM_RED_MAXK, sizeof(double) * j, cudaMemcpyHostToDevice, c->stream));
      RC(cdm_k_maxpy_dev(c, n, j, y_dev, V, ld, x, nullptr));
      CDM_CUDA(c, cudaStreamSynchronize(c->stream));   // red_host is reused
      if (conv || it >= o->max_it) { break; }
   }
   CDM_CUDA(c, cudaEventRecord(c->ev1, c->stream));
   CDM_CUDA(c, cudaEventSynchronize(c->ev1));
   RC(cdm_check_p2p(c));
   if (ilu) { RC(cdm_ilu_check(op)); }
   float ms = 0.f;
   cudaEventElapsedTime(&ms, c->ev0, c->ev1);
   res->iters = it; res->converged = conv; res->final_norm = rnorm; res->hist_len = hl;
   res->seconds = ms * 1e-3;
   return CDM_OK;
}

// mfem::CGSolver.  On partitioned spaces r, z, d are kept ghost-consistent (updates over the whole local vector, dots
// over the true dofs), so the apply inside the loop needs no P exchange.
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
   const bool par = c->nranks > 1 && !sp->peers.empty();
   if (par && op->halo_mode == 2 && op->assembly == 0 && sp->sym.ready == 0) { RC(cdm_halo_sym_setup(sp)); }
   const bool gh = cdm_apply_keeps_ghosts(op);
   const int64_t nall = gh ? sp->ndof : n, ntail = nall - n;
   const double *dinv = nullptr;
   if (o->jacobi == 2) { return cdm_fail(c, CDM_EUNSUP, "cdm_cg: only none / Jacobi preconditioning (ILU(0) is wired into cdm_gmres)"); }
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
      RC(cdm_apply_tail(op, d, r, true, false));
      RC(cdm_k_add(c, n, b, -1.0, r, r));
   }
   if (gh) { RC(cdm_halo_P_space(sp, r)); }                 // r ghost-consistent from here on
   if (dinv) { RC(cdm_k_pmult(c, nall, dinv, r, d)); }
   else { CDM_CUDA(c, cudaMemcpyAsync(d, r, sizeof(double) * nall, cudaMemcpyDeviceToDevice, c->stream)); }
   double nom, den, betanom;
   RC(cdm_k_mdot_dev(c, n, 1, d, r, ld, sc)); RC(cdm_allreduce_sum(c, sc, 1)); RC(fetch(c, sc, 1, &nom));
   const double r0 = std::fmax(nom * o->rtol * o->rtol, o->atol * o->atol);
   if (hist) { hist[hl] = nom; } hl++;
   betanom = nom;
   if (nom <= r0) { conv = 1; }
   else
   {
      RC(cdm_apply_tail(op, d, z, true, gh));
      RC(cdm_k_mdot_dev(c, n, 1, z, d, ld, sc)); RC(cdm_allreduce_sum(c, sc, 1)); RC(fetch(c, sc, 1, &den));
      if (den > 0.0)
      {
         for (it = 1;; it++)
         {
            const double alpha = nom / den;
            RC(cdm_k_cg_update(c, n, alpha, d, z, x, r, sc));       // x += a d ; r -= a z ; sc = (r,r)
            if (ntail > 0) { RC(cdm_k_axpy(c, ntail, -alpha, z + n, r + n)); }
            if (dinv)
            {
               RC(cdm_k_pmult(c, nall, dinv, r, zz));
               RC(cdm_k_mdot_dev(c, n, 1, r, zz, ld, sc));
            }
            RC(cdm_allreduce_sum(c, sc, 1));
            RC(fetch(c, sc, 1, &betanom));
            if (hist) { hist[hl] = betanom; } hl++;
            if (betanom <= r0) { conv = 1; break; }
            if (it >= o->max_it) { break; }
            const double beta = betanom / nom;
            RC(cdm_k_add(c, nall, dinv ? zz : r, beta, d, d));       // d = z + beta d
            RC(cdm_apply_tail(op, d, z, true, gh));
            RC(cdm_k_mdot_dev(c, n, 1, d, z, ld, sc)); RC(cdm_allreduce_sum(c, sc, 1)); RC(fetch(c, sc, 1, &den));
            if (den <= 0.0) { break; }
            nom = betanom;
         }
      }
   }
   CDM_CUDA(c, cudaEventRecord(c->ev1, c->stream));
   CDM_CUDA(c, cudaEventSynchronize(c->ev1));
   RC(cdm_check_p2p(c));
   float ms = 0.f;
   cudaEventElapsedTime(&ms, c->ev0, c->ev1);
   res->iters = it; res->converged = conv; res->final_norm = std::sqrt(std::fabs(betanom)); res->hist_len = hl;
   res->seconds = ms * 1e-3;
   return CDM_OK;
}
