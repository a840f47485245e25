// integrator_pa.cu -- entry points at the BilinearFormIntegrator / prolongation level of MFEM's
// partial-assembly stack, for callers that keep MFEM's own ElementRestriction and
// ParFiniteElementSpace and only want the element kernels (north star: "keeps the
// mfem::Operator::Mult / BilinearFormIntegrator-style API"):
//
//   BilinearFormIntegrator::AddMultPA(x_E, y_E)        -> cdm_integrator_add_mult_pa
//   BilinearFormIntegrator::AssembleDiagonalPA(diag_E) -> cdm_integrator_assemble_diagonal_pa
//   ParBilinearForm::RecoverFEMSolution (u_L = P X, linear_convection_diffusion_2D.cpp:377,
//   per step at linear_convection_diffusion_1D.cpp:569-572)        -> cdm_prolongate
//   P^T (ParLinearForm::ParallelAssemble, linear_convection_diffusion_2D.cpp:343) -> cdm_prolongate_transpose
//
// plus two per-context services of the launchers (kernel configuration cache, peer-memory error word).
#include "cdm_internal.hpp"
#include "kernels_common.cuh"

int cdm_kernel_cfg(cdm_ctx *ctx, const void *kern, int threads, size_t smem, const char *name, int *blocks_per_sm)
{
   const auto key = std::make_pair(kern, smem);
   auto it = ctx->kernel_cfg.find(key);
   if (it != ctx->kernel_cfg.end()) { *blocks_per_sm = it->second; return CDM_OK; }
   // both calls act on the CURRENT device: a context created for another device of the same process
   // needs its own opt-in and its own occupancy answer
   CDM_CUDA(ctx, cudaSetDevice(ctx->device));
   CDM_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
   int b = 0;
   CDM_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, threads, smem));
   if (b < 1) { return cdm_fail(ctx, CDM_ECUDA, std::string(name) + " does not fit on an SM"); }
   if (getenv("CDM_CFG_DEBUG"))                              // resident blocks per SM of every kernel configuration, once each
   {
      cudaFuncAttributes fa;
      if (cudaFuncGetAttributes(&fa, kern) == cudaSuccess)
      {
         fprintf(stderr, "[cdm] %s: %d threads, %zu B dynamic shared memory, %d registers, %zu B local -> %d blocks per SM\n",
                 name, threads, smem, fa.numRegs, (size_t)fa.localSizeBytes, b);
      }
   }
   ctx->kernel_cfg[key] = b;
   *blocks_per_sm = b;
   return CDM_OK;
}

int cdm_check_p2p(cdm_ctx *ctx)
{
   if (!ctx->p2p_err_host || *ctx->p2p_err_host == 0u) { return CDM_OK; }
   const unsigned int w = *ctx->p2p_err_host;
   *ctx->p2p_err_host = 0u;
   return cdm_fail(ctx, CDM_ENCCL, "peer-memory halo exchange timed out waiting for neighbour slot " + std::to_string(w - 1) +
                                   " (a rank died or the ranks issued different call sequences); the result of the last apply is invalid");
}

namespace
{
__global__ void __launch_bounds__(256) k_iota(int64_t n, int32_t *a)
{
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) { a[i] = (int32_t)i; }
}

int ensure_iota(cdm_space *sp)
{
   if (sp->iota_dev) { return CDM_OK; }
   cdm_ctx *ctx = sp->ctx;
   const int64_t n = sp->ne * sp->nd;
   if (n > 2147483000LL) { return cdm_fail(ctx, CDM_EUNSUP, "E-vector entry points: more than 2^31 element dofs"); }
   CDM_CUDA(ctx, cudaMalloc(&sp->iota_dev, sizeof(int32_t) * (size_t)n));
   k_iota<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, sp->iota_dev);
   ctx->launches++;
   CDM_CUDA(ctx, cudaGetLastError());
   return CDM_OK;
}
}  // namespace

extern "C" {

// y_E += B^T D B x_E  (all active integrators fused; E-vectors are [ne][nd], dof-fastest, MFEM ElementRestriction layout)
int cdm_integrator_add_mult_pa(cdm_op *op, const double *xE_dev, double *yE_dev)
{
   if (!op || !xE_dev || !yE_dev) { return CDM_EINVAL; }
   cdm_space *sp = op->sp;
   cdm_ctx *ctx = sp->ctx;
   CDM_REQUIRE_GPU(ctx);
   if (op->assembly == 1) { return cdm_fail(ctx, CDM_EUNSUP, "cdm_integrator_add_mult_pa: operator is in full-assembly mode"); }
   int rc;
   if ((rc = ensure_iota(sp))) { return rc; }
   if ((rc = cdm_k_ensure_yE(op))) { return rc; }
   // the element kernels read x through a gather map and, in the deterministic scatter mode, write plain E-vectors:
   // identity map in, scratch E-vector out, then accumulate (AddMultPA adds)
   const int saved_scatter = op->scatter_mode;
   const bool saved_range = op->range_on;
   op->scatter_mode = 0; op->range_on = false;
   op->gmap_override = sp->iota_dev; op->e_out = op->yE_dev;
   rc = cdm_k_apply(op, xE_dev, nullptr, false);
   op->gmap_override = nullptr; op->e_out = nullptr;
   op->scatter_mode = saved_scatter; op->range_on = saved_range;
   if (rc) { return rc; }
   return cdm_k_axpy(ctx, sp->ne * sp->nd, 1.0, op->yE_dev, yE_dev);
}

// diag_E += diag(B^T D B) element by element
int cdm_integrator_assemble_diagonal_pa(cdm_op *op, double *diagE_dev)
{
   if (!op || !diagE_dev) { return CDM_EINVAL; }
   cdm_space *sp = op->sp;
   cdm_ctx *ctx = sp->ctx;
   CDM_REQUIRE_GPU(ctx);
   int rc;
   if ((rc = cdm_k_ensure_yE(op))) { return rc; }
   if ((rc = cdm_k_diag_evec(op, op->yE_dev))) { return rc; }
   return cdm_k_axpy(ctx, sp->ne * sp->nd, 1.0, op->yE_dev, diagE_dev);
}

// ElementRestriction::Mult / MultTranspose on the device, for callers that work with E-vectors:
// xE[e*nd + d] = xL[gather[e*nd + d]]  ;  yL[g] = sum_{j in [offsets[g], offsets[g+1])} yE[indices[j]]
int cdm_restriction_mult(cdm_space *sp, const double *xL_dev, double *xE_dev)
{
   if (!sp || !xL_dev || !xE_dev) { return CDM_EINVAL; }
   CDM_REQUIRE_GPU(sp->ctx);
   return cdm_k_pack(sp->ctx, sp->ne * sp->nd, sp->gather_dev, xL_dev, xE_dev);
}
int cdm_restriction_mult_transpose(cdm_space *sp, const double *yE_dev, double *yL_dev)
{
   if (!sp || !yE_dev || !yL_dev) { return CDM_EINVAL; }
   CDM_REQUIRE_GPU(sp->ctx);
   return cdm_k_restrict_transpose(sp, yE_dev, yL_dev);
}

// u_L = P x_T : the true dofs are the first ntrue entries of the L-vector (owned-first numbering), the ghost
// tail is filled from the owners.  u_L has cdm_operator_local_size entries; x_T may alias u_L.
int cdm_prolongate(cdm_space *sp, const double *xT_dev, double *uL_dev)
{
   if (!sp || !xT_dev || !uL_dev) { return CDM_EINVAL; }
   cdm_ctx *ctx = sp->ctx;
   CDM_REQUIRE_GPU(ctx);
   if (xT_dev != uL_dev)
      CDM_CUDA(ctx, cudaMemcpyAsync(uL_dev, xT_dev, sizeof(double) * (size_t)sp->ntrue, cudaMemcpyDeviceToDevice, ctx->stream));
   if (ctx->nranks > 1 && !sp->peers.empty()) { return cdm_halo_P_space(sp, uL_dev); }
   return CDM_OK;
}

// b_T = P^T b_L : ghost contributions are summed into their owners (fixed peer order); b_L is modified in place
// (its first ntrue entries are the result) and, when bT_dev differs, copied out.
int cdm_prolongate_transpose(cdm_space *sp, double *bL_dev, double *bT_dev)
{
   if (!sp || !bL_dev) { return CDM_EINVAL; }
   cdm_ctx *ctx = sp->ctx;
   CDM_REQUIRE_GPU(ctx);
   if (ctx->nranks > 1 && !sp->peers.empty()) { const int rc = cdm_halo_PT_space(sp, bL_dev); if (rc) { return rc; } }
   if (bT_dev && bT_dev != bL_dev)
      CDM_CUDA(ctx, cudaMemcpyAsync(bT_dev, bL_dev, sizeof(double) * (size_t)sp->ntrue, cudaMemcpyDeviceToDevice, ctx->stream));
   return CDM_OK;
}

int64_t cdm_space_local_size(const cdm_space *sp) { return sp ? sp->ndof : 0; }
int64_t cdm_space_true_size(const cdm_space *sp) { return sp ? sp->ntrue : 0; }

}  // extern "C"
