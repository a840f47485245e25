// kernels_common.cuh -- small device-side helpers shared by the kernel files.
#pragma once
#include "cdm_internal.hpp"
#include <cstring>

// 1-D basis tables, passed by value as a kernel parameter (constant bank):
// B[q*D1D + d] = l_d(x_q), G[q*D1D + d] = l_d'(x_q)
struct BasisTables
{
   double B[CDM_MAX_Q1D * CDM_MAX_D1D];
   double G[CDM_MAX_Q1D * CDM_MAX_D1D];
};

// fixed launch shape of every reduction: results do not depend on n-independent
// scheduling, so dot products are bit-reproducible run to run
constexpr int CDM_RED_BLOCKS = 592;     // 4 x 148 SMs
constexpr int CDM_RED_THREADS = 256;
constexpr int CDM_RED_MAXK = 64;

__device__ __forceinline__ double warp_sum(double v)
{
   v += __shfl_xor_sync(0xffffffffu, v, 16);
   v += __shfl_xor_sync(0xffffffffu, v, 8);
   v += __shfl_xor_sync(0xffffffffu, v, 4);
   v += __shfl_xor_sync(0xffffffffu, v, 2);
   v += __shfl_xor_sync(0xffffffffu, v, 1);
   return v;
}

// block-wide sum (CDM_RED_THREADS threads); result valid in thread 0
__device__ __forceinline__ double block_sum(double v, double *smem8)
{
   v = warp_sum(v);
   const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
   if (lane == 0) { smem8[wid] = v; }
   __syncthreads();
   double r = 0.0;
   if (wid == 0)
   {
      r = (lane < (int)(blockDim.x >> 5)) ? smem8[lane] : 0.0;
      r = warp_sum(r);
   }
   __syncthreads();
   return r;
}

// ---- bulk-async copy (TMA, cp.async.bulk) + mbarrier helpers shared by the apply kernels
namespace cdmk
{
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
   asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "CW_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra CD_%=;\n"
      "bra CW_%=;\n"
      "CD_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy completing on an mbarrier; L2 evict-first: a stream that is read once must not push the
// gathered vectors out of L2
__device__ __forceinline__ void bulk_g2s_stream(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
   uint64_t pol;
   asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
   asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void red_add_f64(double *addr, double v)
{ asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory"); }
// y[g] += v unless g < 0 (essential dof), as one predicated instruction
__device__ __forceinline__ void red_add_f64_pred(double *y, int g, double v)
{
   asm volatile("{\n.reg .pred p;\nsetp.ge.s32 p, %2, 0;\n@p red.global.add.f64 [%0], %1;\n}\n"
                ::"l"(y + g), "d"(v), "r"(g) : "memory");
}
}  // namespace cdmk
