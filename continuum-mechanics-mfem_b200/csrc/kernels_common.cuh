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
