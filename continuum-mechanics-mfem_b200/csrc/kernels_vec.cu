// kernels_vec.cu -- Krylov vector kernels: axpy-type updates, fused multi-dot,
// fused multi-axpy + norm, fused CG update.  All reductions are two-stage with
// a fixed launch shape and fixed summation order (bit-reproducible).
//
// Stands behind the mfem::Vector kernels used by mfem::CGSolver
// (mesh_recession_handler.cpp:270-276), PETSc VecMDot / VecMAXPY / VecNorm inside
// KSPGMRES (linear_convection_diffusion_2D.cpp:368-370, Input/petsc.opts:2) and
// InnerProduct (newton_petsc_solver.hpp:82-85).
#include "cdm_internal.hpp"
#include "kernels_common.cuh"

#define VEC_BLOCK 256
static inline unsigned vec_grid(int64_t n, int per_thread = 1)
{
   int64_t nb = (n + (int64_t)VEC_BLOCK * per_thread - 1) / ((int64_t)VEC_BLOCK * per_thread);
   return (unsigned)(nb < 1 ? 1 : nb);
}
#define VEC_CHECK(c) do { (c)->launches++; CDM_CUDA(c, cudaGetLastError()); return CDM_OK; } while (0)

__global__ void k_set(int64_t n, double v, double *__restrict__ x)
{
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) { x[i] = v; }
}
__global__ void k_axpy(int64_t n, double a, const double *__restrict__ x, double *__restrict__ y)
{
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) { y[i] += a * x[i]; }
}
__global__ void k_add(int64_t n, const double *__restrict__ x, double a, const double *__restrict__ y,
                      double *__restrict__ z)
{
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) { z[i] = x[i] + a * y[i]; }
}
__global__ void k_pmult(int64_t n, const double *__restrict__ d, const double *__restrict__ x, double *__restrict__ y)
{
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) { y[i] = d[i] * x[i]; }
}
__global__ void k_recip(int64_t n, const double *__restrict__ d, double *__restrict__ r)
{
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) { r[i] = 1.0 / d[i]; }
}
__global__ void k_scale(int64_t n, double a, const double *__restrict__ w, double *__restrict__ v)
{
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) { v[i] = a * w[i]; }
}
__global__ void k_scale_rnorm(int64_t n, const double *__restrict__ norm2, const double *__restrict__ w,
                              double *__restrict__ v)
{
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   const double a = 1.0 / sqrt(norm2[0]);
   if (i < n) { v[i] = a * w[i]; }
}
__global__ void k_copy_idx(int64_t n, const int32_t *__restrict__ idx, const double *__restrict__ x, double *__restrict__ y)
{
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) { const int32_t g = idx[i]; y[g] = x[g]; }
}
__global__ void k_set_idx(int64_t n, const int32_t *__restrict__ idx, double v, double *__restrict__ y)
{
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) { y[idx[i]] = v; }
}
__global__ void k_pack(int64_t n, const int32_t *__restrict__ idx, const double *__restrict__ x, double *__restrict__ buf)
{
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) { buf[i] = x[idx[i]]; }
}
__global__ void k_unpack(int64_t n, const int32_t *__restrict__ idx, const double *__restrict__ buf,
                         double *__restrict__ x, int add)
{
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) { if (add) { x[idx[i]] += buf[i]; } else { x[idx[i]] = buf[i]; } }
}

__global__ void k_unpack_add_csr(int64_t n, const int32_t *__restrict__ dof, const int32_t *__restrict__ off,
                                 const int32_t *__restrict__ src, const double *__restrict__ buf, double *__restrict__ x)
{
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) { return; }
   double s = x[dof[i]];
   for (int32_t j = off[i]; j < off[i + 1]; j++) { s += buf[src[j]]; }
   x[dof[i]] = s;
}
int cdm_k_unpack_add_csr(cdm_ctx *c, int64_t n, const int32_t *dof, const int32_t *off, const int32_t *src,
                         const double *buf, double *x)
{ if (n > 0) { k_unpack_add_csr<<<vec_grid(n), VEC_BLOCK, 0, c->stream>>>(n, dof, off, src, buf, x); } VEC_CHECK(c); }

int cdm_k_set(cdm_ctx *c, int64_t n, double v, double *x)
{ if (n > 0) { k_set<<<vec_grid(n), VEC_BLOCK, 0, c->stream>>>(n, v, x); } VEC_CHECK(c); }
int cdm_k_axpy(cdm_ctx *c, int64_t n, double a, const double *x, double *y)
{ if (n > 0) { k_axpy<<<vec_grid(n), VEC_BLOCK, 0, c->stream>>>(n, a, x, y); } VEC_CHECK(c); }
int cdm_k_add(cdm_ctx *c, int64_t n, const double *x, double a, const double *y, double *z)
{ if (n > 0) { k_add<<<vec_grid(n), VEC_BLOCK, 0, c->stream>>>(n, x, a, y, z); } VEC_CHECK(c); }
int cdm_k_pmult(cdm_ctx *c, int64_t n, const double *d, const double *x, double *y)
{ if (n > 0) { k_pmult<<<vec_grid(n), VEC_BLOCK, 0, c->stream>>>(n, d, x, y); } VEC_CHECK(c); }
int cdm_k_recip(cdm_ctx *c, int64_t n, const double *d, double *r)
{ if (n > 0) { k_recip<<<vec_grid(n), VEC_BLOCK, 0, c->stream>>>(n, d, r); } VEC_CHECK(c); }
int cdm_k_scale(cdm_ctx *c, int64_t n, double a, const double *w, double *v)
{ if (n > 0) { k_scale<<<vec_grid(n), VEC_BLOCK, 0, c->stream>>>(n, a, w, v); } VEC_CHECK(c); }
int cdm_k_scale_by_rnorm(cdm_ctx *c, int64_t n, const double *norm2_dev, const double *w, double *v)
{ if (n > 0) { k_scale_rnorm<<<vec_grid(n), VEC_BLOCK, 0, c->stream>>>(n, norm2_dev, w, v); } VEC_CHECK(c); }
int cdm_k_copy_idx(cdm_ctx *c, int64_t n, const int32_t *idx, const double *x, double *y)
{ if (n > 0) { k_copy_idx<<<vec_grid(n), VEC_BLOCK, 0, c->stream>>>(n, idx, x, y); } VEC_CHECK(c); }
int cdm_k_set_idx(cdm_ctx *c, int64_t n, const int32_t *idx, double v, double *y)
{ if (n > 0) { k_set_idx<<<vec_grid(n), VEC_BLOCK, 0, c->stream>>>(n, idx, v, y); } VEC_CHECK(c); }
int cdm_k_zero_idx(cdm_ctx *c, int64_t n, const int32_t *idx, double *y) { return cdm_k_set_idx(c, n, idx, 0.0, y); }
int cdm_k_pack(cdm_ctx *c, int64_t n, const int32_t *idx, const double *x, double *buf)
{ if (n > 0) { k_pack<<<vec_grid(n), VEC_BLOCK, 0, c->stream>>>(n, idx, x, buf); } VEC_CHECK(c); }
int cdm_k_unpack(cdm_ctx *c, int64_t n, const int32_t *idx, const double *buf, double *x, int add)
{ if (n > 0) { k_unpack<<<vec_grid(n), VEC_BLOCK, 0, c->stream>>>(n, idx, buf, x, add); } VEC_CHECK(c); }

// ------------------------------------------------------------ reductions
// stage 1: per-block partials, partial[j * CDM_RED_BLOCKS + block]
// stage 2: one block per result sums its CDM_RED_BLOCKS partials in a fixed order

template <int KB>
__global__ void __launch_bounds__(CDM_RED_THREADS)
k_mdot_partial(int64_t n, int kb, const double *__restrict__ w, const double *__restrict__ V, int64_t ldv,
               double *__restrict__ partial)
{
   __shared__ double red[8];
   double acc[KB];
   #pragma unroll
   for (int j = 0; j < KB; j++) { acc[j] = 0.0; }
   const int64_t stride = (int64_t)gridDim.x * blockDim.x;
   for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
   {
      const double wv = w[i];
      #pragma unroll
      for (int j = 0; j < KB; j++) if (j < kb) { acc[j] += wv * V[j * ldv + i]; }
   }
   #pragma unroll
   for (int j = 0; j < KB; j++)
   {
      if (j < kb)
      {
         const double s = block_sum(acc[j], red);
         if (threadIdx.x == 0) { partial[(int64_t)j * CDM_RED_BLOCKS + blockIdx.x] = s; }
      }
   }
}

__global__ void __launch_bounds__(CDM_RED_THREADS)
k_reduce_final(const double *__restrict__ partial, double *__restrict__ out)
{
   __shared__ double red[8];
   const double *p = partial + (int64_t)blockIdx.x * CDM_RED_BLOCKS;
   double v = 0.0;
   for (int i = threadIdx.x; i < CDM_RED_BLOCKS; i += CDM_RED_THREADS) { v += p[i]; }
   const double s = block_sum(v, red);
   if (threadIdx.x == 0) { out[blockIdx.x] = s; }
}

int cdm_k_mdot_dev(cdm_ctx *c, int64_t n, int k, const double *w, const double *V, int64_t ldv, double *out_dev)
{
   if (k < 1 || k > CDM_RED_MAXK) { return cdm_fail(c, CDM_EINVAL, "mdot: k out of range"); }
   double *partial = c->red_dev;
   for (int j0 = 0; j0 < k; j0 += 8)
   {
      const int kb = (k - j0 < 8) ? k - j0 : 8;
      k_mdot_partial<8><<<CDM_RED_BLOCKS, CDM_RED_THREADS, 0, c->stream>>>(
         n, kb, w, V + (int64_t)j0 * ldv, ldv, partial + (int64_t)j0 * CDM_RED_BLOCKS);
      c->launches++;
   }
   k_reduce_final<<<k, CDM_RED_THREADS, 0, c->stream>>>(partial, out_dev);
   VEC_CHECK(c);
}

// w -= sum_j h[j] V_j, and the partial sums of ||w_new||^2
__global__ void __launch_bounds__(CDM_RED_THREADS)
k_maxpy_norm(int64_t n, int k, const double *__restrict__ h, const double *__restrict__ V, int64_t ldv,
             double *__restrict__ w, double *__restrict__ partial)
{
   __shared__ double sh[CDM_RED_MAXK];
   __shared__ double red[8];
   if (threadIdx.x < k) { sh[threadIdx.x] = h[threadIdx.x]; }
   __syncthreads();
   double acc = 0.0;
   const int64_t stride = (int64_t)gridDim.x * blockDim.x;
   for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
   {
      double v = w[i];
      for (int j = 0; j < k; j++) { v -= sh[j] * V[j * ldv + i]; }
      w[i] = v;
      acc += v * v;
   }
   if (partial)
   {
      const double s = block_sum(acc, red);
      if (threadIdx.x == 0) { partial[blockIdx.x] = s; }
   }
}

int cdm_k_maxpy_dev(cdm_ctx *c, int64_t n, int k, const double *h_dev, const double *V, int64_t ldv,
                    double *w, double *norm2_out_dev)
{
   if (k < 0 || k > CDM_RED_MAXK) { return cdm_fail(c, CDM_EINVAL, "maxpy: k out of range"); }
   k_maxpy_norm<<<CDM_RED_BLOCKS, CDM_RED_THREADS, 0, c->stream>>>(n, k, h_dev, V, ldv, w,
                                                                  norm2_out_dev ? c->red_dev : nullptr);
   c->launches++;
   if (norm2_out_dev)
   {
      k_reduce_final<<<1, CDM_RED_THREADS, 0, c->stream>>>(c->red_dev, norm2_out_dev);
      c->launches++;
   }
   CDM_CUDA(c, cudaGetLastError());
   return CDM_OK;
}

// x += a d ; r -= a z ; partial sums of (r_new, r_new)
__global__ void __launch_bounds__(CDM_RED_THREADS)
k_cg_update(int64_t n, double a, const double *__restrict__ d, const double *__restrict__ z,
            double *__restrict__ x, double *__restrict__ r, double *__restrict__ partial)
{
   __shared__ double red[8];
   double acc = 0.0;
   const int64_t stride = (int64_t)gridDim.x * blockDim.x;
   for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
   {
      x[i] += a * d[i];
      const double rv = r[i] - a * z[i];
      r[i] = rv;
      acc += rv * rv;
   }
   const double s = block_sum(acc, red);
   if (threadIdx.x == 0) { partial[blockIdx.x] = s; }
}

int cdm_k_cg_update(cdm_ctx *c, int64_t n, double a, const double *d, const double *z, double *x,
                    double *r, double *rr_out_dev)
{
   k_cg_update<<<CDM_RED_BLOCKS, CDM_RED_THREADS, 0, c->stream>>>(n, a, d, z, x, r, c->red_dev);
   k_reduce_final<<<1, CDM_RED_THREADS, 0, c->stream>>>(c->red_dev, rr_out_dev);
   c->launches += 2;
   CDM_CUDA(c, cudaGetLastError());
   return CDM_OK;
}
