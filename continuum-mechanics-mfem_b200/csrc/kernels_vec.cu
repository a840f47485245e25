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
   if (i < n) { const double v = d[i]; r[i] = (v != 0.0) ? 1.0 / v : 1.0; }      // PCJacobi: zero diagonal -> 1
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
// Streaming kernels use 128-bit loads and keep KB+1 (or more) independent loads in flight per
// thread; the launch shape is fixed, so every result is bit-reproducible.

__device__ __forceinline__ double2 ld2(const double *p, int64_t i2) { return reinterpret_cast<const double2 *>(p)[i2]; }
__device__ __forceinline__ void st2(double *p, int64_t i2, double2 v) { reinterpret_cast<double2 *>(p)[i2] = v; }

// KB dot products of w against V_0..V_{KB-1} (KB is exact: no predicates in the steady state).
// PC: w = dinv .* t is formed on the fly (Jacobi preconditioner fused into the first pass) and
// written to w_out.  VEC: all pointers 16-byte aligned and ldv even -> double2 path, where a block
// walks over tiles of 2*UNROLL*blockDim consecutive doubles of every vector.
template <int KB, int UNROLL, bool PC, bool VEC>
__global__ void __launch_bounds__(CDM_RED_THREADS)
k_mdot_partial(int64_t n, const double *__restrict__ w, const double *__restrict__ dinv,
               double *__restrict__ w_out, const double *__restrict__ V, int64_t ldv,
               double *__restrict__ partial, const double *__restrict__ sc2)
{
   __shared__ double red[8];
   // PC: w_out = sc * dinv .* w with sc = 1/sqrt(*sc2) (lazy normalisation of the Krylov basis: the operator was
   // applied to the un-normalised vector u_j, sc2 = ||u_j||^2); dinv == nullptr stands for the identity
   const double sc = (PC && sc2) ? 1.0 / sqrt(*sc2) : 1.0;
   double acc[KB];
   #pragma unroll
   for (int j = 0; j < KB; j++) { acc[j] = 0.0; }
   if (VEC)
   {
      const int64_t n2 = n >> 1;
      const int64_t tile = (int64_t)UNROLL * blockDim.x;               // in double2 units
      const int64_t nfull = n2 / tile;
      for (int64_t t = blockIdx.x; t < nfull; t += gridDim.x)
      {
         const int64_t base = t * tile + threadIdx.x;
         double2 wv[UNROLL];
         #pragma unroll
         for (int u = 0; u < UNROLL; u++)
         {
            wv[u] = ld2(w, base + u * blockDim.x);
            if (PC)
            {
               const double2 d = dinv ? ld2(dinv, base + u * blockDim.x) : make_double2(1.0, 1.0);
               wv[u].x *= sc * d.x; wv[u].y *= sc * d.y; st2(w_out, base + u * blockDim.x, wv[u]);
            }
         }
         double2 v[KB][UNROLL];
         #pragma unroll
         for (int j = 0; j < KB; j++)
         {
            #pragma unroll
            for (int u = 0; u < UNROLL; u++) { v[j][u] = ld2(V + j * ldv, base + u * blockDim.x); }
         }
         #pragma unroll
         for (int j = 0; j < KB; j++)
         {
            #pragma unroll
            for (int u = 0; u < UNROLL; u++) { acc[j] += wv[u].x * v[j][u].x + wv[u].y * v[j][u].y; }
         }
      }
      // tail (fewer than one tile of double2 plus possibly one odd double): last block only
      if (blockIdx.x == gridDim.x - 1)
      {
         for (int64_t i = nfull * tile + threadIdx.x; i < n2; i += blockDim.x)
         {
            double2 wv = ld2(w, i);
            if (PC) { const double2 d = dinv ? ld2(dinv, i) : make_double2(1.0, 1.0); wv.x *= sc * d.x; wv.y *= sc * d.y; st2(w_out, i, wv); }
            #pragma unroll
            for (int j = 0; j < KB; j++) { const double2 vv = ld2(V + j * ldv, i); acc[j] += wv.x * vv.x + wv.y * vv.y; }
         }
         if ((n & 1) && threadIdx.x == 0)
         {
            double wl = w[n - 1];
            if (PC) { wl *= sc * (dinv ? dinv[n - 1] : 1.0); w_out[n - 1] = wl; }
            #pragma unroll
            for (int j = 0; j < KB; j++) { acc[j] += wl * V[j * ldv + n - 1]; }
         }
      }
   }
   else
   {
      const int64_t stride = (int64_t)gridDim.x * blockDim.x;
      for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
      {
         double wv = w[i];
         if (PC) { wv *= sc * (dinv ? dinv[i] : 1.0); w_out[i] = wv; }
         #pragma unroll
         for (int j = 0; j < KB; j++) { acc[j] += wv * V[j * ldv + i]; }
      }
   }
   #pragma unroll
   for (int j = 0; j < KB; j++)
   {
      const double s = block_sum(acc[j], red);
      if (threadIdx.x == 0) { partial[(int64_t)j * CDM_RED_BLOCKS + blockIdx.x] = s; }
   }
}

__global__ void __launch_bounds__(CDM_RED_THREADS)
k_reduce_final(const double *__restrict__ partial, double *__restrict__ out, const double *__restrict__ nrm2 = nullptr)
{
   __shared__ double red[8];
   const double *p = partial + (int64_t)blockIdx.x * CDM_RED_BLOCKS;
   double v = 0.0;
   for (int i = threadIdx.x; i < CDM_RED_BLOCKS; i += CDM_RED_THREADS) { v += p[i]; }
   const double s = block_sum(v, red);
   // nrm2: the dots were taken against un-normalised basis vectors u_i; (w, v_i) = (w, u_i) / ||u_i||
   if (threadIdx.x == 0) { out[blockIdx.x] = nrm2 ? s / sqrt(nrm2[blockIdx.x]) : s; }
}

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int KB, bool PC>
static void mdot_one(cdm_ctx *c, bool vec, int64_t n, const double *win, const double *dinv, double *w_out,
                     const double *Vj, int64_t ldv, double *pj, const double *sc2)
{
   if (vec) { k_mdot_partial<KB, (KB <= 4 ? 4 : 2), PC, true><<<CDM_RED_BLOCKS, CDM_RED_THREADS, 0, c->stream>>>(n, win, dinv, w_out, Vj, ldv, pj, sc2); }
   else { k_mdot_partial<KB, 1, PC, false><<<CDM_RED_BLOCKS, CDM_RED_THREADS, 0, c->stream>>>(n, win, dinv, w_out, Vj, ldv, pj, sc2); }
}

template <bool PC>
static int mdot_launch(cdm_ctx *c, int64_t n, int k, const double *w, const double *dinv, double *w_out,
                       const double *V, int64_t ldv, double *out_dev, const double *sc2 = nullptr, const double *nrm2 = nullptr)
{
   if (k < 1 || k > CDM_RED_MAXK) { return cdm_fail(c, CDM_EINVAL, "mdot: k out of range"); }
   double *partial = c->red_dev;
   const bool vec = aligned16(w) && aligned16(V) && (ldv % 2 == 0) && (!PC || ((!dinv || aligned16(dinv)) && aligned16(w_out)));
   for (int j0 = 0; j0 < k; j0 += 8)
   {
      const int kb = (k - j0 < 8) ? k - j0 : 8;
      const double *Vj = V + (int64_t)j0 * ldv;
      double *pj = partial + (int64_t)j0 * CDM_RED_BLOCKS;
      // the preconditioner is applied (and w_out written) by the first pass only
      const bool first = PC && j0 == 0;
      const double *win = (PC && !first) ? w_out : w;
#define MDOT_CASE(KB) case KB: if (first) { mdot_one<KB, true>(c, vec, n, win, dinv, w_out, Vj, ldv, pj, sc2); } \
                               else { mdot_one<KB, false>(c, vec, n, win, dinv, w_out, Vj, ldv, pj, nullptr); } break
      switch (kb) { MDOT_CASE(1); MDOT_CASE(2); MDOT_CASE(3); MDOT_CASE(4); MDOT_CASE(5); MDOT_CASE(6); MDOT_CASE(7); MDOT_CASE(8); }
#undef MDOT_CASE
      c->launches++;
   }
   k_reduce_final<<<k, CDM_RED_THREADS, 0, c->stream>>>(partial, out_dev, nrm2);
   VEC_CHECK(c);
}

int cdm_k_mdot_dev(cdm_ctx *c, int64_t n, int k, const double *w, const double *V, int64_t ldv, double *out_dev)
{ return mdot_launch<false>(c, n, k, w, nullptr, nullptr, V, ldv, out_dev); }

// w_out = dinv .* t, then k dots of w_out against V (one pass over t, dinv and the first 16 basis vectors)
int cdm_k_mdot_pc_dev(cdm_ctx *c, int64_t n, int k, const double *t, const double *dinv, double *w_out,
                      const double *V, int64_t ldv, double *out_dev)
{ return mdot_launch<true>(c, n, k, t, dinv, w_out, V, ldv, out_dev); }

// lazily normalised Krylov basis (V holds u_i, nrm2[i] = ||u_i||^2):
//   w_out = dinv .* t / sqrt(*sc2)  (dinv may be null), out[i] = (w_out, u_i) / sqrt(nrm2[i]),  i < k
int cdm_k_mdot_lazy_dev(cdm_ctx *c, int64_t n, int k, const double *t, const double *dinv, const double *sc2,
                        double *w_out, const double *V, int64_t ldv, const double *nrm2, double *out_dev)
{ return mdot_launch<true>(c, n, k, t, dinv, w_out, V, ldv, out_dev, sc2, nrm2); }

// y = dinv .* x / sqrt(*sc2) on a short range (ghost tail of the vector above); dinv, sc2 may be null
__global__ void k_pmult_scaled(int64_t n, const double *__restrict__ dinv, const double *__restrict__ sc2,
                               const double *__restrict__ x, double *__restrict__ y)
{
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   const double sc = sc2 ? 1.0 / sqrt(*sc2) : 1.0;
   if (i < n) { y[i] = sc * (dinv ? dinv[i] : 1.0) * x[i]; }
}
__global__ void k_scale_dots(int k, const double *__restrict__ nrm2, double *__restrict__ h)
{
   const int i = threadIdx.x;
   if (i < k) { h[i] /= sqrt(nrm2[i]); }
}
int cdm_k_scale_dots(cdm_ctx *c, int k, const double *nrm2, double *h)
{ k_scale_dots<<<1, 64, 0, c->stream>>>(k, nrm2, h); VEC_CHECK(c); }
int cdm_k_pmult_scaled(cdm_ctx *c, int64_t n, const double *dinv, const double *sc2, const double *x, double *y)
{ if (n > 0) { k_pmult_scaled<<<vec_grid(n), VEC_BLOCK, 0, c->stream>>>(n, dinv, sc2, x, y); } VEC_CHECK(c); }

// w -= sum_j h[j] V_j, and the partial sums of ||w_new||^2
template <bool VEC>
__global__ void __launch_bounds__(CDM_RED_THREADS)
k_maxpy_norm(int64_t n, int k, const double *__restrict__ h, const double *__restrict__ V, int64_t ldv,
             double *__restrict__ w, double *__restrict__ partial, const double *__restrict__ nrm2)
{
   __shared__ double sh[CDM_RED_MAXK];
   __shared__ double red[8];
   // nrm2: V holds un-normalised vectors u_i, the coefficient of u_i is h_i / ||u_i||
   if (threadIdx.x < k) { sh[threadIdx.x] = nrm2 ? h[threadIdx.x] / sqrt(nrm2[threadIdx.x]) : h[threadIdx.x]; }
   __syncthreads();
   double acc = 0.0;
   const int64_t stride = (int64_t)gridDim.x * blockDim.x;
   const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (VEC)
   {
      const int64_t n2 = n >> 1;
      for (int64_t i = gid; i < n2; i += stride)
      {
         double2 v = ld2(w, i);
         int j0 = 0;
         for (; j0 + 8 <= k; j0 += 8)                         // full groups: 8 independent loads in flight
         {
            double2 t[8];
            #pragma unroll
            for (int j = 0; j < 8; j++) { t[j] = ld2(V + (j0 + j) * ldv, i); }
            #pragma unroll
            for (int j = 0; j < 8; j++) { v.x -= sh[j0 + j] * t[j].x; v.y -= sh[j0 + j] * t[j].y; }
         }
         if (j0 + 4 <= k)
         {
            double2 t[4];
            #pragma unroll
            for (int j = 0; j < 4; j++) { t[j] = ld2(V + (j0 + j) * ldv, i); }
            #pragma unroll
            for (int j = 0; j < 4; j++) { v.x -= sh[j0 + j] * t[j].x; v.y -= sh[j0 + j] * t[j].y; }
            j0 += 4;
         }
         for (; j0 < k; j0++) { const double2 t = ld2(V + j0 * ldv, i); v.x -= sh[j0] * t.x; v.y -= sh[j0] * t.y; }
         st2(w, i, v);
         acc += v.x * v.x + v.y * v.y;
      }
      if ((n & 1) && gid == 0)
      {
         double v = w[n - 1];
         for (int j = 0; j < k; j++) { v -= sh[j] * V[j * ldv + n - 1]; }
         w[n - 1] = v;
         acc += v * v;
      }
   }
   else
   {
      for (int64_t i = gid; i < n; i += stride)
      {
         double v = w[i];
         for (int j = 0; j < k; j++) { v -= sh[j] * V[j * ldv + i]; }
         w[i] = v;
         acc += v * v;
      }
   }
   if (partial)
   {
      const double s = block_sum(acc, red);
      if (threadIdx.x == 0) { partial[blockIdx.x] = s; }
   }
}

int cdm_k_maxpy_dev(cdm_ctx *c, int64_t n, int k, const double *h_dev, const double *V, int64_t ldv,
                    double *w, double *norm2_out_dev)
{ return cdm_k_maxpy_lazy_dev(c, n, k, h_dev, nullptr, V, ldv, w, norm2_out_dev); }

// w -= sum_i (h[i] / sqrt(nrm2[i])) V_i (nrm2 may be null: plain coefficients); optional ||w_new||^2
int cdm_k_maxpy_lazy_dev(cdm_ctx *c, int64_t n, int k, const double *h_dev, const double *nrm2, const double *V, int64_t ldv,
                         double *w, double *norm2_out_dev)
{
   if (k < 0 || k > CDM_RED_MAXK) { return cdm_fail(c, CDM_EINVAL, "maxpy: k out of range"); }
   if (n <= 0) { return CDM_OK; }
   double *partial = norm2_out_dev ? c->red_dev : nullptr;
   if (aligned16(w) && aligned16(V) && (ldv % 2 == 0))
      k_maxpy_norm<true><<<CDM_RED_BLOCKS, CDM_RED_THREADS, 0, c->stream>>>(n, k, h_dev, V, ldv, w, partial, nrm2);
   else
      k_maxpy_norm<false><<<CDM_RED_BLOCKS, CDM_RED_THREADS, 0, c->stream>>>(n, k, h_dev, V, ldv, w, partial, nrm2);
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
