// kernels_apply_sub.cu -- the "register-z" warp kernel of kernels_apply_p3.cu generalised to the low
// orders, where one element does not fill a warp: EPW elements per warp, each owned by a sub-warp of
// 32/EPW lanes (order 2: D1D = 3, Q1D = 4, two elements per warp, one per half warp -- the 16 (qx,qy)
// columns of an element are exactly one 64-bit shared-memory wavefront; order 1: D1D = 2, Q1D = 3).
//
//   y_L += G^T B^T D B G x_L      (Operator::Mult inside the Krylov loop,
//                                  linear_convection_diffusion_2D.cpp:368-370; BASELINE config 3 is order 2)
//
// Same structure as k_apply3d_warp_bg: persistent grid, warp-private shared memory with __syncwarp only,
// D streamed by cp.async.bulk into a per-element ring of Q1D z-slabs guarded by mbarriers (requested one
// element ahead), x/y contractions through two exchange stages with constant-bank coefficients, z
// contraction / D product / transposed z contraction in registers, two-level gather prefetch, red.add.
// Lane roles inside a sub-warp (sl = lane % SUBW):
//   L1 (dy,dz)  sl < D^2  : x-line of nodal values (gather / scatter, x contraction)
//   L2 (qx,dz)  sl < Q D  : y contraction, qx = sl / D, dz = sl % D
//   L3 (qx,qy)  sl < Q^2  : z contraction and the quadrature-point work
// Exchange layouts (doubles): P(qx,dy,dz) = qx + PA dy + PB dz, R(qx,qy,dz) = qx + Q qy + RB dz with the
// strides of SubCfg chosen so that writer and reader lanes of a sub-warp hit distinct 8-byte banks.
#include "cdm_internal.hpp"
#include "kernels_common.cuh"

namespace
{
struct SubTables
{
   double B[CDM_MAX_Q1D * CDM_MAX_D1D];    // B[q*D + d]
   double G[CDM_MAX_Q1D * CDM_MAX_D1D];    // G[q*D + d]
};

template <int D, int Q> struct SubCfg;
// EPW elements per warp; WHOLE: one bulk copy (and one mbarrier) per element instead of one per z-slab;
// EMOD: per-element shared-memory stride modulo 16 doubles, chosen so that sub-warps that share a
// half-warp wavefront (EPW = 3: lanes 0-9 | 10-19 | 20-29) read disjoint banks in the L3 stages.
template <> struct SubCfg<2, 3> { static constexpr int EPW = 3, PA = 3, PB = 6, RB = 9, EMOD = 10; static constexpr bool WHOLE = true; };
template <> struct SubCfg<3, 4> { static constexpr int EPW = 2, PA = 5, PB = 20, RB = 20, EMOD = 0; static constexpr bool WHOLE = false; };

// mbarrier / bulk-async copy (L2 evict-first: D is read exactly once per apply) / predicated red.add helpers:
// kernels_common.cuh (namespace cdmk)
__device__ __forceinline__ void s_mbar_init(uint64_t *bar, uint32_t count) { cdmk::mbar_init(bar, count); }
__device__ __forceinline__ void s_mbar_expect_tx(uint64_t *bar, uint32_t bytes) { cdmk::mbar_expect_tx(bar, bytes); }
__device__ __forceinline__ void s_mbar_wait(uint64_t *bar, uint32_t parity) { cdmk::mbar_wait(bar, parity); }
__device__ __forceinline__ void s_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) { cdmk::bulk_g2s_stream(dst, src, bytes, bar); }
// y[g] += v unless g < 0 (essential dof / element owned by nobody): predicated, no branch
__device__ __forceinline__ void s_red_add_f64_if(double *y, int g, double v) { cdmk::red_add_f64_pred(y, g, v); }

template <int D, int Q>
__host__ __device__ constexpr int sub_elem_doubles(int slab)
{
   // ring of Q slabs + three R arrays + two P arrays, kept 16-byte aligned for the bulk copies
   // (even number of doubles) and congruent to EMOD modulo 16
   int n = (Q * slab + 3 * SubCfg<D, Q>::RB * D + 2 * SubCfg<D, Q>::PB * D + 1) & ~1;
   while ((n & 15) != SubCfg<D, Q>::EMOD) { n += 2; }
   return n;
}

// launch shape (measured on B200, 8 M dofs, % of HBM roofline at p=2 / p=1): 2 warps 91.7 / 89.3,
// 3 warps x 5 CTAs 92.7 / 87.7, 4 warps x 4 CTAs 94.6 / 90.5
#ifndef CDM_SUB_NW
#define CDM_SUB_NW 4
#endif
#ifndef CDM_SUB_MINB
#define CDM_SUB_MINB 4
#endif

template <int D, int Q, int NW, bool DIFF, bool CONV, bool MASS, bool ATOMIC>
__global__ void __launch_bounds__(NW * 32, CDM_SUB_MINB)
k_apply3d_sub(const SubTables tb, const int64_t ne, const int32_t *__restrict__ gmap,
              const double *__restrict__ x, const double *__restrict__ Dg, const int slab,
              double *__restrict__ y)
{
   using C = SubCfg<D, Q>;
   constexpr int EPW = C::EPW, SUBW = 32 / EPW, Q2 = Q * Q, ND = D * D * D;
   constexpr int PA = C::PA, PB = C::PB, RB = C::RB, PS = PB * D, RS = RB * D;
   constexpr bool GRAD = DIFF || CONV;
   static_assert(Q2 <= SUBW && Q * D <= SUBW, "an element must fit its sub-warp");
   extern __shared__ __align__(128) unsigned char smraw[];
   const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
   const int sub = (lane / SUBW < EPW) ? lane / SUBW : EPW - 1;        // spare lanes (EPW = 3: 30, 31) take no role
   const int sl = lane - sub * SUBW;
   constexpr bool WHOLE = C::WHOLE;
   constexpr int NB = WHOLE ? 1 : Q;                                    // mbarriers per warp
   const int elem_doubles = sub_elem_doubles<D, Q>(slab);
   const int warp_doubles = (EPW * elem_doubles + 15) & ~15;
   double *wring = reinterpret_cast<double *>(smraw) + wib * warp_doubles;
   double *ring = wring + sub * elem_doubles;
   uint64_t *wbars = reinterpret_cast<uint64_t *>(reinterpret_cast<double *>(smraw) + NW * warp_doubles) + wib * NB;
   double *sR0 = ring + Q * slab, *sR1 = sR0 + RS, *sR2 = sR1 + RS, *sP0 = sR2 + RS, *sP1 = sP0 + PS;
   const bool l1 = sl < D * D, l2 = sl < Q * D, l3 = sl < Q2;
   const int l3i = l3 ? sl : 0;                                          // L3 role: (qx,qy) = sl
   const int qx2 = l2 ? (sl / D) : 0, dz2 = sl % D;                      // L2 role

   if (lane == 0)
   {
      for (int b = 0; b < NB; b++) { s_mbar_init(&wbars[b], 1); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
   }
   __syncwarp();
   const int64_t gw = (int64_t)blockIdx.x * NW + wib, tw = (int64_t)gridDim.x * NW;
   const int64_t estride = tw * EPW;
   const uint32_t copy_bytes = (uint32_t)slab * 8u * (WHOLE ? Q : 1);
   // lane 0 requests the D tiles of the EPW elements of one warp-iteration (first element e0): per
   // barrier one expect_tx for all EPW copies, then the copies
   auto request = [&](int64_t e0)
   {
      #pragma unroll
      for (int b = 0; b < NB; b++)
      {
         s_mbar_expect_tx(&wbars[b], copy_bytes * EPW);
         #pragma unroll
         for (int s = 0; s < EPW; s++)
         {
            const int64_t es = (e0 + s < ne) ? e0 + s : ne - 1;
            s_bulk_g2s(wring + s * elem_doubles + b * slab, Dg + (es * Q + b) * (int64_t)slab, copy_bytes, &wbars[b]);
         }
      }
   };
   // Control flow is kept warp-uniform: when the last warp-iteration holds fewer than EPW elements the
   // spare sub-warp re-processes the last element with its output disabled (divergent sub-warps would
   // issue every instruction of the element once per sub-warp).
   const int64_t last = ne - 1;
   // gather pipeline: indices two elements ahead, values one element ahead
   int pg[D], pgn[D];
   double px[D];
   #pragma unroll
   for (int d = 0; d < D; d++) { pg[d] = -1; pgn[d] = -1; px[d] = 0.0; }
   if (gw * EPW < ne)
   {
      const int64_t e = (gw * EPW + sub < ne) ? gw * EPW + sub : last;
      if (lane == 0) { request(gw * EPW); }
      if (l1)
      {
         #pragma unroll
         for (int d = 0; d < D; d++) { pg[d] = __ldg(gmap + e * ND + D * sl + d); }
         if (gw * EPW + estride < ne)
         {
            const int64_t e1 = (gw * EPW + estride + sub < ne) ? gw * EPW + estride + sub : last;
            #pragma unroll
            for (int d = 0; d < D; d++) { pgn[d] = __ldg(gmap + e1 * ND + D * sl + d); }
         }
         #pragma unroll
         for (int d = 0; d < D; d++) { px[d] = (pg[d] >= 0) ? __ldg(x + pg[d]) : 0.0; }
      }
   }

   uint32_t parity = 0;
   for (int64_t eb = gw * EPW; eb < ne; eb += estride, parity ^= 1u)
   {
      const bool act = eb + sub < ne;                        // false: duplicate of the last element, output disabled
      const int64_t e = act ? eb + sub : last;
      const bool more = eb + estride < ne;                   // warp-uniform
      int g[D];
      #pragma unroll
      for (int d = 0; d < D; d++) { g[d] = act ? pg[d] : -1; }
      // ---- F1 (L1 lanes): x contraction of the own x-line with B and G
      if (l1)
      {
         #pragma unroll
         for (int q = 0; q < Q; q++)
         {
            double tB = 0.0, tG = 0.0;
            #pragma unroll
            for (int d = 0; d < D; d++) { tB += tb.B[q * D + d] * px[d]; if (GRAD) { tG += tb.G[q * D + d] * px[d]; } }
            const int dy = sl % D, dz = sl / D;
            sP0[q + PA * dy + PB * dz] = tB;
            if (GRAD) { sP1[q + PA * dy + PB * dz] = tG; }
         }
      }
      if (more && l1)
      {
         #pragma unroll
         for (int d = 0; d < D; d++) { pg[d] = pgn[d]; }
         if (eb + 2 * estride < ne)
         {
            const int64_t e2 = (eb + 2 * estride + sub < ne) ? eb + 2 * estride + sub : last;
            #pragma unroll
            for (int d = 0; d < D; d++) { pgn[d] = __ldg(gmap + e2 * ND + D * sl + d); }
         }
         #pragma unroll
         for (int d = 0; d < D; d++) { px[d] = (pg[d] >= 0) ? __ldg(x + pg[d]) : 0.0; }
      }
      __syncwarp();
      // ---- F2 (L2 lanes): y contraction -> (B B), (G B), (B G)
      if (l2)
      {
         double tB[D], tG[D];
         #pragma unroll
         for (int dy = 0; dy < D; dy++)
         {
            tB[dy] = sP0[qx2 + PA * dy + PB * dz2];
            if (GRAD) { tG[dy] = sP1[qx2 + PA * dy + PB * dz2]; }
         }
         #pragma unroll
         for (int q = 0; q < Q; q++)
         {
            double vbb = 0.0, vgb = 0.0, vbg = 0.0;
            #pragma unroll
            for (int dy = 0; dy < D; dy++)
            {
               vbb += tb.B[q * D + dy] * tB[dy];
               if (GRAD) { vgb += tb.B[q * D + dy] * tG[dy]; vbg += tb.G[q * D + dy] * tB[dy]; }
            }
            sR0[qx2 + Q * q + RB * dz2] = vbb;
            if (GRAD) { sR1[qx2 + Q * q + RB * dz2] = vgb; sR2[qx2 + Q * q + RB * dz2] = vbg; }
         }
      }
      __syncwarp();
      // ---- F3 (L3 lanes): z contraction in registers: u, ux, uy, uz on the lane's z-column
      double u[Q], ux[Q], uy[Q], uz[Q];
      {
         double vbb[D], vgb[D], vbg[D];
         #pragma unroll
         for (int dz = 0; dz < D; dz++)
         {
            vbb[dz] = sR0[l3i + RB * dz];
            if (GRAD) { vgb[dz] = sR1[l3i + RB * dz]; vbg[dz] = sR2[l3i + RB * dz]; }
         }
         #pragma unroll
         for (int qz = 0; qz < Q; qz++)
         {
            double a = 0.0, b = 0.0, c = 0.0, d = 0.0;
            #pragma unroll
            for (int dz = 0; dz < D; dz++)
            {
               a += tb.B[qz * D + dz] * vbb[dz];
               if (GRAD) { b += tb.B[qz * D + dz] * vgb[dz]; c += tb.B[qz * D + dz] * vbg[dz]; d += tb.G[qz * D + dz] * vbb[dz]; }
            }
            u[qz] = a; ux[qz] = b; uy[qz] = c; uz[qz] = d;
         }
      }
      // ---- point-wise D at the lane's quadrature points (registers only)
      #pragma unroll
      for (int qz = 0; qz < Q; qz++)
      {
         if (!WHOLE || qz == 0) { s_mbar_wait(&wbars[WHOLE ? 0 : qz], parity); }   // warp-uniform wait
         const double *dp = ring + qz * slab + l3i;
         double fx = 0.0, fy = 0.0, fz = 0.0, s = 0.0;
         int c = 0;
         if (DIFF)
         {
            const double d0 = dp[0], d1 = dp[Q2], d2 = dp[2 * Q2], d3 = dp[3 * Q2], d4 = dp[4 * Q2], d5 = dp[5 * Q2];
            fx = d0 * ux[qz] + d1 * uy[qz] + d2 * uz[qz];
            fy = d1 * ux[qz] + d3 * uy[qz] + d4 * uz[qz];
            fz = d2 * ux[qz] + d4 * uy[qz] + d5 * uz[qz];
            c = 6;
         }
         if (CONV) { s = dp[c * Q2] * ux[qz] + dp[(c + 1) * Q2] * uy[qz] + dp[(c + 2) * Q2] * uz[qz]; c += 3; }
         if (MASS) { s += dp[c * Q2] * u[qz]; }
         ux[qz] = fx; uy[qz] = fy; uz[qz] = fz; u[qz] = s;
      }
      __syncwarp();                                          // D tile and the R buffers are consumed by every lane
      if (lane == 0 && more) { request(eb + estride); }
      // ---- B1 (L3 lanes): transposed z contraction in registers
      if (l3)
      {
         #pragma unroll
         for (int dz = 0; dz < D; dz++)
         {
            double wx = 0.0, wy = 0.0, wb = 0.0;
            #pragma unroll
            for (int qz = 0; qz < Q; qz++)
            {
               wb += tb.B[qz * D + dz] * u[qz];
               if (DIFF) { wx += tb.B[qz * D + dz] * ux[qz]; wy += tb.B[qz * D + dz] * uy[qz]; wb += tb.G[qz * D + dz] * uz[qz]; }
            }
            sR2[sl + RB * dz] = wb;
            if (DIFF) { sR0[sl + RB * dz] = wx; sR1[sl + RB * dz] = wy; }
         }
      }
      __syncwarp();
      // ---- B2 (L2 lanes): transposed y contraction
      if (l2)
      {
         double wx[Q], wy[Q], wb[Q];
         #pragma unroll
         for (int q = 0; q < Q; q++)
         {
            wb[q] = sR2[qx2 + Q * q + RB * dz2];
            if (DIFF) { wx[q] = sR0[qx2 + Q * q + RB * dz2]; wy[q] = sR1[qx2 + Q * q + RB * dz2]; }
         }
         #pragma unroll
         for (int dy = 0; dy < D; dy++)
         {
            double a1 = 0.0, a2 = 0.0;
            #pragma unroll
            for (int q = 0; q < Q; q++)
            {
               a2 += tb.B[q * D + dy] * wb[q];
               if (DIFF) { a1 += tb.B[q * D + dy] * wx[q]; a2 += tb.G[q * D + dy] * wy[q]; }
            }
            sP1[qx2 + PA * dy + PB * dz2] = a2;
            if (DIFF) { sP0[qx2 + PA * dy + PB * dz2] = a1; }
         }
      }
      __syncwarp();
      // ---- B3 (L1 lanes): transposed x contraction of the own x-line, scatter
      if (l1)
      {
         double a1[Q], a2[Q];
         const int dy = sl % D, dz = sl / D;
         #pragma unroll
         for (int q = 0; q < Q; q++) { a2[q] = sP1[q + PA * dy + PB * dz]; if (DIFF) { a1[q] = sP0[q + PA * dy + PB * dz]; } }
         double yv[D];                                       // results first, scatter after: the asm red.add is a compiler barrier
         #pragma unroll
         for (int dx = 0; dx < D; dx++)
         {
            double a = 0.0;
            #pragma unroll
            for (int q = 0; q < Q; q++) { a += tb.B[q * D + dx] * a2[q]; if (DIFF) { a += tb.G[q * D + dx] * a1[q]; } }
            yv[dx] = a;
         }
         #pragma unroll
         for (int dx = 0; dx < D; dx++)
         {
            if (ATOMIC) { s_red_add_f64_if(y, g[dx], yv[dx]); }
            else if (act) { y[e * ND + D * sl + dx] = yv[dx]; }
         }
      }
      __syncwarp();                                          // P buffers are rewritten by the next element's F1
   }
}

template <int D, int Q, int NW, bool DIFF, bool CONV, bool MASS, bool ATOMIC>
int launch_sub(cdm_op *op, const SubTables &tb, const int32_t *gmap, const double *xL, double *out)
{
   cdm_space *sp = op->sp;
   cdm_ctx *ctx = sp->ctx;
   constexpr int EPW = SubCfg<D, Q>::EPW, ND = D * D * D;
   auto kern = k_apply3d_sub<D, Q, NW, DIFF, CONV, MASS, ATOMIC>;
   const int warp_doubles = (EPW * sub_elem_doubles<D, Q>(op->slab) + 15) & ~15;
   const size_t smem = (size_t)(NW * warp_doubles) * sizeof(double) + (size_t)NW * Q * sizeof(uint64_t);
   int blocks_per_sm = 0;
   { const int rc = cdm_kernel_cfg(ctx, (const void *)kern, NW * 32, smem, "k_apply3d_sub", &blocks_per_sm); if (rc) { return rc; } }
   const int64_t e0 = op->range_on ? op->e_begin : 0, e1 = op->range_on ? op->e_end : sp->ne;
   const int64_t n = e1 - e0;
   if (n <= 0) { return CDM_OK; }
   int64_t grid = (int64_t)ctx->sm_count * blocks_per_sm;
   const int64_t need = (n + NW * EPW - 1) / (NW * EPW);
   if (grid > need) { grid = need; }
   if (op->grid_cap > 0 && grid > op->grid_cap) { grid = op->grid_cap; }
   if (ctx->time_main) { cudaEventRecord(ctx->evk0, ctx->stream); }
   kern<<<(unsigned)grid, NW * 32, smem, ctx->stream>>>(tb, n, gmap + e0 * ND, xL, op->D_dev + e0 * Q * (int64_t)op->slab,
                                                        op->slab, ATOMIC ? out : out + e0 * ND);
   if (ctx->time_main) { cudaEventRecord(ctx->evk1, ctx->stream); }
   ctx->launches++;
   CDM_CUDA(ctx, cudaGetLastError());
   return CDM_OK;
}

__global__ void __launch_bounds__(256)
k_restrict_transpose_sub(int64_t ndof, const int32_t *__restrict__ offsets, const int32_t *__restrict__ indices,
                         const double *__restrict__ yE, double *__restrict__ y)
{
   const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (g >= ndof) { return; }
   double s = 0.0;
   for (int32_t j = offsets[g]; j < offsets[g + 1]; j++) { s += yE[indices[j]]; }
   y[g] = s;
}
}  // namespace

#define SUB_ONE(D_, Q_, DF, CV, MS)                                                            \
   (atomic ? launch_sub<D_, Q_, CDM_SUB_NW, DF, CV, MS, true>(op, tb, gmap, xL, out)           \
           : launch_sub<D_, Q_, CDM_SUB_NW, DF, CV, MS, false>(op, tb, gmap, xL, out))
#define SUB_FLAGS(D_, Q_)                                                                      \
   (op->has_diff && op->has_conv && op->has_mass) ? SUB_ONE(D_, Q_, true, true, true)          \
   : (op->has_diff && !op->has_conv && op->has_mass) ? SUB_ONE(D_, Q_, true, false, true)      \
   : (op->has_diff && !op->has_conv && !op->has_mass) ? SUB_ONE(D_, Q_, true, false, false)    \
   : (!op->has_diff && !op->has_conv && op->has_mass) ? SUB_ONE(D_, Q_, false, false, true)    \
   : 1

// returns 1 when the (order, integrator set) is not covered: the caller falls back to the block kernel
int cdm_k_apply_sub(cdm_op *op, const int32_t *gmap, const double *xL, double *yL)
{
   cdm_space *sp = op->sp;
   cdm_ctx *ctx = sp->ctx;
   if (sp->dim != 3 || (sp->p != 1 && sp->p != 2)) { return 1; }
   SubTables tb;
   memset(&tb, 0, sizeof(tb));
   for (int i = 0; i < sp->q1d * sp->d1d; i++) { tb.B[i] = sp->B[i]; tb.G[i] = sp->G[i]; }
   const bool atomic = op->scatter_mode == 1;
   double *out = yL;
   if (!atomic)
   {
      if (!op->yE_dev) { CDM_CUDA(ctx, cudaMalloc(&op->yE_dev, sizeof(double) * (size_t)sp->ne * sp->nd)); }
      out = op->e_out ? op->e_out : op->yE_dev;
   }
   // red.add accumulates: the caller of a ranged launch (overlapped multi-GPU schedule) zeroes y itself
   else if (!op->range_on) { CDM_CUDA(ctx, cudaMemsetAsync(yL, 0, sizeof(double) * (size_t)sp->ndof, ctx->stream)); }
   int rc = (sp->p == 2) ? (SUB_FLAGS(3, 4)) : (SUB_FLAGS(2, 3));
   if (rc) { return rc; }
   if (!atomic && !op->e_out)
   {
      const int64_t nb = (sp->ndof + 255) / 256;
      k_restrict_transpose_sub<<<(unsigned)nb, 256, 0, ctx->stream>>>(sp->ndof, sp->offsets_dev, sp->indices_dev, op->yE_dev, yL);
      ctx->launches++;
      CDM_CUDA(ctx, cudaGetLastError());
   }
   return CDM_OK;
}
