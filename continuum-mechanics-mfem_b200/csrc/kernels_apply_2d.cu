// kernels_apply_2d.cu -- 2D quadrilateral apply, one THREAD per element (kernel option 6, default in 2D for p <= 4).
//
// BASELINE config 1 is the reference as shipped: 2D quads, order 2 (Input/input_2d.yaml, linear_convection_diffusion_2D.cpp
// :311-339).  A 2D element is tiny -- (p+1)^2 dofs, (p+1)^2 quadrature points (rule order 2p+1), 6 (p+1)^2 doubles of
// quadrature data: 432 B at p = 2 -- so a block of threads per element (k_apply2d_generic) spends its time in barriers and
// L1 wavefronts (ncu r02: L1TEX 95 %, DRAM 30 %, 33-55 % of the HBM roofline).  Here a warp owns 32 consecutive elements:
//   * the quadrature data of the 32 elements is one contiguous 32 x slab block of global memory; every lane issues ONE
//     cp.async.bulk for its element's slab into its own shared-memory row, all completing on one mbarrier; rows are padded
//     to a stride = 2 (mod 4) doubles so that 128-bit LDS of a quarter warp hit 8 different bank quads (conflict-free);
//     double-buffered (or single-buffered with the refill issued as soon as the data has been consumed) and evict-first in
//     L2 so the stream does not displace x / y;
//   * everything else lives in the thread's registers: gather of the (p+1)^2 dofs, x and y contractions with the 1-D
//     tables from the constant bank, the point-wise D, the transposed contractions fused row by row, fp64 red.add scatter
//     (or E-vector stores for the deterministic mode).  No barriers except one __syncwarp per chunk.
// Algorithmic bytes per element: 8 n_D Q^2 + 4 D^2 (+ 16 B per unique dof), SURVEY.md 8(d).
#include "cdm_internal.hpp"
#include "kernels_common.cuh"
#include <type_traits>

namespace
{
using namespace cdmk;
#ifndef CDM_2D_IDX32_FROM
#define CDM_2D_IDX32_FROM 3
#endif

template <int P, int NW, int NBUF, int MINB, bool DIFF, bool CONV, bool MASS, bool ATOMIC>
__global__ void __launch_bounds__(NW * 32, MINB)
k_apply2d_thread(const BasisTables tb, const int64_t ne, const int32_t *__restrict__ gmap, const double *__restrict__ x,
                 const double *__restrict__ Dg, const int slab, const int sstride, double *__restrict__ y)
{
   constexpr int D = P + 1, Q = P + 1, ND = D * D, Q2 = Q * Q;
   // chunk / element indices: 32-bit at the orders that sit at the register cap (the launcher refuses more than 2e9 elements)
   using eidx = typename std::conditional<(P >= CDM_2D_IDX32_FROM), int, int64_t>::type;
   constexpr int OC = DIFF ? 3 : 0, OM = OC + (CONV ? 2 : 0);
   constexpr bool PF = P <= 3;              // gather one chunk ahead (order 4 has no registers to spare for it)
   // ... and the indices two chunks ahead, so that the value loads never wait for them: measured +2.5 % at p = 1, +3 % at
   // p = 3, -1 % at p = 2 (profiles/r02_sweep2d_variants.md)
   constexpr bool PF2 = P == 1 || P == 3;
   // (order 4, tried: not keeping the 25 dof indices across the middle stage but re-reading them before the scatter halves the
   // spills and costs 14 points: 50.5 -> 36.9 %)
   extern __shared__ __align__(128) unsigned char smraw[];
   const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
   double *wbuf = reinterpret_cast<double *>(smraw) + (size_t)wib * NBUF * 32 * sstride;
   uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<double *>(smraw) + (size_t)NW * NBUF * 32 * sstride) + wib * NBUF;
   if (lane == 0)
   {
      for (int b = 0; b < NBUF; b++) { mbar_init(&bars[b], 1); }
      mbar_init_fence();
   }
   __syncwarp();
   const eidx nwarps = (eidx)gridDim.x * NW, gw = (eidx)blockIdx.x * NW + wib;
   const eidx nel = (eidx)ne, nchunks = (eidx)((ne + 31) >> 5);
   const uint32_t slab_bytes = (uint32_t)slab * 8u;
   auto issue = [&](eidx chunk, int b)
   {
      const eidx e = chunk * 32 + lane;
      const eidx left = nel - chunk * 32;
      if (lane == 0) { mbar_expect_tx(&bars[b], (uint32_t)(left < 32 ? left : 32) * slab_bytes); }
      __syncwarp();
      if (e < nel) { bulk_g2s_stream(wbuf + (size_t)(b * 32 + lane) * sstride, Dg + e * (int64_t)slab, slab_bytes, &bars[b]); }
   };
   if (gw < nchunks) { issue(gw, 0); }
   // gather pipeline: the dofs of the next chunk are fetched while the current one is being processed
   int32_t g[ND], gn[ND], gnn[PF2 ? ND : 1];
   double u[ND];
   {
      const eidx e = gw * 32 + lane;
      #pragma unroll
      for (int i = 0; i < ND; i++) { gn[i] = (PF && gw < nchunks && e < nel) ? __ldg(gmap + (int64_t)e * ND + i) : -1; }
      if (PF2)
      {
         const eidx e2 = (gw + nwarps) * 32 + lane;
         #pragma unroll
         for (int i = 0; i < ND; i++) { gnn[i] = (gw + nwarps < nchunks && e2 < nel) ? __ldg(gmap + (int64_t)e2 * ND + i) : -1; }
      }
      #pragma unroll
      for (int i = 0; i < ND; i++) { u[i] = (PF && gn[i] >= 0) ? __ldg(x + gn[i]) : 0.0; }
   }
   int it = 0;
   for (eidx chunk = gw; chunk < nchunks; chunk += nwarps, it++)
   {
      const int b = (NBUF == 2) ? (it & 1) : 0;
      const eidx next = chunk + nwarps;
      if (NBUF == 2 && next < nchunks) { issue(next, b ^ 1); }     // the other buffer was consumed one iteration ago
      const eidx e = chunk * 32 + lane;
      const bool valid = e < nel;
      if (PF)
      {
         #pragma unroll
         for (int i = 0; i < ND; i++) { g[i] = gn[i]; }
      }
      else
      {
         #pragma unroll
         for (int i = 0; i < ND; i++) { g[i] = valid ? __ldg(gmap + (int64_t)e * ND + i) : -1; }
         #pragma unroll
         for (int i = 0; i < ND; i++) { u[i] = (g[i] >= 0) ? __ldg(x + g[i]) : 0.0; }
      }
      // ---- x contraction: bu[dy][qx] = sum_dx B[qx][dx] u[dy][dx], gu with G
      double bu[D][Q], gu[D][Q];
      #pragma unroll
      for (int dy = 0; dy < D; dy++)
      {
         #pragma unroll
         for (int qx = 0; qx < Q; qx++)
         {
            double s0 = 0.0, s1 = 0.0;
            #pragma unroll
            for (int dx = 0; dx < D; dx++)
            {
               s0 += tb.B[qx * D + dx] * u[dy * D + dx];
               if (DIFF || CONV) { s1 += tb.G[qx * D + dx] * u[dy * D + dx]; }
            }
            bu[dy][qx] = s0; gu[dy][qx] = s1;
         }
      }
      // ---- u is dead: fetch the dofs of the next chunk (used one iteration from now)
      if (PF2)
      {
         #pragma unroll
         for (int i = 0; i < ND; i++) { gn[i] = gnn[i]; }
         #pragma unroll
         for (int i = 0; i < ND; i++) { u[i] = (gn[i] >= 0) ? __ldg(x + gn[i]) : 0.0; }
         const eidx n2 = next + nwarps, e2 = n2 * 32 + lane;
         const bool nv = n2 < nchunks && e2 < nel;
         #pragma unroll
         for (int i = 0; i < ND; i++) { gnn[i] = nv ? __ldg(gmap + (int64_t)e2 * ND + i) : -1; }
      }
      else if (PF)
      {
         const eidx en = next * 32 + lane;
         const bool nv = next < nchunks && en < nel;
         #pragma unroll
         for (int i = 0; i < ND; i++) { gn[i] = nv ? __ldg(gmap + (int64_t)en * ND + i) : -1; }
         #pragma unroll
         for (int i = 0; i < ND; i++) { u[i] = (gn[i] >= 0) ? __ldg(x + gn[i]) : 0.0; }
      }
      // ---- quadrature data of this chunk
      mbar_wait(&bars[b], (uint32_t)((NBUF == 2 ? (it >> 1) : it) & 1));
      const double2 *dp = reinterpret_cast<const double2 *>(wbuf + (size_t)(b * 32 + lane) * sstride);
      auto dv = [&](int k) -> double { const double2 t = dp[k >> 1]; return (k & 1) ? t.y : t.x; };
      // ---- y contraction, point-wise D, transposed y contraction, one row of points at a time
      double a[D][Q], c[D][Q];                                     // [dy][qx]: a pairs with G^T in x, c with B^T in x
      #pragma unroll
      for (int dy = 0; dy < D; dy++)
      {
         #pragma unroll
         for (int qx = 0; qx < Q; qx++) { a[dy][qx] = 0.0; c[dy][qx] = 0.0; }
      }
      #pragma unroll
      for (int qy = 0; qy < Q; qy++)
      {
         #pragma unroll
         for (int qx = 0; qx < Q; qx++)
         {
            double U = 0.0, Ux = 0.0, Uy = 0.0;
            #pragma unroll
            for (int dy = 0; dy < D; dy++)
            {
               U += tb.B[qy * D + dy] * bu[dy][qx];
               if (DIFF || CONV) { Ux += tb.B[qy * D + dy] * gu[dy][qx]; Uy += tb.G[qy * D + dy] * bu[dy][qx]; }
            }
            const int q = qx + Q * qy;
            double fx = 0.0, fy = 0.0, s = 0.0;
            if (DIFF)
            {
               const double d0 = dv(q), d1 = dv(Q2 + q), d2 = dv(2 * Q2 + q);
               fx = d0 * Ux + d1 * Uy; fy = d1 * Ux + d2 * Uy;
            }
            if (CONV) { s = dv(OC * Q2 + q) * Ux + dv((OC + 1) * Q2 + q) * Uy; }
            if (MASS) { s += dv(OM * Q2 + q) * U; }
            #pragma unroll
            for (int dy = 0; dy < D; dy++)
            {
               c[dy][qx] += tb.B[qy * D + dy] * s;
               if (DIFF) { a[dy][qx] += tb.B[qy * D + dy] * fx; c[dy][qx] += tb.G[qy * D + dy] * fy; }
            }
         }
      }
      __syncwarp();                                                // every lane has consumed its row of the buffer
      if (NBUF == 1 && next < nchunks) { issue(next, 0); }
      // ---- transposed x contraction and scatter
      #pragma unroll
      for (int dy = 0; dy < D; dy++)
      {
         #pragma unroll
         for (int dx = 0; dx < D; dx++)
         {
            double v = 0.0;
            #pragma unroll
            for (int qx = 0; qx < Q; qx++)
            {
               v += tb.B[qx * D + dx] * c[dy][qx];
               if (DIFF) { v += tb.G[qx * D + dx] * a[dy][qx]; }
            }
            if (ATOMIC) { if (g[dy * D + dx] >= 0) { red_add_f64(y + g[dy * D + dx], v); } }
            else if (valid) { y[(int64_t)e * ND + dy * D + dx] = v; }
         }
      }
   }
}

// rows of the staging buffer: >= slab doubles, = 2 (mod 4) so 128-bit shared loads of a quarter warp are conflict-free
int staging_stride(int slab) { int s = slab; while ((s & 3) != 2) { s++; } return s; }

template <int P, int NW, int NBUF, int MINB, bool DIFF, bool CONV, bool MASS, bool ATOMIC>
int launch_2d(cdm_op *op, const BasisTables &tb, const int32_t *gmap, const double *xL, double *out)
{
   cdm_space *sp = op->sp;
   cdm_ctx *ctx = sp->ctx;
   constexpr int ND = (P + 1) * (P + 1);
   auto kern = k_apply2d_thread<P, NW, NBUF, MINB, DIFF, CONV, MASS, ATOMIC>;
   const int sstride = staging_stride(op->slab);
   const size_t smem = (size_t)NW * NBUF * 32 * sstride * sizeof(double) + (size_t)NW * NBUF * sizeof(uint64_t);
   int blocks_per_sm = 0;
   { const int rc = cdm_kernel_cfg(ctx, (const void *)kern, NW * 32, smem, "k_apply2d_thread", &blocks_per_sm); if (rc) { return rc; } }
   const int64_t e0 = op->range_on ? op->e_begin : 0, e1 = op->range_on ? op->e_end : sp->ne;
   const int64_t n = e1 - e0;
   if (n <= 0) { return CDM_OK; }
   if (n > 2000000000LL) { return 1; }                        // 32-bit element indices at orders >= 3: the caller falls back
   int64_t grid = (int64_t)ctx->sm_count * blocks_per_sm;
   const int64_t need = (n + 32 * NW - 1) / (32 * NW);
   if (grid > need) { grid = need; }
   if (op->grid_cap > 0 && grid > op->grid_cap) { grid = op->grid_cap; }
   if (ctx->time_main) { cudaEventRecord(ctx->evk0, ctx->stream); }
   kern<<<(unsigned)grid, NW * 32, smem, ctx->stream>>>(tb, n, gmap + e0 * ND, xL, op->D_dev + e0 * (int64_t)op->slab, op->slab, sstride,
                                                         ATOMIC ? out : out + e0 * ND);
   if (ctx->time_main) { cudaEventRecord(ctx->evk1, ctx->stream); }
   ctx->launches++;
   CDM_CUDA(ctx, cudaGetLastError());
   return CDM_OK;
}

template <int P, int NW, int NBUF, int MINB>
int dispatch_2d(cdm_op *op, const BasisTables &tb, const int32_t *gmap, const double *xL, double *out, bool atomic)
{
#define T2D_ONE(DF, CV, MS) (atomic ? launch_2d<P, NW, NBUF, MINB, DF, CV, MS, true>(op, tb, gmap, xL, out) \
                                    : launch_2d<P, NW, NBUF, MINB, DF, CV, MS, false>(op, tb, gmap, xL, out))
   if (op->has_diff && op->has_conv && op->has_mass) { return T2D_ONE(true, true, true); }
   if (op->has_diff && !op->has_conv && op->has_mass) { return T2D_ONE(true, false, true); }
   if (!op->has_diff && !op->has_conv && op->has_mass) { return T2D_ONE(false, false, true); }
   if (op->has_diff && !op->has_conv && !op->has_mass) { return T2D_ONE(true, false, false); }
#undef T2D_ONE
   return 1;
}
}  // namespace

// per order: warps per block (NW), staging buffers per warp (NB), resident blocks the register budget is sized for (MB).
// Measured at 8 M dofs (profiles/r02_sweep2d_variants.md): p=2 (4,2,2) 90.6 % / (4,1,3) 87.7 % / (4,1,4) 81.1 %;
// p=3 (4,2,1) 70.2 % / (4,1,2) 74.5 % / (8,1,1) 75.1 %; p=4 (4,1,1) 49.8 % / (2,1,2) 50.4 % / (2,2,1) 30.2 %.
#ifndef CDM_2D_P1_NW
#define CDM_2D_P1_NW 4
#define CDM_2D_P1_NB 2
#define CDM_2D_P1_MB 4
#endif
#ifndef CDM_2D_P2_NW
#define CDM_2D_P2_NW 4
#define CDM_2D_P2_NB 2
#define CDM_2D_P2_MB 2
#endif
#ifndef CDM_2D_P3_NW
#define CDM_2D_P3_NW 4
#define CDM_2D_P3_NB 1
#define CDM_2D_P3_MB 2
#endif
#ifndef CDM_2D_P4_NW
#define CDM_2D_P4_NW 4
#define CDM_2D_P4_NB 1
#define CDM_2D_P4_MB 1
#endif

// returns 1 when this operator is not covered (caller falls back to the generic kernel)
int cdm_k_apply_2d_thread(cdm_op *op, const int32_t *gmap, const double *xL, double *yL)
{
   cdm_space *sp = op->sp;
   cdm_ctx *ctx = sp->ctx;
   if (sp->dim != 2 || sp->geom != 0 || sp->p < 1 || sp->p > 4) { return 1; }
   if (op->has_conv && !(op->has_diff && op->has_mass)) { return 1; }
   if (!op->has_diff && !op->has_mass) { return 1; }
   BasisTables tb;
   memset(&tb, 0, sizeof(tb));
   for (int i = 0; i < sp->q1d * sp->d1d; i++) { tb.B[i] = sp->B[i]; tb.G[i] = sp->G[i]; }
   const bool atomic = op->scatter_mode == 1;
   double *out = yL;
   if (atomic) { if (!op->range_on) { CDM_CUDA(ctx, cudaMemsetAsync(yL, 0, sizeof(double) * (size_t)sp->ndof, ctx->stream)); } }
   else
   {
      if (!op->yE_dev) { CDM_CUDA(ctx, cudaMalloc(&op->yE_dev, sizeof(double) * (size_t)sp->ne * sp->nd)); }
      out = op->e_out ? op->e_out : op->yE_dev;
   }
   int rc = 1;
   switch (sp->p)
   {
      case 1: rc = dispatch_2d<1, CDM_2D_P1_NW, CDM_2D_P1_NB, CDM_2D_P1_MB>(op, tb, gmap, xL, out, atomic); break;
      case 2: rc = dispatch_2d<2, CDM_2D_P2_NW, CDM_2D_P2_NB, CDM_2D_P2_MB>(op, tb, gmap, xL, out, atomic); break;
      case 3: rc = dispatch_2d<3, CDM_2D_P3_NW, CDM_2D_P3_NB, CDM_2D_P3_MB>(op, tb, gmap, xL, out, atomic); break;
      case 4: rc = dispatch_2d<4, CDM_2D_P4_NW, CDM_2D_P4_NB, CDM_2D_P4_MB>(op, tb, gmap, xL, out, atomic); break;
   }
   if (rc) { return rc; }
   if (!atomic && !op->e_out) { return cdm_k_restrict_transpose(sp, op->yE_dev, yL); }
   return CDM_OK;
}
