// kernels_apply_p3.cu -- the headline kernel: fused 3D apply for orders whose
// quadrature slab fits one warp (Q1D^2 <= 32: p = 1, 2, 3), written for sm_100a.
//
//   y_L += G^T B^T D B G x_L      (Operator::Mult inside the Krylov loop,
//                                  linear_convection_diffusion_2D.cpp:368-370)
//
// Design (see DESIGN.md "apply kernel"):
//   * persistent grid, one warp owns one element at a time; all exchange between
//     the lanes of an element goes through warp-private shared memory guarded by
//     __syncwarp only -- no block barriers in the steady state.
//   * the quadrature data (10 doubles per point, the dominant HBM stream) is
//     pulled by the bulk-async copy engine (cp.async.bulk -> SASS UBLKCP) into a
//     per-warp ring of Q1D z-slabs, each guarded by an mbarrier; the slab of the
//     NEXT element is requested as soon as the current one has been consumed, so
//     ~10 KB per warp are always in flight regardless of occupancy.
//   * gradients are taken with the collocation derivative matrix at the Gauss
//     points (u -> B u, then Dq along each direction), which needs one
//     interpolation instead of four; G = Dq B holds exactly because Q1D > p.
//   * x is gathered with the (essential-dof-encoding) int32 map one element ahead;
//     results are scattered with fp64 red.add (or stored as an E-vector).
#include "cdm_internal.hpp"
#include "kernels_common.cuh"

namespace
{
struct WarpTables
{
   double B[CDM_MAX_Q1D * CDM_MAX_D1D];    // B[q*D + d]
   double Dq[CDM_MAX_Q1D * CDM_MAX_Q1D];   // collocation derivative, Dq[i*Q + k] = l_k'(x_i) on the Gauss points
};

// mbarrier / bulk-async copy / red.add helpers: kernels_common.cuh (namespace cdmk); the D stream is evict-first in L2:
// it is read exactly once per apply and must not push out the x / y vectors, which are re-read (gather) and re-written
// (red.add) by neighbouring elements
using cdmk::smem_u32;
using cdmk::mbar_init;
#ifndef CDM_P3_IDX32
#define CDM_P3_IDX32 1
#endif
#ifndef CDM_P3_WAITALL
#define CDM_P3_WAITALL 1
#endif
using cdmk::mbar_expect_tx;
using cdmk::mbar_wait;
using cdmk::red_add_f64;
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) { cdmk::bulk_g2s_stream(dst, src, bytes, bar); }
// y[g] += v unless g < 0 (essential dof).  A predicated red (-DCDM_P3_PREDICATED_RED) measured the same as
// the branch within run-to-run noise on config 2, so the branch stays.
__device__ __forceinline__ void red_add_f64_if(double *y, int g, double v)
{
#ifndef CDM_P3_PREDICATED_RED
   if (g >= 0) { red_add_f64(y + g, v); }
#else
   cdmk::red_add_f64_pred(y, g, v);
#endif
}

// Lane roles (D1D = 4, Q1D = 5):
//   L1 lanes (dy,dz)  l < 16 : own one x-line of nodal values (gather / scatter, x contraction)
//   L2 lanes (qx,dz)  l < 20 : y contraction                   qx = l>>2, dz = l&3
//   L3 lanes (qx,qy)  l < 25 : z contraction and the quadrature-point work   qx = l%5, qy = l/5
// In every interpolation stage the coefficient index is a compile-time constant, so
// B comes from the constant bank (kernel parameter) and costs no registers or loads.
// Exchange layouts (doubles), chosen so that both the writer and the reader of each
// buffer hit 16 distinct 8-byte bank pairs per half warp:
//   P(qx,dy,dz) = qx + 5 dy + 20 dz        (L1 <-> L2)
//   R(qx,qy,dz) = qx + 5 qy + 28 dz        (L2 <-> L3)
// PH = true ("phased"): the quadrature-point work is split into two sync-free phases -- (A) gradients,
// D, point-wise products for all five slabs, (B) all transposed gradient contractions -- so the loads
// of different slabs overlap; fx, fy of the whole element are exchanged at once.
template <int D, int Q, int NW, bool DIFF, bool CONV, bool MASS, bool ATOMIC, bool PH>
__global__ void __launch_bounds__(NW * 32)
k_apply3d_warp(const WarpTables tb, const int64_t ne, const int32_t *__restrict__ gmap,
               const double *__restrict__ x, const double *__restrict__ Dg, const int slab,
               double *__restrict__ y)
{
   static_assert(D == 4 && Q == 5, "lane roles and exchange layouts are written for order 3");
   constexpr int Q2 = Q * Q, ND = D * D * D;
   constexpr int SU = Q * Q2;                                // u at the quadrature points [qz][qy][qx]
   constexpr int SRW = 28 * (D - 1) + Q2 + 3;                // R buffer (112)
   constexpr int SF = PH ? 2 * Q * Q2 : 2 * 2 * Q2;          // fx, fy: whole element (PH) or double-buffered slabs
   constexpr int SR = PH ? 0 : SRW;                          // PH: the R buffer lives inside the fx/fy region
   extern __shared__ __align__(128) unsigned char smraw[];
   const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
   const int warp_doubles = (Q * slab + SU + SF + SR + 15) & ~15;      // keeps every ring slab 16-byte aligned
   double *wbase = reinterpret_cast<double *>(smraw) + wib * warp_doubles;
   uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<double *>(smraw) + NW * warp_doubles) + wib * Q;
   double *ring = wbase, *sU = ring + Q * slab, *sF = sU + SU, *sP = sF, *sR = PH ? sF + 100 : sF + SF;
   const bool l1 = lane < D * D, l2 = lane < Q * D, l3 = lane < Q2;
   const int qx = l3 ? lane % Q : 0, qy = l3 ? lane / Q : 0;             // L3 role
   const int qx2 = l2 ? (lane >> 2) : 0, dz2 = lane & 3;                 // L2 role

   if (lane == 0)
   {
      for (int q = 0; q < Q; q++) { mbar_init(&bars[q], 1); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
   }
   __syncwarp();
   double dqx_row[Q], dqy_row[Q], dqx_col[Q], dqy_col[Q];
   #pragma unroll
   for (int k = 0; k < Q; k++)
   {
      dqx_row[k] = tb.Dq[qx * Q + k]; dqy_row[k] = tb.Dq[qy * Q + k];
      dqx_col[k] = tb.Dq[k * Q + qx]; dqy_col[k] = tb.Dq[k * Q + qy];
   }

   const int64_t gw = (int64_t)blockIdx.x * NW + wib, tw = (int64_t)gridDim.x * NW;
   const uint32_t slab_bytes = (uint32_t)slab * 8u;
   // prologue: request the first element's slabs and gather its x-lines
   int4 pg = make_int4(-1, -1, -1, -1);
   double px0 = 0.0, px1 = 0.0, px2 = 0.0, px3 = 0.0;
   if (gw < ne)
   {
      if (lane == 0)
      {
         #pragma unroll
         for (int q = 0; q < Q; q++)
         {
            mbar_expect_tx(&bars[q], slab_bytes);
            bulk_g2s(ring + q * slab, Dg + (gw * Q + q) * (int64_t)slab, slab_bytes, &bars[q]);
         }
      }
      if (l1)
      {
         pg = __ldg(reinterpret_cast<const int4 *>(gmap + gw * ND) + lane);
         px0 = (pg.x >= 0) ? __ldg(x + pg.x) : 0.0; px1 = (pg.y >= 0) ? __ldg(x + pg.y) : 0.0;
         px2 = (pg.z >= 0) ? __ldg(x + pg.z) : 0.0; px3 = (pg.w >= 0) ? __ldg(x + pg.w) : 0.0;
      }
   }

   uint32_t parity = 0;
   for (int64_t e = gw; e < ne; e += tw, parity ^= 1u)
   {
      const int64_t en = e + tw;
      const bool more = en < ne;
      const int4 g = pg;
      // ---- F1 (L1 lanes): x contraction of the own x-line, constant-bank coefficients
      if (l1)
      {
         #pragma unroll
         for (int q = 0; q < Q; q++)
         {
            sP[q + Q * lane] = tb.B[q * D + 0] * px0 + tb.B[q * D + 1] * px1 + tb.B[q * D + 2] * px2 + tb.B[q * D + 3] * px3;
         }
      }
      // prefetch the next element's gather (consumed one element later)
      if (more && l1)
      {
         pg = __ldg(reinterpret_cast<const int4 *>(gmap + en * ND) + lane);
         px0 = (pg.x >= 0) ? __ldg(x + pg.x) : 0.0; px1 = (pg.y >= 0) ? __ldg(x + pg.y) : 0.0;
         px2 = (pg.z >= 0) ? __ldg(x + pg.z) : 0.0; px3 = (pg.w >= 0) ? __ldg(x + pg.w) : 0.0;
      }
      __syncwarp();
      // ---- F2 (L2 lanes): y contraction
      if (l2)
      {
         double t[D];
         #pragma unroll
         for (int dy = 0; dy < D; dy++) { t[dy] = sP[qx2 + Q * dy + Q * D * dz2]; }
         #pragma unroll
         for (int q = 0; q < Q; q++)
         {
            sR[qx2 + Q * q + 28 * dz2] = tb.B[q * D + 0] * t[0] + tb.B[q * D + 1] * t[1] + tb.B[q * D + 2] * t[2] + tb.B[q * D + 3] * t[3];
         }
      }
      __syncwarp();
      // ---- F3 (L3 lanes): z contraction in registers, publish u
      double u[Q];
      {
         double v[D];
         #pragma unroll
         for (int dz = 0; dz < D; dz++) { v[dz] = sR[lane % Q2 + 28 * dz]; }
         #pragma unroll
         for (int qz = 0; qz < Q; qz++)
         {
            u[qz] = tb.B[qz * D + 0] * v[0] + tb.B[qz * D + 1] * v[1] + tb.B[qz * D + 2] * v[2] + tb.B[qz * D + 3] * v[3];
         }
      }
      if (l3)
      {
         #pragma unroll
         for (int qz = 0; qz < Q; qz++) { sU[qz * Q2 + lane] = u[qz]; }
      }
      __syncwarp();
      double out[Q];
      if (PH)
      {
         // ---- phase A: gradients, point-wise D for all slabs (no synchronisation inside)
         double fzv[Q], sv[Q];
         #pragma unroll
         for (int qz = 0; qz < Q; qz++)
         {
            double gx = 0.0, gy = 0.0, gz = 0.0;
            if (DIFF || CONV)
            {
               #pragma unroll
               for (int k = 0; k < Q; k++)
               {
                  gx += dqx_row[k] * sU[qz * Q2 + qy * Q + k];
                  gy += dqy_row[k] * sU[qz * Q2 + k * Q + qx];
                  gz += tb.Dq[qz * Q + k] * u[k];
               }
            }
            mbar_wait(&bars[qz], parity);
            const double *dp = ring + qz * slab + (l3 ? lane : 0);
            double fx = 0.0, fy = 0.0, fz = 0.0, s = 0.0;
            int c = 0;
            if (DIFF)
            {
               const double d0 = dp[0], d1 = dp[Q2], d2 = dp[2 * Q2], d3 = dp[3 * Q2], d4 = dp[4 * Q2], d5 = dp[5 * Q2];
               fx = d0 * gx + d1 * gy + d2 * gz;
               fy = d1 * gx + d3 * gy + d4 * gz;
               fz = d2 * gx + d4 * gy + d5 * gz;
               c = 6;
            }
            if (CONV) { s = dp[c * Q2] * gx + dp[(c + 1) * Q2] * gy + dp[(c + 2) * Q2] * gz; c += 3; }
            if (MASS) { s += dp[c * Q2] * u[qz]; }
            fzv[qz] = fz; sv[qz] = s;
            if (DIFF && l3) { sF[qz * Q2 + lane] = fx; sF[Q * Q2 + qz * Q2 + lane] = fy; }
         }
         __syncwarp();
         // the whole D tile of this element is consumed: request the next element's tile
         if (lane == 0 && more)
         {
            #pragma unroll
            for (int q = 0; q < Q; q++)
            {
               mbar_expect_tx(&bars[q], slab_bytes);
               bulk_g2s(ring + q * slab, Dg + (en * Q + q) * (int64_t)slab, slab_bytes, &bars[q]);
            }
         }
         // ---- phase B: transposed gradient contractions
         #pragma unroll
         for (int qz = 0; qz < Q; qz++)
         {
            double r = sv[qz];
            if (DIFF)
            {
               #pragma unroll
               for (int k = 0; k < Q; k++)
               {
                  r += dqx_col[k] * sF[qz * Q2 + qy * Q + k] + dqy_col[k] * sF[Q * Q2 + qz * Q2 + k * Q + qx];
                  r += tb.Dq[k * Q + qz] * fzv[k];
               }
            }
            out[qz] = r;
         }
         __syncwarp();                                       // fx/fy region is reused by the R buffer
      }
      else
      {
      // ---- slab loop: gradients, point-wise D, transposed gradients
      #pragma unroll
      for (int qz = 0; qz < Q; qz++) { out[qz] = 0.0; }
      #pragma unroll
      for (int qz = 0; qz < Q; qz++)
      {
         double gx = 0.0, gy = 0.0, gz = 0.0;
         if (DIFF || CONV)
         {
            #pragma unroll
            for (int k = 0; k < Q; k++)
            {
               gx += dqx_row[k] * sU[qz * Q2 + qy * Q + k];
               gy += dqy_row[k] * sU[qz * Q2 + k * Q + qx];
               gz += tb.Dq[qz * Q + k] * u[k];
            }
         }
         mbar_wait(&bars[qz], parity);
         const double *dp = ring + qz * slab + (l3 ? lane : 0);
         double fx = 0.0, fy = 0.0, fz = 0.0, s = 0.0;
         int c = 0;
         if (DIFF)
         {
            const double d0 = dp[0], d1 = dp[Q2], d2 = dp[2 * Q2], d3 = dp[3 * Q2], d4 = dp[4 * Q2], d5 = dp[5 * Q2];
            fx = d0 * gx + d1 * gy + d2 * gz;
            fy = d1 * gx + d3 * gy + d4 * gz;
            fz = d2 * gx + d4 * gy + d5 * gz;
            c = 6;
         }
         if (CONV) { s = dp[c * Q2] * gx + dp[(c + 1) * Q2] * gy + dp[(c + 2) * Q2] * gz; c += 3; }
         if (MASS) { s += dp[c * Q2] * u[qz]; }
         double r = s;
         if (DIFF)
         {
            double *sFx = sF + (qz & 1) * (2 * Q2), *sFy = sFx + Q2;
            if (l3) { sFx[lane] = fx; sFy[lane] = fy; }
         }
         __syncwarp();
         // slab qz of this element is consumed by every lane: request it for the next element
         if (lane == 0 && more)
         {
            mbar_expect_tx(&bars[qz], slab_bytes);
            bulk_g2s(ring + qz * slab, Dg + (en * Q + qz) * (int64_t)slab, slab_bytes, &bars[qz]);
         }
         if (DIFF)
         {
            const double *sFx = sF + (qz & 1) * (2 * Q2), *sFy = sFx + Q2;
            #pragma unroll
            for (int k = 0; k < Q; k++)
            {
               r += dqx_col[k] * sFx[qy * Q + k] + dqy_col[k] * sFy[k * Q + qx];
               out[k] += tb.Dq[qz * Q + k] * fz;
            }
         }
         out[qz] += r;
      }
      }
      // ---- B1 (L3 lanes): transposed z contraction in registers, publish w(qx,qy,dz)
      if (l3)
      {
         #pragma unroll
         for (int dz = 0; dz < D; dz++)
         {
            double t = 0.0;
            #pragma unroll
            for (int qz = 0; qz < Q; qz++) { t += tb.B[qz * D + dz] * out[qz]; }
            sR[lane + 28 * dz] = t;
         }
      }
      __syncwarp();
      // ---- B2 (L2 lanes): transposed y contraction
      if (l2)
      {
         double t[Q];
         #pragma unroll
         for (int q = 0; q < Q; q++) { t[q] = sR[qx2 + Q * q + 28 * dz2]; }
         #pragma unroll
         for (int dy = 0; dy < D; dy++)
         {
            double a = 0.0;
            #pragma unroll
            for (int q = 0; q < Q; q++) { a += tb.B[q * D + dy] * t[q]; }
            sP[qx2 + Q * dy + Q * D * dz2] = a;
         }
      }
      __syncwarp();
      // ---- B3 (L1 lanes): transposed x contraction of the own x-line, scatter
      if (l1)
      {
         double t[Q], yv[D];
         #pragma unroll
         for (int q = 0; q < Q; q++) { t[q] = sP[q + Q * lane]; }
         #pragma unroll
         for (int dx = 0; dx < D; dx++)
         {
            double a = 0.0;
            #pragma unroll
            for (int q = 0; q < Q; q++) { a += tb.B[q * D + dx] * t[q]; }
            yv[dx] = a;
         }
         if (ATOMIC)
         {
            if (g.x >= 0) { red_add_f64(y + g.x, yv[0]); }
            if (g.y >= 0) { red_add_f64(y + g.y, yv[1]); }
            if (g.z >= 0) { red_add_f64(y + g.z, yv[2]); }
            if (g.w >= 0) { red_add_f64(y + g.w, yv[3]); }
         }
         else
         {
            double2 *dst = reinterpret_cast<double2 *>(y + e * ND + D * lane);
            dst[0] = make_double2(yv[0], yv[1]);
            dst[1] = make_double2(yv[2], yv[3]);
         }
      }
      __syncwarp();                                          // sP / sR reuse by the next element
   }
}

// ---------------------------------------------------------------------------------------------
// k_apply3d_warp_bg -- "register-z" variant (kernel option 3, the default for order 3).
// u, du/dx, du/dy, du/dz are produced by separate B / G contractions: x and y in two exchange
// stages with constant-bank coefficients, z in registers, so every lane ends up holding the four
// fields on its own z-column of quadrature points.  The point-wise D product and the transposed
// z contraction then run entirely in registers: no gradient exchange through shared memory, no
// per-lane coefficient registers, 6 warp syncs per element.  Shared-memory instructions per
// element: 45 (forward) + 50 (D) + 45 (backward), against ~200 in the collocated variants.
struct WarpTablesBG
{
   double B[CDM_MAX_Q1D * CDM_MAX_D1D];    // B[q*D + d]
   double G[CDM_MAX_Q1D * CDM_MAX_D1D];    // G[q*D + d]
};

template <int NW, bool DIFF, bool CONV, bool MASS, bool ATOMIC>
__global__ void __launch_bounds__(NW * 32)
k_apply3d_warp_bg(const WarpTablesBG tb, const int64_t ne, const int32_t *__restrict__ gmap,
                  const double *__restrict__ x, const double *__restrict__ Dg, const int slab,
                  double *__restrict__ y)
{
   constexpr int D = 4, Q = 5, Q2 = Q * Q, ND = D * D * D;
   constexpr bool GRAD = DIFF || CONV;
   constexpr int PS = 80, RS = 112;                          // one P-layout / R-layout array (doubles)
   extern __shared__ __align__(128) unsigned char smraw[];
   const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
   const int warp_doubles = (Q * slab + 3 * RS + 2 * PS + 15) & ~15;
   double *wbase = reinterpret_cast<double *>(smraw) + wib * warp_doubles;
   uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<double *>(smraw) + NW * warp_doubles) + wib * Q;
   double *ring = wbase, *sR0 = ring + Q * slab, *sR1 = sR0 + RS, *sR2 = sR1 + RS, *sP0 = sR2 + RS, *sP1 = sP0 + PS;
   const bool l1 = lane < D * D, l2 = lane < Q * D, l3 = lane < Q2;
   const int l3i = l3 ? lane : 0;                                        // L3 role: (qx,qy) = lane
   const int qx2 = l2 ? (lane >> 2) : 0, dz2 = lane & 3;                 // L2 role

   if (lane == 0)
   {
      for (int q = 0; q < Q; q++) { mbar_init(&bars[q], 1); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
   }
   __syncwarp();
   // element indices: 64-bit by default; 32-bit as a measured variant (CDM_P3_IDX32; the launcher then refuses > 2e9 elements)
#if CDM_P3_IDX32
   using eidx = int;
#else
   using eidx = int64_t;
#endif
   const eidx gw = (eidx)blockIdx.x * NW + wib, tw = (eidx)gridDim.x * NW, nel = (eidx)ne;
   const uint32_t slab_bytes = (uint32_t)slab * 8u;
   // gather pipeline: indices are fetched two elements ahead, values one element ahead, so that no
   // load is consumed in the element that issued it
   int4 pg = make_int4(-1, -1, -1, -1), pgn = make_int4(-1, -1, -1, -1);
   double px[D] = {0.0, 0.0, 0.0, 0.0};
   if (gw < nel)
   {
      if (lane == 0)
      {
         #pragma unroll
         for (int q = 0; q < Q; q++)
         {
            mbar_expect_tx(&bars[q], slab_bytes);
            bulk_g2s(ring + q * slab, Dg + ((int64_t)gw * Q + q) * slab, slab_bytes, &bars[q]);
         }
      }
      if (l1)
      {
         pg = __ldg(reinterpret_cast<const int4 *>(gmap + (int64_t)gw * ND) + lane);
         if (gw + tw < nel) { pgn = __ldg(reinterpret_cast<const int4 *>(gmap + (int64_t)(gw + tw) * ND) + lane); }
         px[0] = (pg.x >= 0) ? __ldg(x + pg.x) : 0.0; px[1] = (pg.y >= 0) ? __ldg(x + pg.y) : 0.0;
         px[2] = (pg.z >= 0) ? __ldg(x + pg.z) : 0.0; px[3] = (pg.w >= 0) ? __ldg(x + pg.w) : 0.0;
      }
   }

   uint32_t parity = 0;
   for (eidx e = gw; e < nel; e += tw, parity ^= 1u)
   {
      const eidx en = e + tw;
      const bool more = en < nel;
      const int4 g = pg;
      // ---- F1 (L1 lanes): x contraction of the own x-line with B and G
      if (l1)
      {
         #pragma unroll
         for (int q = 0; q < Q; q++)
         {
            double tB = 0.0, tG = 0.0;
            #pragma unroll
            for (int d = 0; d < D; d++) { tB += tb.B[q * D + d] * px[d]; if (GRAD) { tG += tb.G[q * D + d] * px[d]; } }
            sP0[q + Q * lane] = tB;
            if (GRAD) { sP1[q + Q * lane] = tG; }
         }
      }
      if (more && l1)
      {
         pg = pgn;                                           // indices of element en, loaded one element ago
         if (en + tw < nel) { pgn = __ldg(reinterpret_cast<const int4 *>(gmap + (int64_t)(en + tw) * ND) + lane); }
         px[0] = (pg.x >= 0) ? __ldg(x + pg.x) : 0.0; px[1] = (pg.y >= 0) ? __ldg(x + pg.y) : 0.0;
         px[2] = (pg.z >= 0) ? __ldg(x + pg.z) : 0.0; px[3] = (pg.w >= 0) ? __ldg(x + pg.w) : 0.0;
      }
      __syncwarp();
      // ---- F2 (L2 lanes): y contraction -> (B B), (G B), (B G)
      if (l2)
      {
         double tB[D], tG[D];
         #pragma unroll
         for (int dy = 0; dy < D; dy++)
         {
            tB[dy] = sP0[qx2 + Q * dy + Q * D * dz2];
            if (GRAD) { tG[dy] = sP1[qx2 + Q * dy + Q * D * dz2]; }
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
            sR0[qx2 + Q * q + 28 * dz2] = vbb;
            if (GRAD) { sR1[qx2 + Q * q + 28 * dz2] = vgb; sR2[qx2 + Q * q + 28 * dz2] = vbg; }
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
            vbb[dz] = sR0[l3i + 28 * dz];
            if (GRAD) { vgb[dz] = sR1[l3i + 28 * dz]; vbg[dz] = sR2[l3i + 28 * dz]; }
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
      // ---- point-wise D at the lane's five quadrature points (registers only)
#if CDM_P3_WAITALL
      // the five slabs were requested one element ago: wait for all of them first, so that the wait loops do not
      // fence the loads of one slab from the products of the previous one
      #pragma unroll
      for (int qz = 0; qz < Q; qz++) { mbar_wait(&bars[qz], parity); }
#endif
      #pragma unroll
      for (int qz = 0; qz < Q; qz++)
      {
#if !CDM_P3_WAITALL
         mbar_wait(&bars[qz], parity);
#endif
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
      if (lane == 0 && more)
      {
         #pragma unroll
         for (int q = 0; q < Q; q++)
         {
            mbar_expect_tx(&bars[q], slab_bytes);
            bulk_g2s(ring + q * slab, Dg + ((int64_t)en * Q + q) * slab, slab_bytes, &bars[q]);
         }
      }
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
            sR2[lane + 28 * dz] = wb;
            if (DIFF) { sR0[lane + 28 * dz] = wx; sR1[lane + 28 * dz] = wy; }
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
            wb[q] = sR2[qx2 + Q * q + 28 * dz2];
            if (DIFF) { wx[q] = sR0[qx2 + Q * q + 28 * dz2]; wy[q] = sR1[qx2 + Q * q + 28 * dz2]; }
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
            sP1[qx2 + Q * dy + Q * D * dz2] = a2;
            if (DIFF) { sP0[qx2 + Q * dy + Q * D * dz2] = a1; }
         }
      }
      __syncwarp();
      // ---- B3 (L1 lanes): transposed x contraction of the own x-line, scatter
      if (l1)
      {
         double a1[Q], a2[Q], yv[D];
         #pragma unroll
         for (int q = 0; q < Q; q++) { a2[q] = sP1[q + Q * lane]; if (DIFF) { a1[q] = sP0[q + Q * lane]; } }
         #pragma unroll
         for (int dx = 0; dx < D; dx++)
         {
#ifdef CDM_P3_SPLIT_B3
            double a = 0.0, ag = 0.0;
            #pragma unroll
            for (int q = 0; q < Q; q++) { a += tb.B[q * D + dx] * a2[q]; if (DIFF) { ag += tb.G[q * D + dx] * a1[q]; } }
            yv[dx] = a + ag;
#else
            double a = 0.0;
            #pragma unroll
            for (int q = 0; q < Q; q++) { a += tb.B[q * D + dx] * a2[q]; if (DIFF) { a += tb.G[q * D + dx] * a1[q]; } }
            yv[dx] = a;
#endif
         }
         if (ATOMIC)
         {
            red_add_f64_if(y, g.x, yv[0]);
            red_add_f64_if(y, g.y, yv[1]);
            red_add_f64_if(y, g.z, yv[2]);
            red_add_f64_if(y, g.w, yv[3]);
         }
         else
         {
            double2 *dst = reinterpret_cast<double2 *>(y + (int64_t)e * ND + D * lane);
            dst[0] = make_double2(yv[0], yv[1]);
            dst[1] = make_double2(yv[2], yv[3]);
         }
      }
      __syncwarp();                                          // P buffers are rewritten by the next element's F1
   }
}

template <int NW, bool DIFF, bool CONV, bool MASS, bool ATOMIC>
int launch_bg(cdm_op *op, const WarpTablesBG &tb, const int32_t *gmap, const double *xL, double *out)
{
   cdm_space *sp = op->sp;
   cdm_ctx *ctx = sp->ctx;
   auto kern = k_apply3d_warp_bg<NW, DIFF, CONV, MASS, ATOMIC>;
   const int warp_doubles = (5 * op->slab + 3 * 112 + 2 * 80 + 15) & ~15;
   const size_t smem = (size_t)(NW * warp_doubles) * sizeof(double) + (size_t)NW * 5 * sizeof(uint64_t);
   int blocks_per_sm = 0;
   { const int rc = cdm_kernel_cfg(ctx, (const void *)kern, NW * 32, smem, "k_apply3d_warp_bg", &blocks_per_sm); if (rc) { return rc; } }
   // element range [e0, e1): the kernel sees a shifted view of the per-element arrays
   const int64_t e0 = op->range_on ? op->e_begin : 0, e1 = op->range_on ? op->e_end : sp->ne;
   const int64_t n = e1 - e0;
   if (n <= 0) { return CDM_OK; }
#if CDM_P3_IDX32
   if (n > 2000000000LL) { return 1; }
#endif
   int64_t grid = (int64_t)ctx->sm_count * blocks_per_sm;
   const int64_t need = (n + NW - 1) / NW;
   if (grid > need) { grid = need; }
   if (op->grid_cap > 0 && grid > op->grid_cap) { grid = op->grid_cap; }
   if (ctx->time_main) { cudaEventRecord(ctx->evk0, ctx->stream); }
   kern<<<(unsigned)grid, NW * 32, smem, ctx->stream>>>(tb, n, gmap + e0 * 64, xL, op->D_dev + e0 * 5 * (int64_t)op->slab,
                                                        op->slab, ATOMIC ? out : out + e0 * 64);
   if (ctx->time_main) { cudaEventRecord(ctx->evk1, ctx->stream); }
   ctx->launches++;
   CDM_CUDA(ctx, cudaGetLastError());
   return CDM_OK;
}

// collocation derivative matrix on the Gauss points: Dq[i][k] = l_k'(x_i)
void collocation_matrix(int q1d, const double *xq, double *Dq)
{
   std::vector<double> bw(q1d);
   for (int j = 0; j < q1d; j++)
   {
      double c = 1.0;
      for (int k = 0; k < q1d; k++) if (k != j) { c *= xq[j] - xq[k]; }
      bw[j] = 1.0 / c;
   }
   for (int i = 0; i < q1d; i++)
   {
      double diag = 0.0;
      for (int k = 0; k < q1d; k++)
      {
         if (k == i) { continue; }
         Dq[i * q1d + k] = (bw[k] / bw[i]) / (xq[i] - xq[k]);
         diag -= Dq[i * q1d + k];
      }
      Dq[i * q1d + i] = diag;
   }
}

template <int D, int Q, int NW, bool DIFF, bool CONV, bool MASS, bool ATOMIC, bool PH>
int launch(cdm_op *op, const WarpTables &tb, const int32_t *gmap, const double *xL, double *out)
{
   cdm_space *sp = op->sp;
   cdm_ctx *ctx = sp->ctx;
   auto kern = k_apply3d_warp<D, Q, NW, DIFF, CONV, MASS, ATOMIC, PH>;
   const int warp_doubles = PH ? ((Q * op->slab + Q * Q * Q + 2 * Q * Q * Q + 15) & ~15)
                               : ((Q * op->slab + Q * Q * Q + 2 * 2 * Q * Q + (28 * (D - 1) + Q * Q + 3) + 15) & ~15);
   const size_t smem = (size_t)(NW * warp_doubles) * sizeof(double) + (size_t)NW * Q * sizeof(uint64_t);
   int blocks_per_sm = 0;
   { const int rc = cdm_kernel_cfg(ctx, (const void *)kern, NW * 32, smem, "k_apply3d_warp", &blocks_per_sm); if (rc) { return rc; } }
   int64_t grid = (int64_t)ctx->sm_count * blocks_per_sm;
   const int64_t need = (sp->ne + NW - 1) / NW;
   if (grid > need) { grid = need; }
   if (op->grid_cap > 0 && grid > op->grid_cap) { grid = op->grid_cap; }
   if (ctx->time_main) { cudaEventRecord(ctx->evk0, ctx->stream); }
   kern<<<(unsigned)grid, NW * 32, smem, ctx->stream>>>(tb, sp->ne, gmap, xL, op->D_dev, op->slab, out);
   if (ctx->time_main) { cudaEventRecord(ctx->evk1, ctx->stream); }
   ctx->launches++;
   CDM_CUDA(ctx, cudaGetLastError());
   return CDM_OK;
}

__global__ void __launch_bounds__(256)
k_restrict_transpose_p3(int64_t ndof, const int32_t *__restrict__ offsets, const int32_t *__restrict__ indices,
                        const double *__restrict__ yE, double *__restrict__ yL)
{
   const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (g >= ndof) { return; }
   double s = 0.0;
   for (int32_t j = offsets[g]; j < offsets[g + 1]; j++) { s += yE[indices[j]]; }
   yL[g] = s;
}
}  // namespace

#define NWARPS 4
#define DISPATCH_ONE(D_, Q_, DF, CV, MS)                                                                   \
   rc = ph ? (atomic ? launch<D_, Q_, NWARPS, DF, CV, MS, true, true>(op, tb, gmap, xL, out)              \
                     : launch<D_, Q_, NWARPS, DF, CV, MS, false, true>(op, tb, gmap, xL, out))            \
           : (atomic ? launch<D_, Q_, NWARPS, DF, CV, MS, true, false>(op, tb, gmap, xL, out)             \
                     : launch<D_, Q_, NWARPS, DF, CV, MS, false, false>(op, tb, gmap, xL, out))
#define DISPATCH_FLAGS(D_, Q_)                                                                            \
   if (op->has_diff && op->has_conv && op->has_mass) { DISPATCH_ONE(D_, Q_, true, true, true); }          \
   else if (op->has_diff && !op->has_conv && op->has_mass) { DISPATCH_ONE(D_, Q_, true, false, true); }   \
   else if (!op->has_diff && !op->has_conv && op->has_mass) { DISPATCH_ONE(D_, Q_, false, false, true); } \
   else if (op->has_diff && !op->has_conv && !op->has_mass) { DISPATCH_ONE(D_, Q_, true, false, false); } \
   else { rc = 1; }

// returns 1 when this operator is not covered (caller falls back to the generic kernel)
int cdm_k_apply_p3(cdm_op *op, const int32_t *gmap, const double *xL, double *yL)
{
   cdm_space *sp = op->sp;
   cdm_ctx *ctx = sp->ctx;
   if (sp->dim != 3 || sp->p != 3) { return 1; }
   if (!((op->has_diff && op->has_mass) || (!op->has_conv && (op->has_diff || op->has_mass)))) { return 1; }
   if (op->has_conv && !(op->has_diff && op->has_mass)) { return 1; }
   WarpTables tb;
   memset(&tb, 0, sizeof(tb));
   for (int i = 0; i < sp->q1d * sp->d1d; i++) { tb.B[i] = sp->B[i]; }
   collocation_matrix(sp->q1d, sp->qx.data(), tb.Dq);
   const bool atomic = op->scatter_mode == 1;
   const bool ph = op->kernel_variant >= 2;
   double *out = yL;
   if (atomic) { if (!op->range_on) { CDM_CUDA(ctx, cudaMemsetAsync(yL, 0, sizeof(double) * (size_t)sp->ndof, ctx->stream)); } }
   else
   {
      if (!op->yE_dev) { CDM_CUDA(ctx, cudaMalloc(&op->yE_dev, sizeof(double) * (size_t)sp->ne * sp->nd)); }
      out = op->e_out ? op->e_out : op->yE_dev;
   }
   int rc = 0;
   if (op->kernel_variant >= 3)
   {
      WarpTablesBG tg;
      memset(&tg, 0, sizeof(tg));
      for (int i = 0; i < sp->q1d * sp->d1d; i++) { tg.B[i] = sp->B[i]; tg.G[i] = sp->G[i]; }
#define BG_ONE(DF, CV, MS)                                                              \
      rc = atomic ? launch_bg<NWARPS, DF, CV, MS, true>(op, tg, gmap, xL, out)          \
                  : launch_bg<NWARPS, DF, CV, MS, false>(op, tg, gmap, xL, out)
      if (op->has_diff && op->has_conv && op->has_mass) { BG_ONE(true, true, true); }
      else if (op->has_diff && !op->has_conv && op->has_mass) { BG_ONE(true, false, true); }
      else if (!op->has_diff && !op->has_conv && op->has_mass) { BG_ONE(false, false, true); }
      else if (op->has_diff && !op->has_conv && !op->has_mass) { BG_ONE(true, false, false); }
      else { rc = 1; }
#undef BG_ONE
   }
   else
   {
      DISPATCH_FLAGS(4, 5)
   }
   if (rc) { return rc; }
   if (!atomic && !op->e_out)
   {
      const unsigned nb = (unsigned)((sp->ndof + 255) / 256);
      k_restrict_transpose_p3<<<nb, 256, 0, ctx->stream>>>(sp->ndof, sp->offsets_dev, sp->indices_dev, op->yE_dev, yL);
      ctx->launches++;
      CDM_CUDA(ctx, cudaGetLastError());
   }
   return CDM_OK;
}
