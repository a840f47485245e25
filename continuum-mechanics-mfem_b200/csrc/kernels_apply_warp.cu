// kernels_apply_warp.cu -- the register-z / bulk-async apply kernel of kernels_apply_p3.cu
// generalised to every order p = 1..6 in 3D (kernel option 4).
//
// One *group* of T = Q1D^2 threads owns one element at a time:
//   p = 1 : T =  9, three groups per warp        p = 4 : T = 36, three warps for two elements (32 + 32 + 2 x 4 lanes)
//   p = 2 : T = 16, two groups per warp          p = 5 : T = 49, two warps per group
//   p = 3 : T = 25, one group per warp           p = 6 : T = 64, two warps per group
// Sub-warp groups synchronise with __syncwarp, two-warp groups and the packed order-4 blocks with the block barrier.
// Everything else is the design documented in kernels_apply_p3.cu: quadrature data streamed by cp.async.bulk into a
// per-group ring of Q1D z-slabs guarded by mbarriers, x / y contractions through conflict-aware exchange buffers with
// constant-bank coefficients (orders 5, 6: in place, see GroupCfg), z contraction + point-wise D + transposed z
// contraction in registers, two-level gather prefetch, fp64 red.add scatter (or E-vector stores).
//
// Compile-time variants (-D...; defaults are the measured best, profiles/r02_sweep_config4.md):
//   CDM_G4_TRIO 1        order 4: three warps for two elements (0: two warps per element, 80.7 instead of 95.3 %)
//   CDM_G5_INPLACE 1, CDM_G6_INPLACE 1   y contractions in place (0: separate P buffers, one resident group fewer)
//   CDM_G_LANEMAP 1      padded half-warp lane maps of the x / y roles at orders 5, 6
//   CDM_G_WAITALL 1, CDM_G5_WAITALL 0    point-wise stage waits for all slab barriers first
//   CDM_G_ALIAS 1        (only without INPLACE) P and R buffers share a region, two more barriers per element
//   CDM_G5_PENTA 0       order 5: five warps for three elements (79.1 instead of 81.5 %)
//   CDM_G4_FUSEZ / G5 / G6 0   z contraction, point-wise D, transposed z contraction fused per level (slower; always on for
//                              order 4 with E-vector output, where the separate stages spill)
//   CDM_G_BALANCED 0     order 4 without TRIO: both warps run the x / y stages (neutral)
//   CDM_G5_IDX32 0       32-bit element indices at order 5 (order 4 with red.add output always uses them)
#include "cdm_internal.hpp"
#include "kernels_common.cuh"
#include <type_traits>

namespace
{
struct GroupTables
{
   double B[CDM_MAX_Q1D * CDM_MAX_D1D];
   double G[CDM_MAX_Q1D * CDM_MAX_D1D];
};

// mbarrier / bulk-async copy / red.add helpers: kernels_common.cuh (namespace cdmk)
__device__ __forceinline__ void g_mbar_init(uint64_t *bar, uint32_t count) { cdmk::mbar_init(bar, count); }
__device__ __forceinline__ void g_mbar_expect_tx(uint64_t *bar, uint32_t bytes) { cdmk::mbar_expect_tx(bar, bytes); }
__device__ __forceinline__ void g_mbar_wait(uint64_t *bar, uint32_t parity) { cdmk::mbar_wait(bar, parity); }
__device__ __forceinline__ void g_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) { cdmk::bulk_g2s_stream(dst, src, bytes, bar); }
__device__ __forceinline__ void g_red_add(double *addr, double v) { cdmk::red_add_f64(addr, v); }

#ifndef CDM_G4_TRIO
#define CDM_G4_TRIO 1
#endif
#ifndef CDM_G5_IDX32
#define CDM_G5_IDX32 0
#endif
#ifndef CDM_G5_PENTA
#define CDM_G5_PENTA 0               // measured: 79.1 % against 81.5 % for two warps per element (10 instead of 12 warps per SM)
#endif
template <int P> struct GroupCfg
{
   static constexpr int D = P + 1, Q = P + 2, T = Q * Q, ND = D * D * D;
   static constexpr int EPW = (T <= 10) ? 3 : (T <= 16 ? 2 : 1);   // groups per warp
   static constexpr int WPG = (T <= 32) ? 1 : 2;                   // warps per group
   #ifndef CDM_G4_GPB
#define CDM_G4_GPB 2
#endif
#ifndef CDM_G5_GPB
#define CDM_G5_GPB 1
#endif
   static constexpr int GPB = (P <= 3) ? 4 * EPW : (P == 4 ? CDM_G4_GPB : (P == 5 ? (CDM_G5_PENTA ? 3 : CDM_G5_GPB) : 1));   // groups per block (two-warp groups: the block IS the group)
   // TRIO (p=4): the 36 columns of an element are 32 + 4 lanes, so the second warp of a group issued the whole z /
   // point-wise stream for four lanes.  A block is now three warps for two elements: warp 0 / warp 1 own columns 0..31 of
   // element 0 / 1 (and their x / y roles), lanes 0..3 / 4..7 of warp 2 own columns 32..35 of element 0 / 1.  Three
   // instead of four instruction streams per two elements, block-wide barriers, 168 instead of 128 registers per thread.
   static constexpr bool TRIO = (P == 4) && CDM_G4_TRIO;
   // PENTA (p=5): 49 columns are 32 + 17 lanes.  A block is five warps for three elements: warp g < 3 owns columns 0..31 of
   // element g and the first two half-warps of its x / y roles; half-warp k < 3 of the two helper warps owns columns 32..47
   // and the third x / y half-warp of element k; three more lanes own column 48 of the three elements.  Five instead of six
   // instruction streams per three elements (group size 4420 doubles = 4 mod 16: the three column-48 lanes hit distinct banks).
   static constexpr bool PENTA = (P == 5) && CDM_G5_PENTA;
   static constexpr bool PACKED = TRIO || PENTA;               // warps shared between the elements of a block, block-wide barriers
   static constexpr int THREADS = TRIO ? 96 : (PENTA ? 160 : ((WPG == 1) ? 32 * (GPB / EPW) : 64 * GPB));
// resident blocks per SM the register budget is sized for.  p=5: 5 blocks fit the shared memory, but the
// 204-register cap of MINB = 5 made ptxas spill 200 B inside the element loop (53 % of the roofline);
// MINB = 4 -> 160 registers, no spills, 73 %.  p=4 (two groups per block): MINB 3 / 4 / 5 -> 46 % / 76 % / 56 %; one group per block x 8 blocks: 52 %.
#ifndef CDM_G_BALANCED
#define CDM_G_BALANCED 0
#endif
#ifndef CDM_G_ALIAS
#define CDM_G_ALIAS 1
#endif
#ifndef CDM_G_SPLITP
#define CDM_G_SPLITP 5
#endif
#ifndef CDM_G_WAITALL
#define CDM_G_WAITALL 1
#endif
#ifndef CDM_G5_WAITALL
#define CDM_G5_WAITALL 0             // p=5: 192 instead of 161 registers, i.e. 5 instead of 6 resident groups
#endif
#ifndef CDM_G5_INPLACE
#define CDM_G5_INPLACE 1
#endif
#ifndef CDM_G6_INPLACE
#define CDM_G6_INPLACE 1
#endif
#ifndef CDM_G5_MINB
#define CDM_G5_MINB (CDM_G5_PENTA ? 1 : ((CDM_G_ALIAS && !CDM_G5_INPLACE) ? 6 : 4))
#endif
#ifndef CDM_G6_MINB
#define CDM_G6_MINB ((CDM_G_ALIAS && !CDM_G6_INPLACE) ? 4 : 3)
#endif
#ifndef CDM_G4_MINB
#define CDM_G4_MINB (CDM_G4_TRIO ? 3 : 4)     // TRIO: 165 registers under (96, 3); four blocks still fit (ptxas spills ~400 B under (96, 4))
#endif
   static constexpr int MINB = (P < 4) ? 4 : (P == 4 ? CDM_G4_MINB : (P == 5 ? CDM_G5_MINB : CDM_G6_MINB));        // resident blocks per SM the register budget is sized for
   // strides of the exchange layouts P(qx, line) = qx + PST line, R(qx,qy,dz) = qx + Q qy + RSTR dz.  For the
   // two-warp groups they come from a brute-force search over the 64-bit bank pattern of the four access
   // shapes (F1 write / F2 read of P, F2 write / F3 read of R): wavefronts per access, old -> new:
   // p=4  P 8 -> 4, R 4 -> 2 (ideal 4, 2);  p=6  P 41 -> 8, R 11 -> 4 (ideal 8, 4);  p=5 unchanged (8 / 5, ideal 6 / 3).
   // INPLACE (p = 5, 6, where shared memory bounds the resident groups): the P content lives in R0 / R1 with the R strides,
   // P(q; dy,dz) = q + Q dy + RSTR dz.  An L2 thread (qx,dz) of F2 / B2 then reads and writes only entries
   // qx + Q j + RSTR dz of its own (qx,dz), all inputs before the first output: the y contractions are in-place
   // transforms, no P buffers and no extra barrier.  p=5: 35.4 instead of 38.8 KB per group = 6 instead of 5 groups per
   // SM with the same code shape (ptxas spills under the 168-register cap of MINB >= 5, so MINB stays 4 and the 6th group
   // fits by itself at 160 registers); p=6: 52.8 instead of 66.1 KB = 4 instead of 3 groups.  Measured at 8 M dofs:
   // p=5 74 -> 81 %, p=6 74 -> 86 % of the HBM roofline (82 / 90 % with the lane maps below).
   static constexpr bool INPLACE = (P == 5 && CDM_G5_INPLACE) || (P == 6 && CDM_G6_INPLACE);
   // Lane maps of the L1 (x-line) and L2 (qx,dz) roles when INPLACE: a half-warp is one 64-bit shared-memory
   // wavefront, so it is given only as many lines as have distinct banks, the rest of its lanes idle:
   //   L1MAP 1 (p=5): 12 of 16 lanes, dy = r % D, dz = 2h + r / D   (banks 7 dy + 8 (dz % 2): 12 distinct)
   //   L1MAP 2 (p=6): 14 of 16 lanes, dy = 2h + r / D, dz = r % D   (banks 8 (dy % 2) + 7 dz: 14 distinct)
   //   L2MAP 1 (p=5): 14 of 16 lanes, qx = r % Q, dz = 2h + r / Q   (banks qx + 8 (dz % 2))
   // with h = half-warp of the group, r = lane in it.  Wavefronts per element of the four exchange shapes:
   // p=5 558 -> 426, p=6 696 -> 600, both the ideal count for their lane numbers.
#ifndef CDM_G_LANEMAP
#define CDM_G_LANEMAP 1
#endif
   static constexpr int L1MAP = (INPLACE && CDM_G_LANEMAP) ? (P == 5 ? 1 : 2) : 0;
   static constexpr int L2MAP = (INPLACE && CDM_G_LANEMAP && P == 5) ? 1 : 0;
   static constexpr int RSTR = (P == 3) ? 28 : (P == 4 ? 45 : (P == 6 ? 71 : (INPLACE ? (L2MAP ? 56 : 58) : (Q * Q + ((Q * Q) % 2 == 0 ? 1 : 2)))));
   static constexpr int PST = (P == 4) ? 9 : (P == 6 ? 17 : Q);
   static constexpr int PSY = INPLACE ? Q : PST, PSZ = INPLACE ? RSTR : PST * D;   // P(q; dy,dz) = q + PSY dy + PSZ dz
   static constexpr int RS = ((D - 1) * RSTR + Q * Q + 1) & ~1;
   static constexpr int PS = (PST * D * D + 1) & ~1;
   // ALIAS (variant, off when INPLACE): the P and the R exchange buffers share one region; F2 and B2 read all their
   // inputs into registers, meet at the group barrier, and write afterwards (two more barriers per element).  p=6: 4
   // groups, 87 %; p=5: the barriers push ptxas to 254 registers (or spills under a cap): 67 %.  Superseded by INPLACE.
   static constexpr bool ALIAS = (P >= 5) && CDM_G_ALIAS && !INPLACE;
   // the D stage waits for all slab barriers before its first product (the slabs were requested one element ago): the
   // wait loops no longer fence the loads of one slab from the products of the previous one
   static constexpr bool WAITALL = (P == 5) ? CDM_G5_WAITALL : CDM_G_WAITALL;
#ifndef CDM_G4_FUSEZ
#define CDM_G4_FUSEZ 0
#endif
#ifndef CDM_G5_FUSEZ
#define CDM_G5_FUSEZ 0
#endif
#ifndef CDM_G6_FUSEZ
#define CDM_G6_FUSEZ 0
#endif
   // z contraction, point-wise D and transposed z contraction fused per quadrature level (see the kernel)
   static constexpr bool FUSEZ = (P == 4 && CDM_G4_FUSEZ) || (P == 5 && CDM_G5_FUSEZ) || (P == 6 && CDM_G6_FUSEZ);
   static constexpr int XS = INPLACE ? 3 * RS : (ALIAS ? (3 * RS > 2 * PS ? 3 * RS : 2 * PS) : 3 * RS + 2 * PS);   // exchange doubles per group
};

// ---- the four exchange stages as functions of a compile-time output range [LO, HI): a two-warp group can then give
// each of its warps one half of the outputs of a stage (warp-uniform branch, coefficients stay in the constant bank)
template <int P, bool GRAD, int QLO, int QHI>
__device__ __forceinline__ void grp_f1(const GroupTables &tb, const double *px, double *sP0, double *sP1, int pb1)
{
   constexpr int D = P + 1;
   #pragma unroll
   for (int q = QLO; q < QHI; q++)
   {
      double tB = 0.0, tG = 0.0;
      #pragma unroll
      for (int d = 0; d < D; d++) { tB += tb.B[q * D + d] * px[d]; if (GRAD) { tG += tb.G[q * D + d] * px[d]; } }
      sP0[q + pb1] = tB;
      if (GRAD) { sP1[q + pb1] = tG; }
   }
}
template <int P, bool GRAD>
__device__ __forceinline__ void grp_f2_load(const double *sP0, const double *sP1, int qx2, int dz2, double *tB, double *tG)
{
   constexpr int D = P + 1, PSY = GroupCfg<P>::PSY, PSZ = GroupCfg<P>::PSZ;
   #pragma unroll
   for (int dy = 0; dy < D; dy++)
   {
      tB[dy] = sP0[qx2 + PSY * dy + PSZ * dz2];
      if (GRAD) { tG[dy] = sP1[qx2 + PSY * dy + PSZ * dz2]; }
   }
}
template <int P, bool GRAD, int QLO, int QHI>
__device__ __forceinline__ void grp_f2(const GroupTables &tb, const double *tB, const double *tG, double *sR0, double *sR1, double *sR2,
                                       int qx2, int dz2)
{
   constexpr int D = P + 1, Q = P + 2, RSTR = GroupCfg<P>::RSTR;
   #pragma unroll
   for (int q = QLO; q < QHI; q++)
   {
      double vbb = 0.0, vgb = 0.0, vbg = 0.0;
      #pragma unroll
      for (int dy = 0; dy < D; dy++)
      {
         vbb += tb.B[q * D + dy] * tB[dy];
         if (GRAD) { vgb += tb.B[q * D + dy] * tG[dy]; vbg += tb.G[q * D + dy] * tB[dy]; }
      }
      sR0[qx2 + Q * q + RSTR * dz2] = vbb;
      if (GRAD) { sR1[qx2 + Q * q + RSTR * dz2] = vgb; sR2[qx2 + Q * q + RSTR * dz2] = vbg; }
   }
}
template <int P, bool DIFF>
__device__ __forceinline__ void grp_b2_load(const double *sR0, const double *sR1, const double *sR2, int qx2, int dz2,
                                            double *wx, double *wy, double *wb)
{
   constexpr int Q = P + 2, RSTR = GroupCfg<P>::RSTR;
   #pragma unroll
   for (int q = 0; q < Q; q++)
   {
      wb[q] = sR2[qx2 + Q * q + RSTR * dz2];
      if (DIFF) { wx[q] = sR0[qx2 + Q * q + RSTR * dz2]; wy[q] = sR1[qx2 + Q * q + RSTR * dz2]; }
   }
}
template <int P, bool DIFF, int DLO, int DHI>
__device__ __forceinline__ void grp_b2(const GroupTables &tb, const double *wx, const double *wy, const double *wb, double *sP0, double *sP1,
                                       int qx2, int dz2)
{
   constexpr int D = P + 1, Q = P + 2, PSY = GroupCfg<P>::PSY, PSZ = GroupCfg<P>::PSZ;
   #pragma unroll
   for (int dy = DLO; dy < DHI; dy++)
   {
      double a1 = 0.0, a2 = 0.0;
      #pragma unroll
      for (int q = 0; q < Q; q++)
      {
         a2 += tb.B[q * D + dy] * wb[q];
         if (DIFF) { a1 += tb.B[q * D + dy] * wx[q]; a2 += tb.G[q * D + dy] * wy[q]; }
      }
      sP1[qx2 + PSY * dy + PSZ * dz2] = a2;
      if (DIFF) { sP0[qx2 + PSY * dy + PSZ * dz2] = a1; }
   }
}
template <int P, bool DIFF, bool ATOMIC, int DLO, int DHI>
__device__ __forceinline__ void grp_b3(const GroupTables &tb, const double *sP0, const double *sP1, int pb1, int t1, const int32_t *g, double *y, int64_t e)
{
   constexpr int D = P + 1, Q = P + 2, ND = D * D * D;
   double a1[Q], a2[Q];
   #pragma unroll
   for (int q = 0; q < Q; q++) { a2[q] = sP1[q + pb1]; if (DIFF) { a1[q] = sP0[q + pb1]; } }
   // all results first (independent accumulation chains), then the scatter: a red.add inside the dx loop is a
   // compiler barrier (asm volatile, "memory") and serialised the chains -- 15 % of the stall samples at p = 5, 6
   constexpr bool SPLIT = P >= CDM_G_SPLITP;                          // p = 4 runs at its 128-register cap: one accumulator per dx
   double yb[D], yg[SPLIT ? D : 1];
   #pragma unroll
   for (int dx = DLO; dx < DHI; dx++)
   {
      double sb = 0.0, sg = 0.0;
      #pragma unroll
      for (int q = 0; q < Q; q++)
      {
         sb += tb.B[q * D + dx] * a2[q];
         if (DIFF) { if (SPLIT) { sg += tb.G[q * D + dx] * a1[q]; } else { sb += tb.G[q * D + dx] * a1[q]; } }
      }
      yb[dx] = sb;
      if (SPLIT) { yg[dx] = sg; }
   }
   #pragma unroll
   for (int dx = DLO; dx < DHI; dx++)
   {
      const double a = SPLIT ? yb[dx] + yg[dx] : yb[dx];
      // (plain stores for the dofs inside the element, which belong to it alone, were measured: 2 points SLOWER than red.add)
      if (ATOMIC) { if (g[dx] >= 0) { g_red_add(y + g[dx], a); } }
      else { y[e * ND + D * t1 + dx] = a; }
   }
}

template <int P, bool DIFF, bool CONV, bool MASS, bool ATOMIC>
__global__ void __launch_bounds__(GroupCfg<P>::THREADS, GroupCfg<P>::MINB)
k_apply3d_group(const GroupTables tb, const int64_t ne, const int32_t *__restrict__ gmap,
                const double *__restrict__ x, const double *__restrict__ Dg, const int slab,
                double *__restrict__ y)
{
   using C = GroupCfg<P>;
   constexpr int D = C::D, Q = C::Q, T = C::T, ND = C::ND, Q2 = Q * Q;
   constexpr int RS = C::RS, PS = C::PS, RSTR = C::RSTR;
   constexpr bool GRAD = DIFF || CONV, ALIAS = C::ALIAS;
   // p=4 with E-vector output: the separate stages spill ~500 B at the 128-register cap, the fused form does not
   constexpr bool FUSEZ = C::FUSEZ || (P == 4 && !ATOMIC);
   extern __shared__ __align__(128) unsigned char smraw[];
   const int group_doubles = (Q * slab + C::XS + 1) & ~1;
   // ---- which group am I, which thread of the group
   int gib, t;                                               // group in block, thread in group
   if (C::WPG == 1)
   {
      const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
      const int sub = lane / T;
      gib = (sub < C::EPW) ? wib * C::EPW + sub : -1;
      t = lane - sub * T;
   }
   else if (C::TRIO)
   {
      const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
      if (wib < 2) { gib = wib; t = lane; }
      else { gib = (lane < 8) ? (lane >> 2) : -1; t = 32 + (lane & 3); }
   }
   else if (C::PENTA)
   {
      const int tid = threadIdx.x;
      if (tid < 96) { gib = tid >> 5; t = tid & 31; }
      else if (tid < 144) { gib = (tid - 96) >> 4; t = 32 + (tid & 15); }
      else { gib = (tid < 147) ? tid - 144 : -1; t = 48; }
   }
   else { gib = threadIdx.x >> 6; t = threadIdx.x & 63; }
   const bool member = gib >= 0 && t < T;
   const int gsafe = gib >= 0 ? gib : 0;
   double *gbase = reinterpret_cast<double *>(smraw) + gsafe * group_doubles;
   uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<double *>(smraw) + C::GPB * group_doubles) + gsafe * Q;
   double *ring = gbase, *sR0 = ring + Q * slab, *sR1 = sR0 + RS, *sR2 = sR1 + RS, *sP0 = (ALIAS || C::INPLACE) ? sR0 : sR2 + RS, *sP1 = C::INPLACE ? sR1 : sP0 + PS;
   // BAL (two-warp groups with D^2, Q D <= 32, i.e. p = 4): the x / y exchange stages, which need only D^2 = 25 or
   // Q D = 30 threads, are run by BOTH warps, each producing one half of the stage's outputs for every line; without
   // it the second warp idles through four of the seven stages and `barrier` is the top stall of the kernel
   constexpr bool BAL = C::WPG == 2 && Q * D <= 32 && CDM_G_BALANCED;
   static_assert(!(BAL && C::PACKED), "the two-warp x / y stages and the packed blocks exclude each other");
   constexpr int QH = Q / 2, DH = (D + 1) / 2;
   const int hw = BAL ? (t >> 5) : 0;                                  // warp of the group (warp-uniform)
   const int tl = BAL ? (t & 31) : t;
   const int hh = t >> 4, hr = t & 15;                                 // half-warp of the group, lane in it (lane maps)
   const bool l1 = BAL ? (gib >= 0 && tl < D * D)
                       : (C::L1MAP == 1 ? (gib >= 0 && hr < 2 * D && 2 * hh + hr / D < D)
                          : (C::L1MAP == 2 ? (gib >= 0 && hr < 2 * D && 2 * hh + hr / D < D) : (member && t < D * D)));
   const bool l2 = BAL ? (gib >= 0 && tl < Q * D)
                       : (C::L2MAP == 1 ? (gib >= 0 && hr < 2 * Q && 2 * hh + hr / Q < D) : (member && t < Q * D));
   const int t3 = member ? t : 0;                                      // L3 role: (qx,qy) = t
   const int qx2 = !l2 ? 0 : (C::L2MAP == 1 ? hr % Q : tl / D);        // L2 role
   const int dz2 = !l2 ? 0 : (C::L2MAP == 1 ? 2 * hh + hr / Q : tl % D);
   // L1 role: x-line (dy,dz), its index t1 = dy + D dz in the element, its row pb1 in the P layout
   const int dy1 = !l1 ? 0 : (C::L1MAP == 1 ? hr % D : (C::L1MAP == 2 ? 2 * hh + hr / D : tl % D));
   const int dz1 = !l1 ? 0 : (C::L1MAP == 1 ? 2 * hh + hr / D : (C::L1MAP == 2 ? hr % D : tl / D));
   const int t1 = C::INPLACE ? dy1 + D * dz1 : (l1 ? tl : 0);
   const int pb1 = C::INPLACE ? C::PSY * dy1 + C::PSZ * dz1 : C::PST * t1;
   auto gsync = [&]()
   {
      if (C::WPG == 1) { __syncwarp(); }
#ifndef CDM_G_REGBAR
      else if (C::GPB == 1 || C::PACKED) { __syncthreads(); }  // the block is the group: barrier 0, immediate operand (a register
                                                             // barrier id showed up as 4-12 % branch_resolving stalls)
#endif
      else { asm volatile("bar.sync %0, 64;" ::"r"(gsafe + 1) : "memory"); }
   };

   if (member && t == 0)
   {
      for (int q = 0; q < Q; q++) { g_mbar_init(&bars[q], 1); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
   }
   gsync();
   // Element indices: 32-bit where measured faster (the launcher refuses more than 2e9 elements).  In the packed layouts
   // gg, e, en are per-thread values; as 64-bit integers they cost 7-29 registers.  p=4 with red.add output: 92.4 -> 95.3 %
   // of the HBM roofline; p=6 loses 6 points and p=4 with E-vector output 18 % (a different ptxas schedule), p=5 0.8.
   constexpr bool IDX32 = (C::TRIO && ATOMIC) || C::PENTA || (P == 5 && CDM_G5_IDX32);
   using eidx = typename std::conditional<IDX32, int, int64_t>::type;
   const eidx ngroups = (eidx)gridDim.x * C::GPB;
   const eidx gg = (eidx)blockIdx.x * C::GPB + gsafe;
   // every group of a warp runs the same number of rounds (the lowest group of the block needs the most)
   const eidx g0 = (eidx)blockIdx.x * C::GPB;
   const eidx nel = (eidx)ne;
   const eidx rounds = (g0 < nel) ? (nel - g0 + ngroups - 1) / ngroups : 0;
   const uint32_t slab_bytes = (uint32_t)slab * 8u;

   // gather pipeline: indices two elements ahead, values one element ahead
   int32_t pg[D], pgn[D];
   double px[D];
   #pragma unroll
   for (int i = 0; i < D; i++) { pg[i] = -1; pgn[i] = -1; px[i] = 0.0; }
   if (gib >= 0 && gg < nel)
   {
      if (member && t == 0)
      {
         for (int q = 0; q < Q; q++)
         {
            g_mbar_expect_tx(&bars[q], slab_bytes);
            g_bulk_g2s(ring + q * slab, Dg + ((int64_t)gg * Q + q) * slab, slab_bytes, &bars[q]);
         }
      }
      if (l1)
      {
         #pragma unroll
         for (int i = 0; i < D; i++) { pg[i] = __ldg(gmap + (int64_t)gg * ND + D * t1 + i); }
         if (gg + ngroups < nel)
         {
            #pragma unroll
            for (int i = 0; i < D; i++) { pgn[i] = __ldg(gmap + (int64_t)(gg + ngroups) * ND + D * t1 + i); }
         }
         #pragma unroll
         for (int i = 0; i < D; i++) { px[i] = (pg[i] >= 0) ? __ldg(x + pg[i]) : 0.0; }
      }
   }

   uint32_t parity = 0;
   for (eidx r = 0; r < rounds; r++, parity ^= 1u)
   {
      const eidx e = gg + r * ngroups;
      const bool valid = gib >= 0 && e < nel;
      const eidx en = e + ngroups;
      const bool more = gib >= 0 && en < nel;
      int32_t g[D];
      #pragma unroll
      for (int i = 0; i < D; i++) { g[i] = pg[i]; }
      // ---- F1 (L1 threads): x contraction of the own x-line with B and G
      if (l1 && valid)
      {
         if (!BAL) { grp_f1<P, GRAD, 0, Q>(tb, px, sP0, sP1, pb1); }
         else if (hw == 0) { grp_f1<P, GRAD, 0, QH>(tb, px, sP0, sP1, pb1); }
         else { grp_f1<P, GRAD, QH, Q>(tb, px, sP0, sP1, pb1); }
      }
      if (l1 && more)
      {
         #pragma unroll
         for (int i = 0; i < D; i++) { pg[i] = pgn[i]; }
         if (en + ngroups < nel)
         {
            #pragma unroll
            for (int i = 0; i < D; i++) { pgn[i] = __ldg(gmap + (int64_t)(en + ngroups) * ND + D * t1 + i); }
         }
         #pragma unroll
         for (int i = 0; i < D; i++) { px[i] = (pg[i] >= 0) ? __ldg(x + pg[i]) : 0.0; }
      }
      gsync();
      // ---- F2 (L2 threads): y contraction -> (B B), (G B), (B G)
      {
         double tB[D], tG[D];
         if (l2 && valid) { grp_f2_load<P, GRAD>(sP0, sP1, qx2, dz2, tB, tG); }
         if (ALIAS) { gsync(); }                             // R overwrites P: every input is in registers first
         if (l2 && valid)
         {
            if (!BAL) { grp_f2<P, GRAD, 0, Q>(tb, tB, tG, sR0, sR1, sR2, qx2, dz2); }
            else if (hw == 0) { grp_f2<P, GRAD, 0, QH>(tb, tB, tG, sR0, sR1, sR2, qx2, dz2); }
            else { grp_f2<P, GRAD, QH, Q>(tb, tB, tG, sR0, sR1, sR2, qx2, dz2); }
         }
      }
      gsync();
      if (FUSEZ)
      {
         // ---- F3 + D + B1 fused per quadrature level (L3 threads): the level's u, grad u are formed from the D z-inputs,
         // multiplied by the level's D values and accumulated straight into the 3 D transposed-z sums, so the 4 Q point
         // values never exist together (6 D + ~10 instead of 3 D + 4 Q live doubles)
         double vbb[D], vgb[D], vbg[D], awx[D], awy[D], awb[D];
         #pragma unroll
         for (int dz = 0; dz < D; dz++)
         {
            vbb[dz] = sR0[t3 + RSTR * dz];
            if (GRAD) { vgb[dz] = sR1[t3 + RSTR * dz]; vbg[dz] = sR2[t3 + RSTR * dz]; }
            awx[dz] = 0.0; awy[dz] = 0.0; awb[dz] = 0.0;
         }
         if (valid)
         {
            if (C::WAITALL)
            {
               #pragma unroll
               for (int qz = 0; qz < Q; qz++) { g_mbar_wait(&bars[qz], parity); }
               if (C::PACKED) { __syncwarp(__activemask()); }   // a helper warp waited on the barriers of several elements: reconverge
            }
            #pragma unroll
            for (int qz = 0; qz < Q; qz++)
            {
               double u0 = 0.0, ux0 = 0.0, uy0 = 0.0, uz0 = 0.0;
               #pragma unroll
               for (int dz = 0; dz < D; dz++)
               {
                  u0 += tb.B[qz * D + dz] * vbb[dz];
                  if (GRAD) { ux0 += tb.B[qz * D + dz] * vgb[dz]; uy0 += tb.B[qz * D + dz] * vbg[dz]; uz0 += tb.G[qz * D + dz] * vbb[dz]; }
               }
               if (!C::WAITALL) { g_mbar_wait(&bars[qz], parity); }
               const double *dp = ring + qz * slab + t3;
               double fx = 0.0, fy = 0.0, fz = 0.0, sv = 0.0;
               int c = 0;
               if (DIFF)
               {
                  const double d0 = dp[0], d1 = dp[Q2], d2 = dp[2 * Q2], d3 = dp[3 * Q2], d4 = dp[4 * Q2], d5 = dp[5 * Q2];
                  fx = d0 * ux0 + d1 * uy0 + d2 * uz0;
                  fy = d1 * ux0 + d3 * uy0 + d4 * uz0;
                  fz = d2 * ux0 + d4 * uy0 + d5 * uz0;
                  c = 6;
               }
               if (CONV) { sv = dp[c * Q2] * ux0 + dp[(c + 1) * Q2] * uy0 + dp[(c + 2) * Q2] * uz0; c += 3; }
               if (MASS) { sv += dp[c * Q2] * u0; }
               #pragma unroll
               for (int dz = 0; dz < D; dz++)
               {
                  awb[dz] += tb.B[qz * D + dz] * sv;
                  if (DIFF) { awx[dz] += tb.B[qz * D + dz] * fx; awy[dz] += tb.B[qz * D + dz] * fy; awb[dz] += tb.G[qz * D + dz] * fz; }
               }
            }
         }
         gsync();                                            // D tile and the R buffers are consumed
         if (member && t == 0 && more)
         {
            for (int q = 0; q < Q; q++)
            {
               g_mbar_expect_tx(&bars[q], slab_bytes);
               g_bulk_g2s(ring + q * slab, Dg + ((int64_t)en * Q + q) * slab, slab_bytes, &bars[q]);
            }
         }
         if (member && valid)
         {
            #pragma unroll
            for (int dz = 0; dz < D; dz++)
            {
               sR2[t3 + RSTR * dz] = awb[dz];
               if (DIFF) { sR0[t3 + RSTR * dz] = awx[dz]; sR1[t3 + RSTR * dz] = awy[dz]; }
            }
         }
      }
      else
      {
         // ---- F3 (L3 threads): z contraction in registers
         double u[Q], ux[Q], uy[Q], uz[Q];
         {
            double vbb[D], vgb[D], vbg[D];
            #pragma unroll
            for (int dz = 0; dz < D; dz++)
            {
               vbb[dz] = sR0[t3 + RSTR * dz];
               if (GRAD) { vgb[dz] = sR1[t3 + RSTR * dz]; vbg[dz] = sR2[t3 + RSTR * dz]; }
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
         // ---- point-wise D at the thread's Q quadrature points (registers only)
         if (valid)
         {
            if (C::WAITALL)
            {
               #pragma unroll
               for (int qz = 0; qz < Q; qz++) { g_mbar_wait(&bars[qz], parity); }
               if (C::PACKED) { __syncwarp(__activemask()); }   // a helper warp waited on the barriers of several elements: reconverge
            }
            #pragma unroll
            for (int qz = 0; qz < Q; qz++)
            {
               if (!C::WAITALL) { g_mbar_wait(&bars[qz], parity); }
               const double *dp = ring + qz * slab + t3;
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
         }
         gsync();                                               // D tile and the R buffers are consumed
         if (member && t == 0 && more)
         {
            for (int q = 0; q < Q; q++)
            {
               g_mbar_expect_tx(&bars[q], slab_bytes);
               g_bulk_g2s(ring + q * slab, Dg + ((int64_t)en * Q + q) * slab, slab_bytes, &bars[q]);
            }
         }
         // ---- B1 (L3 threads): transposed z contraction in registers
         if (member && valid)
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
               sR2[t3 + RSTR * dz] = wb;
               if (DIFF) { sR0[t3 + RSTR * dz] = wx; sR1[t3 + RSTR * dz] = wy; }
            }
         }
      }
      gsync();
      // ---- B2 (L2 threads): transposed y contraction
      {
         double wx[Q], wy[Q], wb[Q];
         if (l2 && valid) { grp_b2_load<P, DIFF>(sR0, sR1, sR2, qx2, dz2, wx, wy, wb); }
         if (ALIAS) { gsync(); }                             // P overwrites R
         if (l2 && valid)
         {
            if (!BAL) { grp_b2<P, DIFF, 0, D>(tb, wx, wy, wb, sP0, sP1, qx2, dz2); }
            else if (hw == 0) { grp_b2<P, DIFF, 0, DH>(tb, wx, wy, wb, sP0, sP1, qx2, dz2); }
            else { grp_b2<P, DIFF, DH, D>(tb, wx, wy, wb, sP0, sP1, qx2, dz2); }
         }
      }
      gsync();
      // ---- B3 (L1 threads): transposed x contraction of the own x-line, scatter
      if (l1 && valid)
      {
         if (!BAL) { grp_b3<P, DIFF, ATOMIC, 0, D>(tb, sP0, sP1, pb1, t1, g, y, e); }
         else if (hw == 0) { grp_b3<P, DIFF, ATOMIC, 0, DH>(tb, sP0, sP1, pb1, t1, g, y, e); }
         else { grp_b3<P, DIFF, ATOMIC, DH, D>(tb, sP0, sP1, pb1, t1, g, y, e); }
      }
      gsync();                                               // P buffers are rewritten by the next round's F1
   }
}

template <int P, bool DIFF, bool CONV, bool MASS, bool ATOMIC>
int launch_group(cdm_op *op, const GroupTables &tb, const int32_t *gmap, const double *xL, double *out)
{
   using C = GroupCfg<P>;
   cdm_space *sp = op->sp;
   cdm_ctx *ctx = sp->ctx;
   auto kern = k_apply3d_group<P, DIFF, CONV, MASS, ATOMIC>;
   const int group_doubles = (C::Q * op->slab + C::XS + 1) & ~1;
   const size_t smem = (size_t)C::GPB * group_doubles * sizeof(double) + (size_t)C::GPB * C::Q * sizeof(uint64_t);
   int blocks_per_sm = 0;
   { const int rc = cdm_kernel_cfg(ctx, (const void *)kern, C::THREADS, smem, "k_apply3d_group", &blocks_per_sm); if (rc) { return rc; } }
   // element range [e0, e1): the kernel sees a shifted view of the per-element arrays
   const int64_t e0 = op->range_on ? op->e_begin : 0, e1 = op->range_on ? op->e_end : sp->ne;
   const int64_t n = e1 - e0;
   if (n <= 0) { return CDM_OK; }
   if (n > 2000000000LL) { return 1; }                       // 32-bit element indices in the kernel: the caller falls back
   int64_t grid = (int64_t)ctx->sm_count * blocks_per_sm;
   const int64_t need = (n + C::GPB - 1) / C::GPB;
   if (grid > need) { grid = need; }
   if (op->grid_cap > 0 && grid > op->grid_cap) { grid = op->grid_cap; }
   if (ctx->time_main) { cudaEventRecord(ctx->evk0, ctx->stream); }
   kern<<<(unsigned)grid, C::THREADS, smem, ctx->stream>>>(tb, n, gmap + e0 * C::ND, xL,
                                                           op->D_dev + e0 * C::Q * (int64_t)op->slab, op->slab,
                                                           ATOMIC ? out : out + e0 * C::ND);
   if (ctx->time_main) { cudaEventRecord(ctx->evk1, ctx->stream); }
   ctx->launches++;
   CDM_CUDA(ctx, cudaGetLastError());
   return CDM_OK;
}

template <int P>
int dispatch_order(cdm_op *op, const GroupTables &tb, const int32_t *gmap, const double *xL, double *out, bool atomic)
{
#define GRP_ONE(DF, CV, MS) (atomic ? launch_group<P, DF, CV, MS, true>(op, tb, gmap, xL, out) \
                                    : launch_group<P, DF, CV, MS, false>(op, tb, gmap, xL, out))
   if (op->has_diff && op->has_conv && op->has_mass) { return GRP_ONE(true, true, true); }
   if (op->has_diff && !op->has_conv && op->has_mass) { return GRP_ONE(true, false, true); }
   if (!op->has_diff && !op->has_conv && op->has_mass) { return GRP_ONE(false, false, true); }
   if (op->has_diff && !op->has_conv && !op->has_mass) { return GRP_ONE(true, false, false); }
#undef GRP_ONE
   return 1;
}

__global__ void __launch_bounds__(256)
k_restrict_transpose_g(int64_t ndof, const int32_t *__restrict__ offsets, const int32_t *__restrict__ indices,
                       const double *__restrict__ yE, double *__restrict__ yL)
{
   const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (g >= ndof) { return; }
   double s = 0.0;
   for (int32_t j = offsets[g]; j < offsets[g + 1]; j++) { s += yE[indices[j]]; }
   yL[g] = s;
}
}  // namespace

// returns 1 when this operator is not covered (caller falls back to the generic kernel)
int cdm_k_apply_group(cdm_op *op, const int32_t *gmap, const double *xL, double *yL)
{
   cdm_space *sp = op->sp;
   cdm_ctx *ctx = sp->ctx;
   if (sp->dim != 3 || sp->p < 1 || sp->p > 6) { return 1; }
   if (op->has_conv && !(op->has_diff && op->has_mass)) { return 1; }
   if (!op->has_diff && !op->has_mass) { return 1; }
   GroupTables tb;
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
      case 1: rc = dispatch_order<1>(op, tb, gmap, xL, out, atomic); break;
      case 2: rc = dispatch_order<2>(op, tb, gmap, xL, out, atomic); break;
      case 3: rc = dispatch_order<3>(op, tb, gmap, xL, out, atomic); break;
      case 4: rc = dispatch_order<4>(op, tb, gmap, xL, out, atomic); break;
      case 5: rc = dispatch_order<5>(op, tb, gmap, xL, out, atomic); break;
      case 6: rc = dispatch_order<6>(op, tb, gmap, xL, out, atomic); break;
   }
   if (rc) { return rc; }
   if (!atomic && !op->e_out)
   {
      const unsigned nb = (unsigned)((sp->ndof + 255) / 256);
      k_restrict_transpose_g<<<nb, 256, 0, ctx->stream>>>(sp->ndof, sp->offsets_dev, sp->indices_dev, op->yE_dev, yL);
      ctx->launches++;
      CDM_CUDA(ctx, cudaGetLastError());
   }
   return CDM_OK;
}
