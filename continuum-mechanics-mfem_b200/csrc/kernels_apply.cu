// kernels_apply.cu -- quadrature-data setup, generic sum-factorised apply (all
// orders, 2D and 3D), PA diagonal and the restriction transpose.
//
// Stands behind (reference call sites, linear_convection_diffusion_2D.cpp):
//   :335-339  a.Assemble()  -> BilinearFormIntegrator::AssemblePA   (k_setup_qdata)
//   :368-370  inside the Krylov solve: Operator::Mult                (k_apply*)
//   Input/petsc.opts:6  -pc_type jacobi -> AssembleDiagonalPA        (k_diag)
// D-tensor algebra: SURVEY.md Appendix C.4 (MFEM bilininteg_*_pa, upstream).
//
// Quadrature-data layout in HBM (3D): one contiguous tile per element, split
// into Q1D z-slabs; slab = [component][qy][qx] with the active components in
// the order diffusion(11,21,31,22,32,33) convection(1,2,3) mass.  A slab is
// padded to an even number of doubles so every slab starts 16-byte aligned
// (bulk-async copies in kernels_apply_p3.cu rely on it).
#include "cdm_internal.hpp"
#include "kernels_common.cuh"

// ---------------------------------------------------------------- setup (K1)

struct SetupArgs
{
   int dim, q1d, ncomp, slab;
   int has_diff, has_conv, has_mass;
   int kappa_kind, kappa_ncomp, vel_kind, mass_kind;
   double kappa_c[6], vel_c[3], mass_c, alpha;
   const double *kappa_q, *vel_q, *mass_q;     // device, point-major
   double qx[CDM_MAX_Q1D], qw[CDM_MAX_Q1D];
};

constexpr int SETUP_EPB = 8;      // elements per block column of k_setup_qdata

template <int DIM>
__global__ void __launch_bounds__(256)
k_setup_qdata(SetupArgs a, int64_t ne, const double *__restrict__ elem_x, double *__restrict__ Dq)
{
   const int q1d = a.q1d;
   const int q2 = q1d * q1d;
   const int nq = (DIM == 3) ? q2 * q1d : q2;
   // a block owns SETUP_EPB consecutive elements: the (element, point) split is a 32-bit division of a small
   // block-local index instead of a 64-bit division of the global one
   const unsigned l = blockIdx.y * blockDim.x + threadIdx.x;
   if (l >= (unsigned)(SETUP_EPB * nq)) { return; }
   const unsigned el = l / (unsigned)nq;
   const int64_t e = (int64_t)blockIdx.x * SETUP_EPB + el;
   if (e >= ne) { return; }
   const int q = (int)(l - el * (unsigned)nq);
   const int qx = q % q1d, qy = (q / q1d) % q1d, qz = (DIM == 3) ? q / q2 : 0;
   const double x = a.qx[qx], y = a.qx[qy], z = (DIM == 3) ? a.qx[qz] : 0.0;
   const double w = a.qw[qx] * a.qw[qy] * ((DIM == 3) ? a.qw[qz] : 1.0);
   constexpr int NV = (DIM == 3) ? 8 : 4;
   const double *X = elem_x + e * NV * DIM;
   double J[DIM][DIM];
   if (DIM == 2)
   {
      // bilinear map, vertices counter-clockwise from (0,0)
      const double dN[4][2] = {{-(1 - y), -(1 - x)}, {(1 - y), -x}, {y, x}, {-y, (1 - x)}};
      for (int r = 0; r < DIM; r++)
         for (int c = 0; c < DIM; c++)
         {
            double s = 0.0;
            for (int k = 0; k < 4; k++) { s += X[k * DIM + r] * dN[k][c]; }
            J[r][c] = s;
         }
   }
   else
   {
      const double mx = 1 - x, my = 1 - y, mz = 1 - z;
      const double dN[8][3] =
      {
         {-my * mz, -mx * mz, -mx * my}, { my * mz, -x * mz, -x * my},
         { y * mz,   x * mz,  -x * y},   {-y * mz,  mx * mz, -mx * y},
         {-my * z,  -mx * z,   mx * my}, { my * z,  -x * z,   x * my},
         { y * z,    x * z,    x * y},   {-y * z,   mx * z,   mx * y}
      };
      for (int r = 0; r < DIM; r++)
         for (int c = 0; c < DIM; c++)
         {
            double s = 0.0;
            for (int k = 0; k < 8; k++) { s += X[k * DIM + r] * dN[k][c]; }
            J[r][c] = s;
         }
   }
   double A[DIM][DIM], det;
   if (DIM == 2)
   {
      A[0][0] = J[1][1]; A[0][1] = -J[0][1]; A[1][0] = -J[1][0]; A[1][1] = J[0][0];
      det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
   }
   else
   {
      A[0][0] = J[1][1] * J[2][2] - J[1][2] * J[2][1];
      A[0][1] = J[2][1] * J[0][2] - J[0][1] * J[2][2];
      A[0][2] = J[0][1] * J[1][2] - J[1][1] * J[0][2];
      A[1][0] = J[2][0] * J[1][2] - J[1][0] * J[2][2];
      A[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0];
      A[1][2] = J[1][0] * J[0][2] - J[0][0] * J[1][2];
      A[2][0] = J[1][0] * J[2][1] - J[2][0] * J[1][1];
      A[2][1] = J[2][0] * J[0][1] - J[0][0] * J[2][1];
      A[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
      det = J[0][0] * A[0][0] + J[1][0] * A[0][1] + J[2][0] * A[0][2];
   }
   // destination of component c of this point
   const int qxy = qx + q1d * qy;
   double *dst = (DIM == 3) ? Dq + ((e * q1d + qz) * (int64_t)a.slab + qxy)
                            : Dq + (e * (int64_t)a.slab + qxy);
   int c = 0;
   if (a.has_diff)
   {
      constexpr int NS = DIM * (DIM + 1) / 2;
      double M[DIM][DIM];
      const double *kp = (a.kappa_kind == CDM_COEFF_QPT) ? a.kappa_q + (e * nq + q) * a.kappa_ncomp : a.kappa_c;
      if (a.kappa_ncomp == 1)
      {
         for (int r = 0; r < DIM; r++) for (int s = 0; s < DIM; s++) { M[r][s] = (r == s) ? kp[0] : 0.0; }
      }
      else
      {
         // packed symmetric: 2D (11,21,22), 3D (11,21,31,22,32,33)
         if (DIM == 2) { M[0][0] = kp[0]; M[1][0] = M[0][1] = kp[1]; M[1][1] = kp[2]; }
         else
         {
            M[0][0] = kp[0]; M[1][0] = M[0][1] = kp[1]; M[2][0] = M[0][2] = kp[2];
            M[1][1] = kp[3]; M[2][1] = M[1][2] = kp[4]; M[2][2] = kp[5];
         }
      }
      const double wd = w / det;
      // (w/det) A M A^T, lower triangle column by column: (0,0),(1,0),(2,0),(1,1),(2,1),(2,2)
      double AM[DIM][DIM];
      for (int r = 0; r < DIM; r++)
         for (int s = 0; s < DIM; s++)
         {
            double t = 0.0;
            for (int k = 0; k < DIM; k++) { t += A[r][k] * M[k][s]; }
            AM[r][s] = t;
         }
      for (int col = 0; col < DIM; col++)
         for (int row = col; row < DIM; row++)
         {
            double t = 0.0;
            for (int k = 0; k < DIM; k++) { t += AM[row][k] * A[col][k]; }
            dst[(int64_t)c * q2] = wd * t; c++;
         }
      (void)NS;
   }
   if (a.has_conv)
   {
      const double *vp = (a.vel_kind == CDM_COEFF_QPT) ? a.vel_q + (e * nq + q) * DIM : a.vel_c;
      for (int r = 0; r < DIM; r++)
      {
         double t = 0.0;
         for (int k = 0; k < DIM; k++) { t += A[r][k] * vp[k]; }
         dst[(int64_t)c * q2] = a.alpha * w * t; c++;
      }
   }
   if (a.has_mass)
   {
      const double ms = (a.mass_kind == CDM_COEFF_QPT) ? a.mass_q[e * nq + q] : a.mass_c;
      dst[(int64_t)c * q2] = w * ms * det;
   }
}

int cdm_k_setup_qdata(cdm_op *op, const cdm_coeff *kappa, const cdm_coeff *vel, double alpha,
                      const cdm_coeff *mass)
{
   cdm_space *sp = op->sp;
   cdm_ctx *ctx = sp->ctx;
   SetupArgs a;
   memset(&a, 0, sizeof(a));
   a.dim = sp->dim; a.q1d = sp->q1d; a.ncomp = op->ncomp; a.slab = op->slab;
   a.has_diff = op->has_diff; a.has_conv = op->has_conv; a.has_mass = op->has_mass;
   a.alpha = alpha;
   for (int i = 0; i < sp->q1d; i++) { a.qx[i] = sp->qx[i]; a.qw[i] = sp->qw[i]; }
   const int64_t npts = sp->ne * sp->nq;
   // per-point coefficient arrays: device pointers are used in place (no copy); host arrays go through a
   // staging buffer kept on the operator, so a per-step cdm_operator_update (ALE drivers,
   // diffusion_mms_ale.cpp:1017-1023) does not allocate
   bool staged_host = false;
   auto stage = [&](const cdm_coeff *c, int ncomp, double *cst, const double **qptr, int *kind, int slot) -> int
   {
      *kind = c->kind;
      if (c->kind == CDM_COEFF_CONST) { for (int i = 0; i < ncomp; i++) { cst[i] = c->data[i]; } return CDM_OK; }
      cudaPointerAttributes at;
      if (cudaPointerGetAttributes(&at, c->data) == cudaSuccess &&
          (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged)) { *qptr = c->data; return CDM_OK; }
      cudaGetLastError();
      const size_t bytes = (size_t)npts * ncomp * sizeof(double);
      if (op->coef_bytes[slot] < bytes)
      {
         if (op->coef_dev[slot]) { cudaFree(op->coef_dev[slot]); op->coef_dev[slot] = nullptr; op->coef_bytes[slot] = 0; }
         CDM_CUDA(ctx, cudaMalloc(&op->coef_dev[slot], bytes));
         op->coef_bytes[slot] = bytes;
      }
      CDM_CUDA(ctx, cudaMemcpyAsync(op->coef_dev[slot], c->data, bytes, cudaMemcpyHostToDevice, ctx->stream));
      *qptr = op->coef_dev[slot];
      staged_host = true;
      return CDM_OK;
   };
   int rc = CDM_OK;
   if (op->has_diff) { a.kappa_ncomp = kappa->ncomp; rc = stage(kappa, kappa->ncomp, a.kappa_c, &a.kappa_q, &a.kappa_kind, 0); }
   if (rc == CDM_OK && op->has_conv) { rc = stage(vel, sp->dim, a.vel_c, &a.vel_q, &a.vel_kind, 1); }
   if (rc == CDM_OK && op->has_mass) { rc = stage(mass, 1, &a.mass_c, &a.mass_q, &a.mass_kind, 2); }
   if (rc == CDM_OK)
   {
      const int bs = 256;
      const dim3 grid((unsigned)((sp->ne + SETUP_EPB - 1) / SETUP_EPB), (unsigned)((SETUP_EPB * sp->nq + bs - 1) / bs));
      if (sp->dim == 2) { k_setup_qdata<2><<<grid, bs, 0, ctx->stream>>>(a, sp->ne, sp->elem_x_dev, op->D_dev); }
      else { k_setup_qdata<3><<<grid, bs, 0, ctx->stream>>>(a, sp->ne, sp->elem_x_dev, op->D_dev); }
      ctx->launches++;
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) { rc = cdm_fail(ctx, CDM_ECUDA, std::string("k_setup_qdata: ") + cudaGetErrorString(e)); }
   }
   // the caller may reuse (pinned) host arrays as soon as this returns
   if (staged_host) { cudaStreamSynchronize(ctx->stream); }
   return rc;
}

// -------------------------------------------------- generic apply, 3D (K3 v1)
//
// One element per Q1D x Q1D thread slice, NBZ elements per block.  x/y
// contractions through shared memory, z contraction in registers.

template <int D, int Q, int NBZ>
__global__ void __launch_bounds__(Q * Q * NBZ)
k_apply3d_generic(BasisTables bs, int64_t ne, const int32_t *__restrict__ gather,
                  const double *__restrict__ x, const double *__restrict__ Dq, int slab,
                  int has_diff, int has_conv, int has_mass, double *__restrict__ y, int atomic)
{
   constexpr int Q2 = Q * Q, ND = D * D * D;
   const int tx = threadIdx.x, ty = threadIdx.y, tz = threadIdx.z;
   const int64_t e = (int64_t)blockIdx.x * NBZ + tz;
   const bool live = e < ne;
   __shared__ double sB[Q * D], sG[Q * D];
   __shared__ double sm[NBZ][5][D * Q2];
   {
      const int t = tx + Q * (ty + Q * tz);
      for (int i = t; i < Q * D; i += Q2 * NBZ) { sB[i] = bs.B[i]; sG[i] = bs.G[i]; }
   }
   double (*s0) = sm[tz][0], (*s1) = sm[tz][1], (*s2) = sm[tz][2], (*s3) = sm[tz][3], (*s4) = sm[tz][4];
   int32_t gi[D];
   // gather: thread (dx,dy) holds the z-column
   // low orders (Q1D <= 4): prefetch this thread's whole D column (Q1D slabs x ncomp values) into
   // registers, so that the HBM latency overlaps the gather and the x / y contractions
   constexpr bool PRE = (Q <= 4);
   double dr[PRE ? Q : 1][10];
   if (PRE && live)
   {
      const int ncomp = (has_diff ? 6 : 0) + (has_conv ? 3 : 0) + (has_mass ? 1 : 0);
      #pragma unroll
      for (int qz = 0; qz < (PRE ? Q : 1); qz++)
      {
         const double *dp = Dq + ((e * Q + qz) * (int64_t)slab + tx + Q * ty);
         #pragma unroll
         for (int c = 0; c < 10; c++) { dr[qz][c] = (c < ncomp) ? __ldg(dp + c * Q2) : 0.0; }
      }
   }
   if (live && tx < D && ty < D)
   {
      #pragma unroll
      for (int dz = 0; dz < D; dz++)
      {
         const int32_t g = gather[e * ND + tx + D * (ty + D * dz)];
         gi[dz] = g;
         s0[(dz * D + ty) * D + tx] = (g >= 0) ? x[g] : 0.0;
      }
   }
   __syncthreads();
   // x contraction: (qx, dy)
   if (live && ty < D)
   {
      #pragma unroll
      for (int dz = 0; dz < D; dz++)
      {
         double bu = 0.0, gu = 0.0;
         #pragma unroll
         for (int dx = 0; dx < D; dx++)
         {
            const double v = s0[(dz * D + ty) * D + dx];
            bu += sB[tx * D + dx] * v; gu += sG[tx * D + dx] * v;
         }
         s1[(dz * D + ty) * Q + tx] = bu; s2[(dz * D + ty) * Q + tx] = gu;
      }
   }
   __syncthreads();
   // y contraction: (qx, qy) -> registers
   double r_bb[D], r_gb[D], r_bg[D];
   #pragma unroll
   for (int dz = 0; dz < D; dz++)
   {
      double bb = 0.0, gb = 0.0, bg = 0.0;
      #pragma unroll
      for (int dy = 0; dy < D; dy++)
      {
         const double b = sB[ty * D + dy], g = sG[ty * D + dy];
         const double vb = s1[(dz * D + dy) * Q + tx], vg = s2[(dz * D + dy) * Q + tx];
         bb += b * vb; gb += b * vg; bg += g * vb;
      }
      r_bb[dz] = bb; r_gb[dz] = gb; r_bg[dz] = bg;
   }
   // z contraction, point-wise D, transposed z contraction
   double t_x[D], t_y[D], t_b[D];
   #pragma unroll
   for (int dz = 0; dz < D; dz++) { t_x[dz] = t_y[dz] = t_b[dz] = 0.0; }
   const int qxy = tx + Q * ty;
   #pragma unroll
   for (int qz = 0; qz < Q; qz++)
   {
      double u = 0.0, ux = 0.0, uy = 0.0, uz = 0.0;
      #pragma unroll
      for (int dz = 0; dz < D; dz++)
      {
         const double b = bs.B[qz * D + dz], g = bs.G[qz * D + dz];
         u += b * r_bb[dz]; ux += b * r_gb[dz]; uy += b * r_bg[dz]; uz += g * r_bb[dz];
      }
      double fx = 0.0, fy = 0.0, fz = 0.0, s = 0.0;
      if (live)
      {
         if (PRE)
         {
            const double *dv = dr[PRE ? qz : 0];
            if (has_diff)
            {
               fx = dv[0] * ux + dv[1] * uy + dv[2] * uz;
               fy = dv[1] * ux + dv[3] * uy + dv[4] * uz;
               fz = dv[2] * ux + dv[4] * uy + dv[5] * uz;
            }
            if (has_conv)
            {
               const double c0 = has_diff ? dv[6] : dv[0], c1 = has_diff ? dv[7] : dv[1], c2 = has_diff ? dv[8] : dv[2];
               s = c0 * ux + c1 * uy + c2 * uz;
            }
            if (has_mass) { s += (has_diff ? (has_conv ? dv[9] : dv[6]) : (has_conv ? dv[3] : dv[0])) * u; }
         }
         else
         {
            const double *dp = Dq + ((e * Q + qz) * (int64_t)slab + qxy);
            if (has_diff)
            {
               const double d0 = dp[0], d1 = dp[Q2], d2 = dp[2 * Q2], d3 = dp[3 * Q2], d4 = dp[4 * Q2], d5 = dp[5 * Q2];
               fx = d0 * ux + d1 * uy + d2 * uz;
               fy = d1 * ux + d3 * uy + d4 * uz;
               fz = d2 * ux + d4 * uy + d5 * uz;
               dp += 6 * Q2;
            }
            if (has_conv) { s = dp[0] * ux + dp[Q2] * uy + dp[2 * Q2] * uz; dp += 3 * Q2; }
            if (has_mass) { s += dp[0] * u; }
         }
      }
      #pragma unroll
      for (int dz = 0; dz < D; dz++)
      {
         const double b = bs.B[qz * D + dz], g = bs.G[qz * D + dz];
         t_x[dz] += b * fx; t_y[dz] += b * fy; t_b[dz] += g * fz + b * s;
      }
   }
   __syncthreads();
   #pragma unroll
   for (int dz = 0; dz < D; dz++)
   {
      s0[(dz * Q + ty) * Q + tx] = t_x[dz]; s1[(dz * Q + ty) * Q + tx] = t_y[dz]; s2[(dz * Q + ty) * Q + tx] = t_b[dz];
   }
   __syncthreads();
   // transposed y contraction: (qx, dy)
   if (ty < D)
   {
      #pragma unroll
      for (int dz = 0; dz < D; dz++)
      {
         double a = 0.0, b = 0.0;
         #pragma unroll
         for (int qy = 0; qy < Q; qy++)
         {
            const double bq = sB[qy * D + ty], gq = sG[qy * D + ty];
            a += bq * s0[(dz * Q + qy) * Q + tx];
            b += gq * s1[(dz * Q + qy) * Q + tx] + bq * s2[(dz * Q + qy) * Q + tx];
         }
         s3[(dz * D + ty) * Q + tx] = a; s4[(dz * D + ty) * Q + tx] = b;
      }
   }
   __syncthreads();
   // transposed x contraction: (dx, dy), scatter
   if (live && tx < D && ty < D)
   {
      #pragma unroll
      for (int dz = 0; dz < D; dz++)
      {
         double v = 0.0;
         #pragma unroll
         for (int qx = 0; qx < Q; qx++)
         {
            v += sG[qx * D + tx] * s3[(dz * D + ty) * Q + qx] + sB[qx * D + tx] * s4[(dz * D + ty) * Q + qx];
         }
         if (atomic) { if (gi[dz] >= 0) { atomicAdd(&y[gi[dz]], v); } }
         else { y[e * ND + tx + D * (ty + D * dz)] = v; }
      }
   }
}

// -------------------------------------------------- generic apply, 2D (K5)

template <int D, int Q, int NBZ>
__global__ void __launch_bounds__(Q * Q * NBZ)
k_apply2d_generic(BasisTables bs, int64_t ne, const int32_t *__restrict__ gather,
                  const double *__restrict__ x, const double *__restrict__ Dq, int estride,
                  int has_diff, int has_conv, int has_mass, double *__restrict__ y, int atomic)
{
   constexpr int Q2 = Q * Q, ND = D * D;
   const int tx = threadIdx.x, ty = threadIdx.y, tz = threadIdx.z;
   const int64_t e = (int64_t)blockIdx.x * NBZ + tz;
   const bool live = e < ne;
   __shared__ double sB[Q * D], sG[Q * D];
   __shared__ double sm[NBZ][5][Q2];
   {
      const int t = tx + Q * (ty + Q * tz);
      for (int i = t; i < Q * D; i += Q2 * NBZ) { sB[i] = bs.B[i]; sG[i] = bs.G[i]; }
   }
   double *s0 = sm[tz][0], *s1 = sm[tz][1], *s2 = sm[tz][2], *s3 = sm[tz][3], *s4 = sm[tz][4];
   int32_t g = -1;
   if (live && tx < D && ty < D)
   {
      g = gather[e * ND + tx + D * ty];
      s0[ty * D + tx] = (g >= 0) ? x[g] : 0.0;
   }
   __syncthreads();
   if (ty < D)
   {
      double bu = 0.0, gu = 0.0;
      #pragma unroll
      for (int dx = 0; dx < D; dx++) { const double v = s0[ty * D + dx]; bu += sB[tx * D + dx] * v; gu += sG[tx * D + dx] * v; }
      s1[ty * Q + tx] = bu; s2[ty * Q + tx] = gu;
   }
   __syncthreads();
   double u = 0.0, ux = 0.0, uy = 0.0;
   #pragma unroll
   for (int dy = 0; dy < D; dy++)
   {
      const double b = sB[ty * D + dy], gg = sG[ty * D + dy];
      u += b * s1[dy * Q + tx]; ux += b * s2[dy * Q + tx]; uy += gg * s1[dy * Q + tx];
   }
   double fx = 0.0, fy = 0.0, s = 0.0;
   if (live)
   {
      const double *dp = Dq + (e * (int64_t)estride + tx + Q * ty);
      if (has_diff)
      {
         const double d0 = dp[0], d1 = dp[Q2], d2 = dp[2 * Q2];
         fx = d0 * ux + d1 * uy; fy = d1 * ux + d2 * uy; dp += 3 * Q2;
      }
      if (has_conv) { s = dp[0] * ux + dp[Q2] * uy; dp += 2 * Q2; }
      if (has_mass) { s += dp[0] * u; }
   }
   __syncthreads();
   s0[ty * Q + tx] = fx; s1[ty * Q + tx] = fy; s2[ty * Q + tx] = s;
   __syncthreads();
   if (ty < D)
   {
      double a = 0.0, b = 0.0;
      #pragma unroll
      for (int qy = 0; qy < Q; qy++)
      {
         const double bq = sB[qy * D + ty], gq = sG[qy * D + ty];
         a += bq * s0[qy * Q + tx];
         b += gq * s1[qy * Q + tx] + bq * s2[qy * Q + tx];
      }
      s3[ty * Q + tx] = a; s4[ty * Q + tx] = b;
   }
   __syncthreads();
   if (live && tx < D && ty < D)
   {
      double v = 0.0;
      #pragma unroll
      for (int qx = 0; qx < Q; qx++) { v += sG[qx * D + tx] * s3[ty * Q + qx] + sB[qx * D + tx] * s4[ty * Q + qx]; }
      if (atomic) { if (g >= 0) { atomicAdd(&y[g], v); } }
      else { y[e * ND + tx + D * ty] = v; }
   }
}

// -------------------------------------------- restriction transpose (K4)
// yL[g] = sum_{j in [offsets[g], offsets[g+1])} yE[indices[j]]  -- fixed order, no atomics
__global__ void __launch_bounds__(256)
k_restrict_transpose(int64_t ndof, const int32_t *__restrict__ offsets, const int32_t *__restrict__ indices,
                     const double *__restrict__ yE, double *__restrict__ yL)
{
   const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (g >= ndof) { return; }
   const int32_t j0 = offsets[g], j1 = offsets[g + 1];
   double s = 0.0;
   for (int32_t j = j0; j < j1; j++) { s += yE[indices[j]]; }
   yL[g] = s;
}

// -------------------------------------------------------- diagonal (K7)
// dE[e, l] = sum_q [ grad(phi_l)^T D grad(phi_l) + phi_l Dc.grad(phi_l) + Dm phi_l^2 ]
template <int DIM>
__global__ void __launch_bounds__(128)
k_diag_elem(BasisTables bs, int d1d, int q1d, int64_t ne, const double *__restrict__ Dq, int slab,
            int has_diff, int has_conv, int has_mass, double *__restrict__ dE)
{
   const int nd = (DIM == 3) ? d1d * d1d * d1d : d1d * d1d;
   const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (gid >= ne * nd) { return; }
   const int64_t e = gid / nd;
   const int l = (int)(gid - e * nd);
   const int lx = l % d1d, ly = (l / d1d) % d1d, lz = (DIM == 3) ? l / (d1d * d1d) : 0;
   const int q2 = q1d * q1d;
   double acc = 0.0;
   const int nqz = (DIM == 3) ? q1d : 1;
   for (int qz = 0; qz < nqz; qz++)
   {
      const double bz = (DIM == 3) ? bs.B[qz * d1d + lz] : 1.0, gz = (DIM == 3) ? bs.G[qz * d1d + lz] : 0.0;
      const double *base = (DIM == 3) ? Dq + (e * q1d + qz) * (int64_t)slab : Dq + e * (int64_t)slab;
      for (int qy = 0; qy < q1d; qy++)
      {
         const double by = bs.B[qy * d1d + ly], gy = bs.G[qy * d1d + ly];
         for (int qx = 0; qx < q1d; qx++)
         {
            const double bx = bs.B[qx * d1d + lx], gx = bs.G[qx * d1d + lx];
            const double phi = bx * by * bz;
            const double g0 = gx * by * bz, g1 = bx * gy * bz, g2 = bx * by * gz;
            const double *dp = base + qx + q1d * qy;
            if (has_diff)
            {
               if (DIM == 3)
               {
                  acc += dp[0] * g0 * g0 + 2.0 * dp[q2] * g0 * g1 + 2.0 * dp[2 * q2] * g0 * g2
                         + dp[3 * q2] * g1 * g1 + 2.0 * dp[4 * q2] * g1 * g2 + dp[5 * q2] * g2 * g2;
                  dp += 6 * q2;
               }
               else
               {
                  acc += dp[0] * g0 * g0 + 2.0 * dp[q2] * g0 * g1 + dp[2 * q2] * g1 * g1;
                  dp += 3 * q2;
               }
            }
            if (has_conv)
            {
               acc += phi * (dp[0] * g0 + dp[q2] * g1 + ((DIM == 3) ? dp[2 * q2] * g2 : 0.0));
               dp += DIM * q2;
            }
            if (has_mass) { acc += dp[0] * phi * phi; }
         }
      }
   }
   dE[gid] = acc;
}

// Sum-factorised 3D diagonal: because phi_l is a tensor product, every term of the diagonal factors into
// 1-D tables BB = B^2, BG = B G, GG = G^2.  One thread per (element, dx, dy) contracts x and y for each z-slab
// into three partial sums (one per z-factor type), then finishes the D1D values of its z-column.
template <int D, int Q>
__global__ void __launch_bounds__(128)
k_diag3d_sumfact(BasisTables bs, int64_t ne, const double *__restrict__ Dq, int slab,
                 int has_diff, int has_conv, int has_mass, double *__restrict__ dE)
{
   constexpr int Q2 = Q * Q;
   const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (gid >= ne * D * D) { return; }
   const int64_t e = gid / (D * D);
   const int l = (int)(gid - e * (D * D));
   const int dx = l % D, dy = l / D;
   double bbx[Q], bgx[Q], ggx[Q], bby[Q], bgy[Q], ggy[Q];
   #pragma unroll
   for (int q = 0; q < Q; q++)
   {
      const double b = bs.B[q * D + dx], g = bs.G[q * D + dx];
      bbx[q] = b * b; bgx[q] = b * g; ggx[q] = g * g;
      const double b2 = bs.B[q * D + dy], g2 = bs.G[q * D + dy];
      bby[q] = b2 * b2; bgy[q] = b2 * g2; ggy[q] = g2 * g2;
   }
   double out[D];
   #pragma unroll
   for (int dz = 0; dz < D; dz++) { out[dz] = 0.0; }
   // component offsets inside a slab
   const int oc = has_diff ? 6 : 0, om = oc + (has_conv ? 3 : 0);
   for (int qz = 0; qz < Q; qz++)
   {
      const double *base = Dq + (e * Q + qz) * (int64_t)slab;
      double sbb = 0.0, sbg = 0.0, sgg = 0.0;          // partial sums that meet B^2, B G, G^2 in z
      for (int qy = 0; qy < Q; qy++)
      {
         #pragma unroll
         for (int qx = 0; qx < Q; qx++)
         {
            const double *dp = base + qx + Q * qy;
            const double xx = bbx[qx] * bby[qy];
            if (has_diff)
            {
               sbb += dp[0] * ggx[qx] * bby[qy] + 2.0 * dp[Q2] * bgx[qx] * bgy[qy] + dp[3 * Q2] * bbx[qx] * ggy[qy];
               sbg += 2.0 * (dp[2 * Q2] * bgx[qx] * bby[qy] + dp[4 * Q2] * bbx[qx] * bgy[qy]);
               sgg += dp[5 * Q2] * xx;
            }
            if (has_conv)
            {
               sbb += dp[oc * Q2] * bgx[qx] * bby[qy] + dp[(oc + 1) * Q2] * bbx[qx] * bgy[qy];
               sbg += dp[(oc + 2) * Q2] * xx;
            }
            if (has_mass) { sbb += dp[om * Q2] * xx; }
         }
      }
      #pragma unroll
      for (int dz = 0; dz < D; dz++)
      {
         const double b = bs.B[qz * D + dz], g = bs.G[qz * D + dz];
         out[dz] += b * b * sbb + b * g * sbg + g * g * sgg;
      }
   }
   #pragma unroll
   for (int dz = 0; dz < D; dz++) { dE[e * (D * D * D) + dx + D * (dy + D * dz)] = out[dz]; }
}

// ------------------------------------------------------------- launchers

static BasisTables make_tables(const cdm_space *sp)
{
   BasisTables t;
   memset(&t, 0, sizeof(t));
   for (int i = 0; i < sp->q1d * sp->d1d; i++) { t.B[i] = sp->B[i]; t.G[i] = sp->G[i]; }
   return t;
}

static int ensure_yE(cdm_op *op)
{
   if (op->yE_dev) { return CDM_OK; }
   cdm_ctx *ctx = op->sp->ctx;
   CDM_CUDA(ctx, cudaMalloc(&op->yE_dev, sizeof(double) * (size_t)op->sp->ne * op->sp->nd));
   return CDM_OK;
}

int cdm_k_apply_p3(cdm_op *op, const int32_t *gmap, const double *xL, double *yL);   // kernels_apply_p3.cu
int cdm_k_apply_group(cdm_op *op, const int32_t *gmap, const double *xL, double *yL);   // kernels_apply_warp.cu
int cdm_k_apply_sub(cdm_op *op, const int32_t *gmap, const double *xL, double *yL);     // kernels_apply_sub.cu
int cdm_k_apply_2d_thread(cdm_op *op, const int32_t *gmap, const double *xL, double *yL);   // kernels_apply_2d.cu

#define LAUNCH3D(P, NBZ)                                                                              \
   case P: {                                                                                          \
      dim3 blk(P + 2, P + 2, NBZ);                                                                    \
      const unsigned nb = (unsigned)((n_el + NBZ - 1) / NBZ);                                         \
      k_apply3d_generic<P + 1, P + 2, NBZ><<<nb, blk, 0, ctx->stream>>>(                              \
         bt, n_el, gmap + e0 * sp->nd, xL, op->D_dev + e0 * (int64_t)sp->q1d * op->slab, op->slab,    \
         op->has_diff, op->has_conv, op->has_mass, atomic ? out : out + e0 * sp->nd, atomic);         \
   } break
#define LAUNCH2D(P, NBZ)                                                                              \
   case P: {                                                                                          \
      dim3 blk(P + 1, P + 1, NBZ);                                                                    \
      const unsigned nb = (unsigned)((n_el + NBZ - 1) / NBZ);                                         \
      k_apply2d_generic<P + 1, P + 1, NBZ><<<nb, blk, 0, ctx->stream>>>(                              \
         bt, n_el, gmap + e0 * sp->nd, xL, op->D_dev + e0 * (int64_t)op->slab, op->slab,              \
         op->has_diff, op->has_conv, op->has_mass, atomic ? out : out + e0 * sp->nd, atomic);         \
   } break

// true when the kernel cdm_k_apply will pick honours op->range_on (everything except the collocated
// order-3 variants 1 and 2, which always sweep the whole mesh)
bool cdm_k_range_capable(const cdm_op *op)
{
   const cdm_space *sp = op->sp;
   if (op->assembly == 1 && op->csr) { return false; }      // the SpMV always sweeps all rows
   return !(sp->dim == 3 && sp->p == 3 && (op->kernel_variant == 1 || op->kernel_variant == 2));
}

// yL = G^T B^T D B G xL on this rank's L-vector (no halo, no essential fix-up)
int cdm_k_apply(cdm_op *op, const double *xL, double *yL, bool constrained)
{
   cdm_space *sp = op->sp;
   cdm_ctx *ctx = sp->ctx;
   if (op->assembly == 1 && op->csr) { return cdm_k_csr_spmv(op, xL, yL, constrained); }   // the reference's literal path
   const int32_t *gmap = op->gmap_override ? op->gmap_override
                                           : ((constrained && op->gather_c_dev) ? op->gather_c_dev : sp->gather_dev);
   if (sp->dim == 2 && op->kernel_variant == 6)
   {
      const int rc = cdm_k_apply_2d_thread(op, gmap, xL, yL);
      if (rc != 1) { return rc; }          // 1: order / integrator set not instantiated -> generic kernel
   }
   else if (sp->dim == 3 && op->kernel_variant == 4)
   {
      const int rc = cdm_k_apply_group(op, gmap, xL, yL);
      if (rc != 1) { return rc; }
   }
   else if (sp->dim == 3 && op->kernel_variant == 5)
   {
      const int rc = cdm_k_apply_sub(op, gmap, xL, yL);
      if (rc != 1) { return rc; }
   }
   else if (sp->dim == 3 && sp->p == 3 && op->kernel_variant >= 1)
   {
      const int rc = cdm_k_apply_p3(op, gmap, xL, yL);
      if (rc != 1) { return rc; }          // 1: integrator combination not instantiated -> generic kernel
   }
   const int atomic = op->scatter_mode == 1;
   double *out = yL;
   // element range [e0, e0 + n_el) (overlapped multi-GPU schedule, pipelined host apply): like the
   // specialised kernels, a ranged launch accumulates into a y the caller has zeroed
   const int64_t e0 = op->range_on ? op->e_begin : 0, n_el = (op->range_on ? op->e_end : sp->ne) - e0;
   if (n_el <= 0) { return CDM_OK; }
   if (atomic) { if (!op->range_on) { CDM_CUDA(ctx, cudaMemsetAsync(yL, 0, sizeof(double) * (size_t)sp->ndof, ctx->stream)); } }
   else { int rc = ensure_yE(op); if (rc) { return rc; } out = op->e_out ? op->e_out : op->yE_dev; }
   const BasisTables bt = make_tables(sp);
   if (ctx->time_main) { cudaEventRecord(ctx->evk0, ctx->stream); }
   if (sp->dim == 3)
   {
      switch (sp->p)
      {
         LAUNCH3D(1, 16); LAUNCH3D(2, 8); LAUNCH3D(3, 5); LAUNCH3D(4, 4); LAUNCH3D(5, 3); LAUNCH3D(6, 2);
         default: return cdm_fail(ctx, CDM_EUNSUP, "order must be 1..6");
      }
   }
   else
   {
      switch (sp->p)
      {
         LAUNCH2D(1, 32); LAUNCH2D(2, 16); LAUNCH2D(3, 8); LAUNCH2D(4, 8); LAUNCH2D(5, 4); LAUNCH2D(6, 4);
         default: return cdm_fail(ctx, CDM_EUNSUP, "order must be 1..6");
      }
   }
   if (ctx->time_main) { cudaEventRecord(ctx->evk1, ctx->stream); }
   ctx->launches++;
   if (!atomic && !op->e_out)
   {
      const unsigned nb = (unsigned)((sp->ndof + 255) / 256);
      k_restrict_transpose<<<nb, 256, 0, ctx->stream>>>(sp->ndof, sp->offsets_dev, sp->indices_dev, op->yE_dev, yL);
      ctx->launches++;
   }
   CDM_CUDA(ctx, cudaGetLastError());
   return CDM_OK;
}

// element-wise diagonal (E-vector layout [ne][nd]) written to dE
int cdm_k_diag_evec(cdm_op *op, double *dE)
{
   cdm_space *sp = op->sp;
   cdm_ctx *ctx = sp->ctx;
   const BasisTables bt = make_tables(sp);
   const int64_t n = sp->ne * sp->nd;
   const unsigned nb = (unsigned)((n + 127) / 128);
   if (sp->dim == 3)
   {
      const unsigned nb3 = (unsigned)((sp->ne * sp->d1d * sp->d1d + 127) / 128);
#define DIAG3(P) case P: k_diag3d_sumfact<P + 1, P + 2><<<nb3, 128, 0, ctx->stream>>>(                    \
         bt, sp->ne, op->D_dev, op->slab, op->has_diff, op->has_conv, op->has_mass, dE); break
      switch (sp->p)
      {
         DIAG3(1); DIAG3(2); DIAG3(3); DIAG3(4); DIAG3(5); DIAG3(6);
         default:
            k_diag_elem<3><<<nb, 128, 0, ctx->stream>>>(bt, sp->d1d, sp->q1d, sp->ne, op->D_dev, op->slab,
                                                        op->has_diff, op->has_conv, op->has_mass, dE);
      }
#undef DIAG3
   }
   else
      k_diag_elem<2><<<nb, 128, 0, ctx->stream>>>(bt, sp->d1d, sp->q1d, sp->ne, op->D_dev, op->slab,
                                                  op->has_diff, op->has_conv, op->has_mass, dE);
   ctx->launches++;
   CDM_CUDA(ctx, cudaGetLastError());
   return CDM_OK;
}

int cdm_k_diag(cdm_op *op, double *dL)
{
   cdm_space *sp = op->sp;
   cdm_ctx *ctx = sp->ctx;
   int rc = ensure_yE(op); if (rc) { return rc; }
   if ((rc = cdm_k_diag_evec(op, op->yE_dev))) { return rc; }
   const unsigned nb2 = (unsigned)((sp->ndof + 255) / 256);
   k_restrict_transpose<<<nb2, 256, 0, ctx->stream>>>(sp->ndof, sp->offsets_dev, sp->indices_dev, op->yE_dev, dL);
   ctx->launches++;
   CDM_CUDA(ctx, cudaGetLastError());
   return CDM_OK;
}

int cdm_k_ensure_yE(cdm_op *op) { return ensure_yE(op); }

// ElementRestriction::MultTranspose: yL[g] = sum of the E-vector entries that map to g (fixed order)
int cdm_k_restrict_transpose(cdm_space *sp, const double *yE, double *yL)
{
   cdm_ctx *ctx = sp->ctx;
   const unsigned nb = (unsigned)((sp->ndof + 255) / 256);
   k_restrict_transpose<<<nb, 256, 0, ctx->stream>>>(sp->ndof, sp->offsets_dev, sp->indices_dev, yE, yL);
   ctx->launches++;
   CDM_CUDA(ctx, cudaGetLastError());
   return CDM_OK;
}

// host copy of the quadrature data in MFEM layout (tests)
int cdm_k_get_qdata(const cdm_op *op, double *Ddiff, double *Dconv, double *Dmass)
{
   const cdm_space *sp = op->sp;
   cdm_ctx *ctx = sp->ctx;
   std::vector<double> h((size_t)op->D_len);
   CDM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
   CDM_CUDA(ctx, cudaMemcpy(h.data(), op->D_dev, sizeof(double) * h.size(), cudaMemcpyDeviceToHost));
   const int dim = sp->dim, q1d = sp->q1d, q2 = q1d * q1d, nq = sp->nq, nsym = dim * (dim + 1) / 2;
   for (int64_t e = 0; e < sp->ne; e++)
      for (int q = 0; q < nq; q++)
      {
         const int qz = (dim == 3) ? q / q2 : 0, qxy = q % q2;
         const double *src = (dim == 3) ? &h[(size_t)((e * q1d + qz) * (int64_t)op->slab + qxy)]
                                        : &h[(size_t)(e * (int64_t)op->slab + qxy)];
         int c = 0;
         if (op->has_diff) { for (int k = 0; k < nsym; k++, c++) if (Ddiff) { Ddiff[(e * nsym + k) * nq + q] = src[(size_t)c * q2]; } }
         if (op->has_conv) { for (int k = 0; k < dim; k++, c++) if (Dconv) { Dconv[(e * dim + k) * nq + q] = src[(size_t)c * q2]; } }
         if (op->has_mass) { if (Dmass) { Dmass[e * nq + q] = src[(size_t)c * q2]; } }
      }
   return CDM_OK;
}
