// kernels_forms.cu -- the steps either side of the solve, on the device:
// linear-form assembly, nodal (boundary) projection and L2 error norms.
//
// Stands behind (reference call sites):
//   linear_convection_diffusion_2D.cpp:341-343  ParLinearForm b; DomainLFIntegrator(f); b.Assemble()
//   diffusion_mms.cpp:433-437                   the same once per time step (time-dependent forcing)
//   linear_convection_diffusion_2D.cpp:347      u.ProjectBdrCoefficient(exact, ess_bdr)
//   linear_convection_diffusion_1D.cpp:545-546  the same per step (time-dependent Dirichlet data)
//   linear_convection_diffusion_2D.cpp:383-392  u.ComputeL2Error(exact, irs), ComputeGlobalLpNorm(2, exact, ...)
// MFEM conventions: DomainLFIntegrator integrates with the rule of order 2p
// (p+1 Gauss-Legendre points per direction); the app's error rules have order max(2, 2p+3) (p+2 points).
// A Coefficient::Eval(T, ip) call site becomes "values at the rule's physical points", which the caller
// gets from cdm_space_rule_coords (host or device array).
//
// All three kernels share the element loop: EPB elements per 256-thread block, TPE threads per element,
// one sum-factorised pass through shared memory, 1-D tables in the constant bank (kernel parameter).
// They are HBM-light (LF: 8*Q^dim B read + D^dim red.add per element; error: D^dim gathers + 8*Q^dim B).
#include "cdm_internal.hpp"
#include "kernels_common.cuh"
#include <cmath>

struct RuleTables
{
   double B[CDM_MAX_Q1D * CDM_MAX_D1D];   // B[q*d1d + d] = l_d(x_q)
   double qx[CDM_MAX_Q1D], qw[CDM_MAX_Q1D];
   int d1d, q1d;
};

template <int DIM>
__device__ __forceinline__ void geom_point(const double *X, double x, double y, double z, double *xp)
{
   if (DIM == 2)
   {
      const double N[4] = {(1 - x) * (1 - y), x * (1 - y), x * y, (1 - x) * y};
      for (int c = 0; c < 2; c++)
      {
         double s = 0.0;
         for (int k = 0; k < 4; k++) { s += N[k] * X[k * 2 + c]; }
         xp[c] = s;
      }
   }
   else
   {
      const double N[8] = {(1 - x) * (1 - y) * (1 - z), x * (1 - y) * (1 - z), x * y * (1 - z), (1 - x) * y * (1 - z),
                           (1 - x) * (1 - y) * z, x * (1 - y) * z, x * y * z, (1 - x) * y * z};
      for (int c = 0; c < 3; c++)
      {
         double s = 0.0;
         for (int k = 0; k < 8; k++) { s += N[k] * X[k * 3 + c]; }
         xp[c] = s;
      }
   }
}

// det of the (bi/tri)linear map, same operation order as k_setup_qdata
template <int DIM>
__device__ __forceinline__ double geom_det(const double *X, double x, double y, double z)
{
   double J[DIM][DIM];
   if (DIM == 2)
   {
      const double dN[4][2] = {{-(1 - y), -(1 - x)}, {(1 - y), -x}, {y, x}, {-y, (1 - x)}};
      for (int r = 0; r < DIM; r++)
         for (int c = 0; c < DIM; c++)
         {
            double s = 0.0;
            for (int k = 0; k < 4; k++) { s += X[k * DIM + r] * dN[k][c]; }
            J[r][c] = s;
         }
      return J[0][0] * J[1][1] - J[0][1] * J[1][0];
   }
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
   const double a00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
   const double a01 = J[2][1] * J[0][2] - J[0][1] * J[2][2];
   const double a02 = J[0][1] * J[1][2] - J[1][1] * J[0][2];
   return J[0][0] * a00 + J[1][0] * a01 + J[2][0] * a02;
}

constexpr int FORM_TPE = 64;     // threads per element
constexpr int FORM_EPB = 4;      // elements per block
constexpr int FORM_THREADS = FORM_TPE * FORM_EPB;
static_assert(FORM_THREADS == CDM_RED_THREADS, "block_sum expects CDM_RED_THREADS threads");

// ------------------------------------------------------------ rule points

template <int DIM>
__global__ void __launch_bounds__(256)
k_rule_coords(RuleTables t, int64_t ne, const double *__restrict__ elem_x, double *__restrict__ xyz)
{
   const int q1d = t.q1d, nq = (DIM == 3) ? q1d * q1d * q1d : q1d * q1d;
   constexpr int NV = (DIM == 3) ? 8 : 4;
   for (int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gid < ne * nq; gid += (int64_t)gridDim.x * blockDim.x)
   {
      const int64_t e = gid / nq;
      const int q = (int)(gid - e * nq);
      double X[NV * DIM], xp[DIM];
      for (int i = 0; i < NV * DIM; i++) { X[i] = elem_x[e * NV * DIM + i]; }
      geom_point<DIM>(X, t.qx[q % q1d], t.qx[(q / q1d) % q1d], (DIM == 3) ? t.qx[q / (q1d * q1d)] : 0.0, xp);
      for (int c = 0; c < DIM; c++) { xyz[gid * DIM + c] = xp[c]; }
   }
}

// ------------------------------------------------------ DomainLFIntegrator
// b[gather[e][i]] += scale * sum_q w_q |J_q| f_q B_i(q): B^T applied direction by direction.

template <int DIM>
__global__ void __launch_bounds__(FORM_THREADS)
k_domain_lf(RuleTables t, int64_t ne, const double *__restrict__ elem_x, const int32_t *__restrict__ gather,
            const double *__restrict__ f_q, double scale, double *__restrict__ b)
{
   extern __shared__ double sm[];
   constexpr int NV = (DIM == 3) ? 8 : 4;
   const int d1d = t.d1d, q1d = t.q1d;
   const int nq = (DIM == 3) ? q1d * q1d * q1d : q1d * q1d;
   const int nd = (DIM == 3) ? d1d * d1d * d1d : d1d * d1d;
   const int n1 = ((DIM == 3) ? q1d * q1d : q1d) * d1d;
   const int n2 = (DIM == 3) ? q1d * d1d * d1d : 0;
   const int per = 24 + nq + n1 + n2;
   const int slice = threadIdx.x / FORM_TPE, tid = threadIdx.x % FORM_TPE;
   double *sX = sm + slice * per, *s0 = sX + 24, *s1 = s0 + nq, *s2 = s1 + n1;
   for (int64_t e0 = (int64_t)blockIdx.x * FORM_EPB; e0 < ne; e0 += (int64_t)gridDim.x * FORM_EPB)
   {
      const int64_t e = e0 + slice;
      const bool act = e < ne;
      __syncthreads();
      if (act) { for (int i = tid; i < NV * DIM; i += FORM_TPE) { sX[i] = elem_x[e * NV * DIM + i]; } }
      __syncthreads();
      if (act)
         for (int q = tid; q < nq; q += FORM_TPE)
         {
            const int qx = q % q1d, qy = (q / q1d) % q1d, qz = (DIM == 3) ? q / (q1d * q1d) : 0;
            const double w = t.qw[qx] * t.qw[qy] * ((DIM == 3) ? t.qw[qz] : 1.0);
            const double det = geom_det<DIM>(sX, t.qx[qx], t.qx[qy], (DIM == 3) ? t.qx[qz] : 0.0);
            s0[q] = scale * w * det * f_q[e * nq + q];
         }
      __syncthreads();
      if (act)
         for (int i = tid; i < n1; i += FORM_TPE)          // x: (r, dx) <- sum_qx B[qx][dx] s0[r][qx]
         {
            const int dx = i % d1d, r = i / d1d;
            double s = 0.0;
            for (int qx = 0; qx < q1d; qx++) { s += t.B[qx * d1d + dx] * s0[r * q1d + qx]; }
            s1[i] = s;
         }
      __syncthreads();
      if (DIM == 2)
      {
         if (act)
            for (int i = tid; i < nd; i += FORM_TPE)       // y: (dy, dx) <- sum_qy B[qy][dy] s1[qy][dx]
            {
               const int dx = i % d1d, dy = i / d1d;
               double s = 0.0;
               for (int qy = 0; qy < q1d; qy++) { s += t.B[qy * d1d + dy] * s1[qy * d1d + dx]; }
               atomicAdd(&b[gather[e * nd + i]], s);
            }
      }
      else
      {
         if (act)
            for (int i = tid; i < n2; i += FORM_TPE)       // y: (qz, dy, dx)
            {
               const int dx = i % d1d, dy = (i / d1d) % d1d, qz = i / (d1d * d1d);
               double s = 0.0;
               for (int qy = 0; qy < q1d; qy++) { s += t.B[qy * d1d + dy] * s1[(qz * q1d + qy) * d1d + dx]; }
               s2[i] = s;
            }
         __syncthreads();
         if (act)
            for (int i = tid; i < nd; i += FORM_TPE)       // z: (dz, dy, dx)
            {
               const int dxy = i % (d1d * d1d), dz = i / (d1d * d1d);
               double s = 0.0;
               for (int qz = 0; qz < q1d; qz++) { s += t.B[qz * d1d + dz] * s2[qz * d1d * d1d + dxy]; }
               atomicAdd(&b[gather[e * nd + i]], s);
            }
      }
   }
}

// ------------------------------------------------------- ComputeL2Error
// partial[block] = sum over the block's elements of sum_q w |J| (u_h(x_q) - uex_q)^2, fixed grid and
// fixed summation order (bit-reproducible like the dot products).

template <int DIM>
__global__ void __launch_bounds__(FORM_THREADS)
k_l2_error(RuleTables t, int64_t ne, const double *__restrict__ elem_x, const int32_t *__restrict__ gather,
           const double *__restrict__ u, const double *__restrict__ uex_q, double *__restrict__ partial)
{
   extern __shared__ double sm[];
   __shared__ double red[8];
   constexpr int NV = (DIM == 3) ? 8 : 4;
   const int d1d = t.d1d, q1d = t.q1d;
   const int nq = (DIM == 3) ? q1d * q1d * q1d : q1d * q1d;
   const int nd = (DIM == 3) ? d1d * d1d * d1d : d1d * d1d;
   const int n1 = ((DIM == 3) ? d1d * d1d : d1d) * q1d;      // (dz, dy, qx)
   const int n2 = (DIM == 3) ? d1d * q1d * q1d : 0;          // (dz, qy, qx)
   const int per = 24 + nd + n1 + n2;
   const int slice = threadIdx.x / FORM_TPE, tid = threadIdx.x % FORM_TPE;
   double *sX = sm + slice * per, *su = sX + 24, *s1 = su + nd, *s2 = s1 + n1;
   double acc = 0.0;
   for (int64_t e0 = (int64_t)blockIdx.x * FORM_EPB; e0 < ne; e0 += (int64_t)gridDim.x * FORM_EPB)
   {
      const int64_t e = e0 + slice;
      const bool act = e < ne;
      __syncthreads();
      if (act)
      {
         for (int i = tid; i < NV * DIM; i += FORM_TPE) { sX[i] = elem_x[e * NV * DIM + i]; }
         if (u) { for (int i = tid; i < nd; i += FORM_TPE) { su[i] = u[gather[e * nd + i]]; } }
      }
      __syncthreads();
      if (act && u)
         for (int i = tid; i < n1; i += FORM_TPE)          // x: (r, qx) <- sum_dx B[qx][dx] u[r][dx]
         {
            const int qx = i % q1d, r = i / q1d;
            double s = 0.0;
            for (int dx = 0; dx < d1d; dx++) { s += t.B[qx * d1d + dx] * su[r * d1d + dx]; }
            s1[i] = s;
         }
      __syncthreads();
      if (DIM == 3)
      {
         if (act && u)
            for (int i = tid; i < n2; i += FORM_TPE)       // y: (dz, qy, qx)
            {
               const int qx = i % q1d, qy = (i / q1d) % q1d, dz = i / (q1d * q1d);
               double s = 0.0;
               for (int dy = 0; dy < d1d; dy++) { s += t.B[qy * d1d + dy] * s1[(dz * d1d + dy) * q1d + qx]; }
               s2[i] = s;
            }
         __syncthreads();
      }
      if (act)
         for (int q = tid; q < nq; q += FORM_TPE)
         {
            const int qx = q % q1d, qy = (q / q1d) % q1d, qz = (DIM == 3) ? q / (q1d * q1d) : 0;
            double uh = 0.0;
            if (u)
            {
               if (DIM == 3) { for (int dz = 0; dz < d1d; dz++) { uh += t.B[qz * d1d + dz] * s2[(dz * q1d + qy) * q1d + qx]; } }
               else { for (int dy = 0; dy < d1d; dy++) { uh += t.B[qy * d1d + dy] * s1[dy * q1d + qx]; } }
            }
            const double w = t.qw[qx] * t.qw[qy] * ((DIM == 3) ? t.qw[qz] : 1.0);
            const double det = geom_det<DIM>(sX, t.qx[qx], t.qx[qy], (DIM == 3) ? t.qx[qz] : 0.0);
            const double d = uh - (uex_q ? uex_q[e * nq + q] : 0.0);
            acc += w * det * d * d;
         }
   }
   const double s = block_sum(acc, red);
   if (threadIdx.x == 0) { partial[blockIdx.x] = s; }
}

__global__ void __launch_bounds__(CDM_RED_THREADS)
k_sum_partials(int n, const double *__restrict__ partial, double *__restrict__ out)
{
   __shared__ double red[8];
   double v = 0.0;
   for (int i = threadIdx.x; i < n; i += CDM_RED_THREADS) { v += partial[i]; }
   const double s = block_sum(v, red);
   if (threadIdx.x == 0) { out[0] = s; }
}

__global__ void __launch_bounds__(256)
k_set_indexed(int64_t n, const int32_t *__restrict__ idx, const double *__restrict__ vals, double *__restrict__ u)
{
   for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) { u[idx[i]] = vals[i]; }
}

// ----------------------------------------------------------------- host side

static bool on_device(const void *p)
{
   cudaPointerAttributes a;
   if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
   return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// device view of a caller array that may live on the host: stream-ordered staging copy when needed
struct Staged
{
   cdm_ctx *c; void *dev = nullptr; bool owned = false;
   explicit Staged(cdm_ctx *ctx) : c(ctx) {}
   int in(const void *p, size_t bytes)
   {
      if (!p || on_device(p)) { dev = const_cast<void *>(p); return CDM_OK; }
      CDM_CUDA(c, cudaMallocAsync(&dev, bytes, c->stream));
      owned = true;
      CDM_CUDA(c, cudaMemcpyAsync(dev, p, bytes, cudaMemcpyHostToDevice, c->stream));
      return CDM_OK;
   }
   ~Staged() { if (owned) { cudaFreeAsync(dev, c->stream); } }
};

static int make_rule(const cdm_space *sp, int q1d, RuleTables *t)
{
   if (q1d < 1 || q1d > CDM_MAX_Q1D) { return cdm_fail(sp->ctx, CDM_EUNSUP, "quadrature rule: points per direction must be 1..8"); }
   memset(t, 0, sizeof(*t));
   double G[CDM_MAX_Q1D * CDM_MAX_D1D], nodes[CDM_MAX_D1D];
   cdm_host_basis(sp->p, q1d, t->B, G, t->qw, nodes, t->qx);
   t->d1d = sp->d1d; t->q1d = q1d;
   return CDM_OK;
}

static int launch_check(cdm_ctx *c, const char *what)
{
   c->launches++;
   cudaError_t e = cudaGetLastError();
   if (e != cudaSuccess) { return cdm_fail(c, CDM_ECUDA, std::string(what) + ": " + cudaGetErrorString(e)); }
   return CDM_OK;
}

static int ensure_work(cdm_space *sp)
{
   if (sp->work_dev) { return CDM_OK; }
   CDM_CUDA(sp->ctx, cudaMalloc(&sp->work_dev, sizeof(double) * (size_t)sp->ndof));
   return CDM_OK;
}

static unsigned form_grid(const cdm_space *sp)
{
   const int64_t need = (sp->ne + FORM_EPB - 1) / FORM_EPB;
   return (unsigned)std::min<int64_t>(need, (int64_t)sp->ctx->sm_count * 8);
}

extern "C" {

int cdm_rule_points(int order) { return order < 0 ? CDM_EINVAL : order / 2 + 1; }

int cdm_space_rule_coords(const cdm_space *sp, int q1d, double *xyz)
{
   if (!sp || !xyz) { return CDM_EINVAL; }
   cdm_ctx *c = sp->ctx;
   CDM_REQUIRE_GPU(c);
   if (sp->geom == 1) { return cdm_simplex_rule_coords(sp, q1d == 0 ? sp->q1d : q1d, xyz); }
   if (q1d == 0) { q1d = sp->q1d; }
   RuleTables t;
   int rc = make_rule(sp, q1d, &t); if (rc) { return rc; }
   const int64_t npts = sp->ne * (int64_t)((sp->dim == 3) ? q1d * q1d * q1d : q1d * q1d);
   const size_t bytes = sizeof(double) * (size_t)npts * sp->dim;
   const bool dev_out = on_device(xyz);
   double *out = xyz;
   if (!dev_out) { CDM_CUDA(c, cudaMallocAsync(&out, bytes, c->stream)); }
   const unsigned nb = (unsigned)std::min<int64_t>((npts + 255) / 256, (int64_t)c->sm_count * 16);
   if (sp->dim == 2) { k_rule_coords<2><<<nb, 256, 0, c->stream>>>(t, sp->ne, sp->elem_x_dev, out); }
   else { k_rule_coords<3><<<nb, 256, 0, c->stream>>>(t, sp->ne, sp->elem_x_dev, out); }
   rc = launch_check(c, "k_rule_coords");
   if (!dev_out)
   {
      if (!rc && cudaMemcpyAsync(xyz, out, bytes, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) { rc = cdm_fail(c, CDM_ECUDA, "cdm_space_rule_coords: copy failed"); }
      cudaFreeAsync(out, c->stream);
      if (cudaStreamSynchronize(c->stream) != cudaSuccess && !rc) { rc = cdm_fail(c, CDM_ECUDA, "cdm_space_rule_coords: sync failed"); }
   }
   return rc;
}

int cdm_domain_lf(cdm_space *sp, int q1d, const double *f_q, double scale, int accumulate, double *b_dev)
{
   if (!sp || !f_q || !b_dev) { return CDM_EINVAL; }
   cdm_ctx *c = sp->ctx;
   CDM_REQUIRE_GPU(c);
   if (q1d == 0) { q1d = sp->p + 1; }                      // DomainLFIntegrator default: order 2p
   if (sp->geom == 1) { return cdm_simplex_domain_lf(sp, q1d, f_q, scale, accumulate, b_dev); }
   RuleTables t;
   int rc = make_rule(sp, q1d, &t); if (rc) { return rc; }
   const int dim = sp->dim, d1d = sp->d1d;
   const int nq = (dim == 3) ? q1d * q1d * q1d : q1d * q1d;
   Staged f(c);
   if ((rc = f.in(f_q, sizeof(double) * (size_t)sp->ne * nq))) { return rc; }
   // every rank of a partitioned space takes part in the exchange, also one that holds no ghosts itself
   // (the lowest rank owns all the dofs it shares)
   const bool part = !sp->peers.empty();
   double *target = b_dev;
   if (part)
   {
      if ((rc = ensure_work(sp))) { return rc; }
      target = sp->work_dev;
      if ((rc = cdm_k_set(c, sp->ndof, 0.0, target))) { return rc; }
   }
   else if (!accumulate) { if ((rc = cdm_k_set(c, sp->ndof, 0.0, target))) { return rc; } }
   const int per = 24 + nq + ((dim == 3) ? q1d * q1d : q1d) * d1d + ((dim == 3) ? q1d * d1d * d1d : 0);
   const size_t smem = sizeof(double) * (size_t)per * FORM_EPB;
   const unsigned nb = form_grid(sp);
   if (dim == 2) { k_domain_lf<2><<<nb, FORM_THREADS, smem, c->stream>>>(t, sp->ne, sp->elem_x_dev, sp->gather_dev, (const double *)f.dev, scale, target); }
   else { k_domain_lf<3><<<nb, FORM_THREADS, smem, c->stream>>>(t, sp->ne, sp->elem_x_dev, sp->gather_dev, (const double *)f.dev, scale, target); }
   if ((rc = launch_check(c, "k_domain_lf"))) { return rc; }
   if (part)
   {
      // ParLinearForm::ParallelAssemble: b_T = P^T b_L
      if ((rc = cdm_halo_PT_space(sp, target))) { return rc; }
      if (accumulate) { rc = cdm_k_axpy(c, sp->ntrue, 1.0, target, b_dev); }
      else { CDM_CUDA(c, cudaMemcpyAsync(b_dev, target, sizeof(double) * (size_t)sp->ntrue, cudaMemcpyDeviceToDevice, c->stream)); }
   }
   return rc;
}

int cdm_l2_error(cdm_space *sp, int q1d, const double *u_dev, const double *uex_q, double *result_host)
{
   if (!sp || !result_host || (!u_dev && !uex_q)) { return CDM_EINVAL; }
   cdm_ctx *c = sp->ctx;
   CDM_REQUIRE_GPU(c);
   if (q1d == 0) { q1d = std::max(2, 2 * sp->p + 3) / 2 + 1; }   // the app's irs: order max(2, 2p+3)
   if (sp->geom == 1) { return cdm_simplex_l2_error(sp, q1d, u_dev, uex_q, result_host); }
   RuleTables t;
   int rc = make_rule(sp, q1d, &t); if (rc) { return rc; }
   const int dim = sp->dim, d1d = sp->d1d, nd = sp->nd;
   const int nq = (dim == 3) ? q1d * q1d * q1d : q1d * q1d;
   Staged ex(c);
   if ((rc = ex.in(uex_q, sizeof(double) * (size_t)sp->ne * nq))) { return rc; }
   const double *uL = u_dev;
   if (u_dev && !sp->peers.empty())
   {
      // u_L = P u_T: ghost values come from their owners
      if ((rc = ensure_work(sp))) { return rc; }
      CDM_CUDA(c, cudaMemcpyAsync(sp->work_dev, u_dev, sizeof(double) * (size_t)sp->ntrue, cudaMemcpyDeviceToDevice, c->stream));
      if ((rc = cdm_halo_P_space(sp, sp->work_dev))) { return rc; }
      uL = sp->work_dev;
   }
   if (!sp->elem_part_dev) { CDM_CUDA(c, cudaMalloc(&sp->elem_part_dev, sizeof(double) * CDM_RED_BLOCKS)); }
   const int per = 24 + nd + ((dim == 3) ? d1d * d1d : d1d) * q1d + ((dim == 3) ? d1d * q1d * q1d : 0);
   const size_t smem = sizeof(double) * (size_t)per * FORM_EPB;
   const unsigned nb = CDM_RED_BLOCKS;                       // fixed launch shape: reproducible sums
   if (dim == 2) { k_l2_error<2><<<nb, FORM_THREADS, smem, c->stream>>>(t, sp->ne, sp->elem_x_dev, sp->gather_dev, uL, (const double *)ex.dev, sp->elem_part_dev); }
   else { k_l2_error<3><<<nb, FORM_THREADS, smem, c->stream>>>(t, sp->ne, sp->elem_x_dev, sp->gather_dev, uL, (const double *)ex.dev, sp->elem_part_dev); }
   if ((rc = launch_check(c, "k_l2_error"))) { return rc; }
   double *out = c->red_dev + (size_t)CDM_RED_MAXK * CDM_RED_BLOCKS;
   k_sum_partials<<<1, CDM_RED_THREADS, 0, c->stream>>>((int)nb, sp->elem_part_dev, out);
   if ((rc = launch_check(c, "k_sum_partials"))) { return rc; }
   if ((rc = cdm_allreduce_sum(c, out, 1))) { return rc; }
   CDM_CUDA(c, cudaMemcpyAsync(c->red_host, out, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
   CDM_CUDA(c, cudaStreamSynchronize(c->stream));
   *result_host = std::sqrt(c->red_host[0]);
   return CDM_OK;
}

int cdm_vec_set_indexed(cdm_ctx *c, int64_t n, const int32_t *idx, const double *vals, double *u_dev)
{
   CDM_REQUIRE_GPU(c);
   if (n < 0 || (n > 0 && (!idx || !vals || !u_dev))) { return CDM_EINVAL; }
   if (n == 0) { return CDM_OK; }
   Staged i(c), v(c);
   int rc;
   if ((rc = i.in(idx, sizeof(int32_t) * (size_t)n))) { return rc; }
   if ((rc = v.in(vals, sizeof(double) * (size_t)n))) { return rc; }
   const unsigned nb = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)c->sm_count * 8);
   k_set_indexed<<<nb, 256, 0, c->stream>>>(n, (const int32_t *)i.dev, (const double *)v.dev, u_dev);
   return launch_check(c, "k_set_indexed");
}

}  // extern "C"
