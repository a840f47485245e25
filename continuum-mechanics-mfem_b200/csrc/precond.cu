// precond.cu -- ILU(0) of the assembled matrix on the device: the preconditioner the reference selects for its
// unstructured / convection-dominated cases,
//     -pc_type bjacobi -sub_ksp_type preonly -sub_pc_type ilu      (Input/petsc_circle.opts:6-8,
//                                                                   Input/petsc_nonlinear.opts:6-8)
// i.e. block-Jacobi with one block per rank and ILU(0) (PETSc's default fill level, natural ordering) inside the block.
// One rank = one block = ILU(0) of the whole matrix FormLinearSystem hands to the solver (essential rows / columns
// eliminated, unit diagonal there).
//
//   symbolic : dependency levels of the rows from the sparsity pattern, on the host (setup, like MFEM's / PETSc's own
//              symbolic phases): row i depends on the rows k < i of its pattern (factorisation, forward solve) and on
//              the rows j > i (backward solve);
//   numeric  : one launch per level, one warp per row, IKJ elimination restricted to the pattern;
//   apply    : z = U^{-1} L^{-1} r, one warp per row (lane-strided products, shuffle reduction).  Two schedules with
//              identical arithmetic:
//              * operator option "ilu_sweep" = 1 (default): ONE launch per sweep.  Rows are issued in level order, a warp
//                polls the result word of each row it needs (pre-filled with a NaN pattern no computation produces; value and
//                flag are one 8-byte word, so no fences) -- the dependency chain then costs one L2 round trip per level
//                instead of a kernel launch (3 850 levels at BASELINE config 2).  Forward progress: a row only waits for rows earlier in the
//                issue order, and thread blocks are dispatched in index order; the spin is bounded and raises an error
//                word instead of hanging the device.
//              * "ilu_sweep" = 0: one launch per level and sweep (no inter-block assumptions; the fallback).
// Either way the result matches the sequential algorithm to round-off, which is what the parity tests check.
#include "cdm_internal.hpp"
#include "kernels_common.cuh"
#include <algorithm>

struct cdm_ilu
{
   int64_t n = 0;
   double *lu_dev = nullptr;             // L (unit diagonal, strictly lower part) and U in the pattern of A
   int64_t *diag_dev = nullptr;          // position of the diagonal entry of every row
   int32_t *fwd_rows_dev = nullptr, *bwd_rows_dev = nullptr;    // rows grouped by level
   std::vector<int64_t> fwd_off, bwd_off;                        // level offsets into the two lists
   double *tmp_dev = nullptr;
   unsigned int *err_dev = nullptr;      // error word of the single-launch sweeps (a row never became available)
};

namespace
{
__device__ __forceinline__ int64_t find_col(const int32_t *__restrict__ colind, int64_t lo, int64_t hi, int32_t col)
{
   // position of `col` in the sorted row [lo, hi), or -1
   int64_t a = lo, b = hi - 1;
   while (a <= b)
   {
      const int64_t mid = (a + b) >> 1;
      const int32_t c = colind[mid];
      if (c == col) { return mid; }
      if (c < col) { a = mid + 1; } else { b = mid - 1; }
   }
   return -1;
}

// the matrix the solver sees: essential rows -> identity, essential columns -> 0
__global__ void __launch_bounds__(256)
k_ilu_copy(int64_t n, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colind, const double *__restrict__ vals,
           const unsigned char *__restrict__ ess, double *__restrict__ lu, int64_t *__restrict__ diag)
{
   const int lane = threadIdx.x & 31;
   const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   if (row >= n) { return; }
   const bool er = ess && ess[row];
   for (int64_t k = rowptr[row] + lane; k < rowptr[row + 1]; k += 32)
   {
      const int32_t c = colind[k];
      double v = vals[k];
      if (er) { v = (c == row) ? 1.0 : 0.0; }
      else if (ess && ess[c]) { v = 0.0; }
      lu[k] = v;
      if (c == row) { diag[row] = k; }
   }
}

// rows of one level: for k < i in the pattern of row i (ascending): l_ik = a_ik / u_kk; a_ij -= l_ik u_kj for j > k in both patterns
__global__ void __launch_bounds__(256)
k_ilu_factor_level(int nrows, const int32_t *__restrict__ rows, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colind,
                   const int64_t *__restrict__ diag, double *__restrict__ lu)
{
   const int lane = threadIdx.x & 31;
   const int w = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
   if (w >= nrows) { return; }
   const int32_t i = rows[w];
   const int64_t r0 = rowptr[i], r1 = rowptr[i + 1], di = diag[i];
   for (int64_t p = r0; p < di; p++)
   {
      const int32_t k = colind[p];
      const int64_t dk = diag[k], k1 = rowptr[k + 1];
      const double lik = lu[p] / lu[dk];
      __syncwarp();
      if (lane == 0) { lu[p] = lik; }
      for (int64_t q = p + 1 + lane; q < r1; q += 32)
      {
         const int64_t s = find_col(colind, dk + 1, k1, colind[q]);
         if (s >= 0) { lu[q] -= lik * lu[s]; }
      }
      __syncwarp();
   }
}

// forward sweep on the rows of one level: y_i = r_i - sum_{k<i} l_ik y_k
__global__ void __launch_bounds__(256)
k_ilu_forward_level(int nrows, const int32_t *__restrict__ rows, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colind,
                    const int64_t *__restrict__ diag, const double *__restrict__ lu, const double *__restrict__ r, double *__restrict__ y)
{
   const int lane = threadIdx.x & 31;
   const int w = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
   if (w >= nrows) { return; }
   const int32_t i = rows[w];
   double s = 0.0;
   for (int64_t p = rowptr[i] + lane; p < diag[i]; p += 32) { s += lu[p] * y[colind[p]]; }
   s = warp_sum(s);
   if (lane == 0) { y[i] = r[i] - s; }
}

// backward sweep: z_i = (y_i - sum_{j>i} u_ij z_j) / u_ii  (in place on y)
__global__ void __launch_bounds__(256)
k_ilu_backward_level(int nrows, const int32_t *__restrict__ rows, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colind,
                     const int64_t *__restrict__ diag, const double *__restrict__ lu, double *__restrict__ z)
{
   const int lane = threadIdx.x & 31;
   const int w = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
   if (w >= nrows) { return; }
   const int32_t i = rows[w];
   const int64_t d = diag[i];
   double s = 0.0;
   for (int64_t p = d + 1 + lane; p < rowptr[i + 1]; p += 32) { s += lu[p] * z[colind[p]]; }
   s = warp_sum(s);
   if (lane == 0) { z[i] = (z[i] - s) / lu[d]; }
}
// ---- single-launch sweeps: the warp of row rows[w] waits for the rows it depends on.  Value and "ready" flag are the
// same 8-byte word: the result vector is pre-filled with a NaN pattern no computation produces, a consumer polls the
// word until it changes (one L2 round trip per dependency level, no fences: an aligned 8-byte store is single-copy
// atomic).  The forward sweep writes y into `tmp` and arms z for the backward sweep; the backward sweep consumes
// tmp[i] and re-arms it for the next application.
constexpr unsigned long long ILU_PENDING = 0xFFF8C0DEC0DEC0DEull;
constexpr unsigned int ILU_SPIN_LIMIT = 1u << 24;        // ~ seconds; a healthy sweep waits micro-seconds

__device__ __forceinline__ unsigned long long ld_volatile_u64(const double *p)
{
   unsigned long long v;
   asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
   return v;
}

// value of row c once it has been published in this sweep (0 and the error word set if it never is)
__device__ __forceinline__ double wait_row(const double *v, unsigned int *err, int32_t c)
{
   unsigned int spins = 0;
   unsigned long long w;
   while ((w = ld_volatile_u64(v + c)) == ILU_PENDING)
   {
      if (++spins > ILU_SPIN_LIMIT) { atomicExch(err, 1u); return 0.0; }
   }
   return __longlong_as_double((long long)w);
}

__global__ void __launch_bounds__(256) k_ilu_arm(int64_t n, double *v)
{
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) { reinterpret_cast<unsigned long long *>(v)[i] = ILU_PENDING; }
}

__global__ void __launch_bounds__(256)
k_ilu_forward_flags(int64_t n, const int32_t *__restrict__ rows, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colind,
                    const int64_t *__restrict__ diag, const double *__restrict__ lu, const double *__restrict__ r, double *y, double *z,
                    unsigned int *err)
{
   const int lane = threadIdx.x & 31;
   const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   if (w >= n) { return; }
   const int32_t i = rows[w];
   double s = 0.0;
   for (int64_t p = rowptr[i] + lane; p < diag[i]; p += 32) { s += lu[p] * wait_row(y, err, colind[p]); }
   s = warp_sum(s);
   if (lane == 0)
   {
      y[i] = r[i] - s;
      reinterpret_cast<unsigned long long *>(z)[i] = ILU_PENDING;       // arm the backward sweep (a later launch)
   }
}

__global__ void __launch_bounds__(256)
k_ilu_backward_flags(int64_t n, const int32_t *__restrict__ rows, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colind,
                     const int64_t *__restrict__ diag, const double *__restrict__ lu, double *y, double *z, unsigned int *err)
{
   const int lane = threadIdx.x & 31;
   const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   if (w >= n) { return; }
   const int32_t i = rows[w];
   const int64_t d = diag[i];
   double s = 0.0;
   for (int64_t p = d + 1 + lane; p < rowptr[i + 1]; p += 32) { s += lu[p] * wait_row(z, err, colind[p]); }
   s = warp_sum(s);
   if (lane == 0)
   {
      z[i] = (y[i] - s) / lu[d];
      reinterpret_cast<unsigned long long *>(y)[i] = ILU_PENDING;       // re-arm the forward sweep of the next application
   }
}
}  // namespace

void cdm_ilu_destroy(cdm_op *op)
{
   cdm_ilu *f = op->ilu;
   if (!f) { return; }
   cudaFree(f->lu_dev); cudaFree(f->diag_dev); cudaFree(f->fwd_rows_dev); cudaFree(f->bwd_rows_dev); cudaFree(f->tmp_dev); cudaFree(f->err_dev);
   delete f;
   op->ilu = nullptr;
}

int cdm_ilu_refactor(cdm_op *op)
{
   cdm_ctx *c = op->sp->ctx;
   cdm_csr *m = op->csr;
   cdm_ilu *f = op->ilu;
   if (!m || !f) { return cdm_fail(c, CDM_EINVAL, "cdm_ilu_refactor: no assembled matrix / factorisation"); }
   const unsigned nb = (unsigned)((m->n * 32 + 255) / 256);
   k_ilu_copy<<<nb, 256, 0, c->stream>>>(m->n, m->rowptr_dev, m->colind_dev, m->vals_dev, op->n_ess > 0 ? m->ess_mark_dev : nullptr, f->lu_dev, f->diag_dev);
   c->launches++;
   for (size_t l = 0; l + 1 < f->fwd_off.size(); l++)
   {
      const int nr = (int)(f->fwd_off[l + 1] - f->fwd_off[l]);
      k_ilu_factor_level<<<(unsigned)(((int64_t)nr * 32 + 255) / 256), 256, 0, c->stream>>>(nr, f->fwd_rows_dev + f->fwd_off[l], m->rowptr_dev, m->colind_dev,
                                                                                          f->diag_dev, f->lu_dev);
      c->launches++;
   }
   CDM_CUDA(c, cudaGetLastError());
   return CDM_OK;
}

int cdm_ilu_setup(cdm_op *op)
{
   cdm_ctx *c = op->sp->ctx;
   cdm_csr *m = op->csr;
   if (!m) { return cdm_fail(c, CDM_EINVAL, "ILU(0) needs the assembled matrix: set option assembly = 1 first"); }
   if (op->ilu) { return CDM_OK; }
   cdm_ilu *f = new cdm_ilu;
   op->ilu = f;
   const int64_t n = m->n;
   f->n = n;
   // dependency levels from the pattern
   std::vector<int32_t> lf(n, 0), lb(n, 0);
   int32_t nf = 0, nbk = 0;
   for (int64_t i = 0; i < n; i++)
   {
      int32_t l = 0;
      for (int64_t p = m->rowptr[i]; p < m->rowptr[i + 1] && m->colind[p] < i; p++) { l = std::max(l, lf[m->colind[p]] + 1); }
      lf[i] = l; nf = std::max(nf, l + 1);
   }
   for (int64_t i = n - 1; i >= 0; i--)
   {
      int32_t l = 0;
      for (int64_t p = m->rowptr[i + 1] - 1; p >= m->rowptr[i] && m->colind[p] > i; p--) { l = std::max(l, lb[m->colind[p]] + 1); }
      lb[i] = l; nbk = std::max(nbk, l + 1);
   }
   auto group = [&](const std::vector<int32_t> &lev, int32_t nl, std::vector<int64_t> &off, std::vector<int32_t> &rows)
   {
      off.assign(nl + 1, 0);
      for (int64_t i = 0; i < n; i++) { off[lev[i] + 1]++; }
      for (int32_t l = 0; l < nl; l++) { off[l + 1] += off[l]; }
      rows.resize(n);
      std::vector<int64_t> cur(off.begin(), off.end() - 1);
      for (int64_t i = 0; i < n; i++) { rows[cur[lev[i]]++] = (int32_t)i; }
   };
   std::vector<int32_t> fr, br;
   group(lf, nf, f->fwd_off, fr);
   group(lb, nbk, f->bwd_off, br);
   auto fail = [&](int rc) { cdm_ilu_destroy(op); return rc; };
   if (cudaMalloc(&f->lu_dev, sizeof(double) * (size_t)m->nnz) != cudaSuccess || cudaMalloc(&f->diag_dev, sizeof(int64_t) * (size_t)n) != cudaSuccess ||
       cudaMalloc(&f->fwd_rows_dev, sizeof(int32_t) * (size_t)n) != cudaSuccess || cudaMalloc(&f->bwd_rows_dev, sizeof(int32_t) * (size_t)n) != cudaSuccess ||
       cudaMalloc(&f->tmp_dev, sizeof(double) * (size_t)n) != cudaSuccess || cudaMalloc(&f->err_dev, sizeof(unsigned int)) != cudaSuccess)
   { cudaGetLastError(); return fail(cdm_fail(c, CDM_ENOMEM, "cdm_ilu_setup: out of device memory")); }
   cudaMemcpyAsync(f->fwd_rows_dev, fr.data(), sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, c->stream);
   cudaMemcpyAsync(f->bwd_rows_dev, br.data(), sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, c->stream);
   cudaMemsetAsync(f->err_dev, 0, sizeof(unsigned int), c->stream);
   k_ilu_arm<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(n, f->tmp_dev);
   int rc = cdm_ilu_refactor(op);
   if (cudaStreamSynchronize(c->stream) != cudaSuccess && !rc) { rc = cdm_fail(c, CDM_ECUDA, "cdm_ilu_setup: device failure"); }
   return rc ? fail(rc) : CDM_OK;
}

int cdm_ilu_apply(cdm_op *op, const double *r, double *z)
{
   cdm_ctx *c = op->sp->ctx;
   cdm_csr *m = op->csr;
   cdm_ilu *f = op->ilu;
   if (!m || !f) { return cdm_fail(c, CDM_EINVAL, "cdm_ilu_apply: no factorisation"); }
   if (op->ilu_sweep == 1)
   {
      const unsigned nb = (unsigned)((f->n * 32 + 255) / 256);
      k_ilu_forward_flags<<<nb, 256, 0, c->stream>>>(f->n, f->fwd_rows_dev, m->rowptr_dev, m->colind_dev, f->diag_dev, f->lu_dev, r, f->tmp_dev, z, f->err_dev);
      k_ilu_backward_flags<<<nb, 256, 0, c->stream>>>(f->n, f->bwd_rows_dev, m->rowptr_dev, m->colind_dev, f->diag_dev, f->lu_dev, f->tmp_dev, z, f->err_dev);
      c->launches += 2;
      CDM_CUDA(c, cudaGetLastError());
      return CDM_OK;
   }
   for (size_t l = 0; l + 1 < f->fwd_off.size(); l++)
   {
      const int nr = (int)(f->fwd_off[l + 1] - f->fwd_off[l]);
      k_ilu_forward_level<<<(unsigned)(((int64_t)nr * 32 + 255) / 256), 256, 0, c->stream>>>(nr, f->fwd_rows_dev + f->fwd_off[l], m->rowptr_dev, m->colind_dev,
                                                                                           f->diag_dev, f->lu_dev, r, z);
   }
   for (size_t l = 0; l + 1 < f->bwd_off.size(); l++)
   {
      const int nr = (int)(f->bwd_off[l + 1] - f->bwd_off[l]);
      k_ilu_backward_level<<<(unsigned)(((int64_t)nr * 32 + 255) / 256), 256, 0, c->stream>>>(nr, f->bwd_rows_dev + f->bwd_off[l], m->rowptr_dev, m->colind_dev,
                                                                                            f->diag_dev, f->lu_dev, z);
   }
   c->launches += (int64_t)(f->fwd_off.size() + f->bwd_off.size()) - 2;
   CDM_CUDA(c, cudaGetLastError());
   return CDM_OK;
}

int cdm_ilu_check(cdm_op *op)
{
   cdm_ctx *c = op->sp->ctx;
   cdm_ilu *f = op->ilu;
   if (!f || op->ilu_sweep != 1) { return CDM_OK; }
   unsigned int err = 0;
   CDM_CUDA(c, cudaMemcpyAsync(&err, f->err_dev, sizeof(err), cudaMemcpyDeviceToHost, c->stream));
   CDM_CUDA(c, cudaStreamSynchronize(c->stream));
   if (err)
   {
      // leave the sweeps usable: clear the word and re-arm the forward buffer
      cudaMemsetAsync(f->err_dev, 0, sizeof(unsigned int), c->stream);
      k_ilu_arm<<<(unsigned)((f->n + 255) / 256), 256, 0, c->stream>>>(f->n, f->tmp_dev);
   }
   if (err) { return cdm_fail(c, CDM_ECUDA, "ILU(0) sweep: a row never became available (set operator option ilu_sweep = 0)"); }
   return CDM_OK;
}

extern "C" {

// number of dependency levels of the two triangular sweeps (diagnostics / tests), after the factorisation exists
int cdm_operator_ilu_levels(cdm_op *op, int *forward, int *backward)
{
   if (!op) { return CDM_EINVAL; }
   CDM_REQUIRE_GPU(op->sp->ctx);
   if (!op->csr) { const int rc = cdm_operator_assemble_csr(op); if (rc) { return rc; } }
   const int rc = cdm_ilu_setup(op);
   if (rc) { return rc; }
   if (forward) { *forward = (int)op->ilu->fwd_off.size() - 1; }
   if (backward) { *backward = (int)op->ilu->bwd_off.size() - 1; }
   return CDM_OK;
}

// z = (LU)^{-1} r with the ILU(0) factors of the matrix the solver sees (PCApply of -pc_type ilu / bjacobi+ilu)
int cdm_operator_ilu_apply(cdm_op *op, const double *r_dev, double *z_dev)
{
   if (!op || !r_dev || !z_dev || r_dev == z_dev) { return CDM_EINVAL; }
   CDM_REQUIRE_GPU(op->sp->ctx);
   if (!op->csr) { const int rc = cdm_operator_assemble_csr(op); if (rc) { return rc; } }
   const int rc = cdm_ilu_setup(op);
   return rc ? rc : cdm_ilu_apply(op, r_dev, z_dev);
}

}  // extern "C"
