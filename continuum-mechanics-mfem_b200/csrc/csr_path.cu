// csr_path.cu -- the reference's literal path on the device: full assembly of the
// Diffusion + Convection + Mass form into a CSR matrix and SpMV as the operator apply.
//
// Stands behind (reference call sites, linear_convection_diffusion_2D.cpp):
//   :339   a.Assemble()                      element matrices -> sparse matrix
//   :351   a.FormLinearSystem(...)           the matrix the solver sees (essential rows / columns eliminated)
//   :368   PetscLinearSolver(A).Mult(B, X)   MatMult inside KSP
// It exists so that the partially-assembled operator can be compared with the algorithm the application
// actually runs, on the same GPU: operator option "assembly" = 1 switches cdm_operator_apply (and therefore
// cdm_gmres / cdm_cg) to the CSR SpMV.  Single rank; quad / hex meshes (where the matrix-free path stays the product)
// and triangle meshes (where it is the only path, as in the reference).
//
// * pattern: on the host from the element restriction (row g = union of the dofs of the elements touching g,
//   sorted) -- bit-exact with the oracle's pattern;
// * values: one warp per matrix row; the element-matrix rows that map to it are computed entry by entry
//   (a_ij = sum_q grad(phi_i).D grad(phi_j) + phi_i Dc.grad(phi_j) + Dm phi_i phi_j from the SAME quadrature data the
//   matrix-free kernels stream, so both paths discretise identically) and added in the fixed order of the
//   ElementRestriction transpose: no atomics, the matrix is bit-reproducible; quads, hexes and triangles;
// * apply: warp per row, essential dofs handled like ConstrainedOperator (row -> identity, column skipped),
//   which is the matrix FormLinearSystem produces with DIAG_ONE.
#include "cdm_internal.hpp"
#include "kernels_common.cuh"
#include <algorithm>


namespace
{
// 1-D factors of basis function i at quadrature point q: value and reference gradient
template <int DIM>
__device__ __forceinline__ void basis_at(const BasisTables &t, int d1d, int q1d, int i, int q, double &phi, double *g)
{
   const int ix = i % d1d, iy = (i / d1d) % d1d, iz = (DIM == 3) ? i / (d1d * d1d) : 0;
   const int qx = q % q1d, qy = (q / q1d) % q1d, qz = (DIM == 3) ? q / (q1d * q1d) : 0;
   const double bx = t.B[qx * d1d + ix], by = t.B[qy * d1d + iy], bz = (DIM == 3) ? t.B[qz * d1d + iz] : 1.0;
   const double gx = t.G[qx * d1d + ix], gy = t.G[qy * d1d + iy];
   phi = bx * by * bz;
   g[0] = gx * by * bz;
   g[1] = bx * gy * bz;
   if (DIM == 3) { g[2] = bx * by * t.G[qz * d1d + iz]; }
}

// One entry of an element matrix: a_ij = sum_q grad(phi_i) . D grad(phi_j) + phi_i Dc . grad(phi_j) + Dm phi_i phi_j.
// MODE 2 / 3: tensor-product quads / hexes (1-D tables in the constant bank); MODE 1: triangles (dense tables B[q][i],
// G[c][q][i] in global memory, quadrature data per element [component][nq]).
template <int MODE>
__device__ __forceinline__ double elem_entry(const BasisTables &t, int d1d, int q1d, int nd, int nq, const double *__restrict__ sB,
                                             const double *__restrict__ sG, const double *__restrict__ Dq, int slab, int64_t e, int i, int j,
                                             int has_diff, int has_conv, int has_mass)
{
   constexpr int DIM = (MODE == 3) ? 3 : 2;
   const int q2 = (MODE == 1) ? nq : q1d * q1d;
   const int oc = has_diff ? DIM * (DIM + 1) / 2 : 0, om = oc + (has_conv ? DIM : 0);
   double a = 0.0;
   for (int q = 0; q < nq; q++)
   {
      double pi, pj, gi[3], gj[3];
      if (MODE == 1)
      {
         pi = sB[(size_t)q * nd + i]; pj = sB[(size_t)q * nd + j];
         gi[0] = sG[(size_t)q * nd + i]; gi[1] = sG[((size_t)nq + q) * nd + i];
         gj[0] = sG[(size_t)q * nd + j]; gj[1] = sG[((size_t)nq + q) * nd + j];
      }
      else
      {
         basis_at<DIM>(t, d1d, q1d, i, q, pi, gi);
         basis_at<DIM>(t, d1d, q1d, j, q, pj, gj);
      }
      const double *dp = (MODE == 3) ? Dq + ((e * q1d + q / q2) * (int64_t)slab + q % q2) : Dq + (e * (int64_t)slab + q);
      if (has_diff)
      {
         if (DIM == 3)
         {
            const double d11 = dp[0], d21 = dp[q2], d31 = dp[2 * q2], d22 = dp[3 * q2], d32 = dp[4 * q2], d33 = dp[5 * q2];
            a += gi[0] * (d11 * gj[0] + d21 * gj[1] + d31 * gj[2]) + gi[1] * (d21 * gj[0] + d22 * gj[1] + d32 * gj[2])
                 + gi[2] * (d31 * gj[0] + d32 * gj[1] + d33 * gj[2]);
         }
         else
         {
            const double d11 = dp[0], d21 = dp[q2], d22 = dp[2 * q2];
            a += gi[0] * (d11 * gj[0] + d21 * gj[1]) + gi[1] * (d21 * gj[0] + d22 * gj[1]);
         }
      }
      if (has_conv)
      {
         double s = dp[oc * q2] * gj[0] + dp[(oc + 1) * q2] * gj[1];
         if (DIM == 3) { s += dp[(oc + 2) * q2] * gj[2]; }
         a += pi * s;
      }
      if (has_mass) { a += dp[om * q2] * pi * pj; }
   }
   return a;
}

// Deterministic fill, one warp per matrix row g: the (element, local row) pairs that map to g are visited in the fixed
// order of the ElementRestriction transpose (offsets / indices); for each of them the lanes compute the nd entries of that
// element-matrix row and add them into the CSR row (distinct columns within one element, a warp barrier between
// elements): no atomics, the matrix is bit-reproducible.
template <int MODE>
__global__ void __launch_bounds__(256)
k_csr_fill_rows(BasisTables t, int d1d, int q1d, int nd, int nq, const double *__restrict__ sB, const double *__restrict__ sG,
                int64_t nrows, const int32_t *__restrict__ offsets, const int32_t *__restrict__ indices, const int32_t *__restrict__ gather,
                const double *__restrict__ Dq, int slab, int has_diff, int has_conv, int has_mass,
                const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colind, double *__restrict__ vals)
{
   const int lane = threadIdx.x & 31;
   const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
   for (int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < nrows; g += nwarps)
   {
      const int64_t r0 = rowptr[g], r1 = rowptr[g + 1];
      for (int64_t k = r0 + lane; k < r1; k += 32) { vals[k] = 0.0; }
      __syncwarp();
      for (int32_t k = offsets[g]; k < offsets[g + 1]; k++)
      {
         const int32_t idx = indices[k];
         const int64_t e = idx / nd;
         const int i = idx - (int32_t)(e * nd);
         for (int j = lane; j < nd; j += 32)
         {
            const double a = elem_entry<MODE>(t, d1d, q1d, nd, nq, sB, sG, Dq, slab, e, i, j, has_diff, has_conv, has_mass);
            const int32_t col = gather[e * nd + j];
            int64_t lo = r0, hi = r1 - 1;
            while (lo < hi)                                    // the pattern contains col by construction
            {
               const int64_t mid = (lo + hi) >> 1;
               if (colind[mid] < col) { lo = mid + 1; } else { hi = mid; }
            }
            vals[lo] += a;
         }
         __syncwarp();
      }
   }
}

__global__ void __launch_bounds__(256)
k_csr_diag(int64_t n, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colind, const double *__restrict__ vals, double *__restrict__ d)
{
   const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (g >= n) { return; }
   int64_t lo = rowptr[g], hi = rowptr[g + 1] - 1;
   while (lo < hi)
   {
      const int64_t mid = (lo + hi) >> 1;
      if (colind[mid] < g) { lo = mid + 1; } else { hi = mid; }
   }
   d[g] = vals[lo];
}

// y = A x (unconstrained) or the FormLinearSystem / ConstrainedOperator matrix: essential rows -> identity,
// essential columns skipped
__global__ void __launch_bounds__(256)
k_csr_spmv(int64_t n, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colind, const double *__restrict__ vals,
           const unsigned char *__restrict__ ess, const double *__restrict__ x, double *__restrict__ y)
{
   const int lane = threadIdx.x & 31;
   const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   if (row >= n) { return; }
   if (ess && ess[row]) { if (lane == 0) { y[row] = x[row]; } return; }
   double s = 0.0;
   for (int64_t k = rowptr[row] + lane; k < rowptr[row + 1]; k += 32)
   {
      const int32_t c = colind[k];
      if (!ess || !ess[c]) { s += vals[k] * x[c]; }
   }
   s = warp_sum(s);
   if (lane == 0) { y[row] = s; }
}
}  // namespace

void cdm_csr_destroy(cdm_op *op)
{
   cdm_csr *m = op->csr;
   cdm_ilu_destroy(op);
   if (!m) { return; }
   cudaFree(m->rowptr_dev); cudaFree(m->colind_dev); cudaFree(m->vals_dev); cudaFree(m->ess_mark_dev);
   delete m;
   op->csr = nullptr;
}

// values only (the pattern is kept): used by cdm_operator_update
static int csr_fill(cdm_op *op)
{
   cdm_space *sp = op->sp;
   cdm_ctx *c = sp->ctx;
   cdm_csr *m = op->csr;
   BasisTables t;
   memset(&t, 0, sizeof(t));
   if (sp->geom == 0) { for (int i = 0; i < sp->q1d * sp->d1d; i++) { t.B[i] = sp->B[i]; t.G[i] = sp->G[i]; } }
   const int64_t warps = std::min<int64_t>(m->n, (int64_t)c->sm_count * 64);
   const unsigned nb = (unsigned)((warps * 32 + 255) / 256);
#define FILL(MODE) k_csr_fill_rows<MODE><<<nb, 256, 0, c->stream>>>(t, sp->d1d, sp->q1d, sp->nd, sp->nq, sp->sB_dev, sp->sG_dev, m->n,     \
      sp->offsets_dev, sp->indices_dev, sp->gather_dev, op->D_dev, op->slab, op->has_diff, op->has_conv, op->has_mass,                     \
      m->rowptr_dev, m->colind_dev, m->vals_dev)
   if (sp->geom == 1) { FILL(1); }
   else if (sp->dim == 2) { FILL(2); }
   else { FILL(3); }
#undef FILL
   c->launches++;
   CDM_CUDA(c, cudaGetLastError());
   if (op->ilu) { return cdm_ilu_refactor(op); }             // the preconditioner follows the new values
   return CDM_OK;
}

int cdm_k_csr_diag(cdm_op *op, double *d)
{
   cdm_ctx *c = op->sp->ctx;
   cdm_csr *m = op->csr;
   if (!m) { return cdm_fail(c, CDM_EINVAL, "cdm_k_csr_diag: no assembled matrix"); }
   k_csr_diag<<<(unsigned)((m->n + 255) / 256), 256, 0, c->stream>>>(m->n, m->rowptr_dev, m->colind_dev, m->vals_dev, d);
   c->launches++;
   CDM_CUDA(c, cudaGetLastError());
   return CDM_OK;
}

int cdm_csr_refill_if_present(cdm_op *op) { return op->csr ? csr_fill(op) : CDM_OK; }

int cdm_k_csr_spmv(cdm_op *op, const double *x, double *y, bool constrained)
{
   cdm_ctx *c = op->sp->ctx;
   cdm_csr *m = op->csr;
   const int64_t threads = m->n * 32;
   const unsigned nb = (unsigned)((threads + 255) / 256);
   if (c->time_main) { cudaEventRecord(c->evk0, c->stream); }
   k_csr_spmv<<<nb, 256, 0, c->stream>>>(m->n, m->rowptr_dev, m->colind_dev, m->vals_dev,
                                         (constrained && op->n_ess > 0) ? m->ess_mark_dev : nullptr, x, y);
   if (c->time_main) { cudaEventRecord(c->evk1, c->stream); }
   c->launches++;
   CDM_CUDA(c, cudaGetLastError());
   return CDM_OK;
}

extern "C" {

int cdm_operator_assemble_csr(cdm_op *op)
{
   if (!op) { return CDM_EINVAL; }
   cdm_space *sp = op->sp;
   cdm_ctx *c = sp->ctx;
   CDM_REQUIRE_GPU(c);
   if (!sp->peers.empty() || sp->ntrue != sp->ndof) { return cdm_fail(c, CDM_EUNSUP, "cdm_operator_assemble_csr: single-rank spaces only"); }
   if (op->csr) { return csr_fill(op); }
   cdm_csr *m = new cdm_csr;
   op->csr = m;
   const int64_t n = sp->ndof;
   const int nd = sp->nd;
   m->n = n;
   m->rowptr.assign(n + 1, 0);
   // row g: the dofs of every element that touches g (offsets / indices = transpose of the gather map)
   std::vector<int32_t> row;
   for (int64_t g = 0; g < n; g++)
   {
      row.clear();
      for (int32_t k = sp->offsets[g]; k < sp->offsets[g + 1]; k++)
      {
         const int64_t e = sp->indices[k] / nd;
         row.insert(row.end(), sp->gather.begin() + e * nd, sp->gather.begin() + (e + 1) * nd);
      }
      std::sort(row.begin(), row.end());
      row.erase(std::unique(row.begin(), row.end()), row.end());
      m->colind.insert(m->colind.end(), row.begin(), row.end());
      m->rowptr[g + 1] = (int64_t)m->colind.size();
   }
   m->nnz = (int64_t)m->colind.size();
   std::vector<unsigned char> mark(n, 0);
   for (int32_t g : op->ess_host) { mark[g] = 1; }
   auto fail = [&](int rc) { cdm_csr_destroy(op); return rc; };
   if (cudaMalloc(&m->rowptr_dev, sizeof(int64_t) * (n + 1)) != cudaSuccess || cudaMalloc(&m->colind_dev, sizeof(int32_t) * (size_t)m->nnz) != cudaSuccess ||
       cudaMalloc(&m->vals_dev, sizeof(double) * (size_t)m->nnz) != cudaSuccess || cudaMalloc(&m->ess_mark_dev, (size_t)n) != cudaSuccess)
   { cudaGetLastError(); return fail(cdm_fail(c, CDM_ENOMEM, "cdm_operator_assemble_csr: out of device memory")); }
   cudaMemcpyAsync(m->rowptr_dev, m->rowptr.data(), sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice, c->stream);
   cudaMemcpyAsync(m->colind_dev, m->colind.data(), sizeof(int32_t) * (size_t)m->nnz, cudaMemcpyHostToDevice, c->stream);
   cudaMemcpyAsync(m->ess_mark_dev, mark.data(), (size_t)n, cudaMemcpyHostToDevice, c->stream);
   int rc = csr_fill(op);
   if (cudaStreamSynchronize(c->stream) != cudaSuccess && !rc) { rc = cdm_fail(c, CDM_ECUDA, "cdm_operator_assemble_csr: device failure"); }
   return rc ? fail(rc) : CDM_OK;
}

int cdm_operator_csr_sizes(const cdm_op *op, int64_t *nrows, int64_t *nnz)
{
   if (!op || !op->csr) { return CDM_EINVAL; }
   if (nrows) { *nrows = op->csr->n; }
   if (nnz) { *nnz = op->csr->nnz; }
   return CDM_OK;
}

int cdm_operator_csr_get(const cdm_op *op, int64_t *rowptr, int32_t *colind, double *vals)
{
   if (!op || !op->csr) { return CDM_EINVAL; }
   const cdm_csr *m = op->csr;
   cdm_ctx *c = op->sp->ctx;
   if (rowptr) { std::copy(m->rowptr.begin(), m->rowptr.end(), rowptr); }
   if (colind) { std::copy(m->colind.begin(), m->colind.end(), colind); }
   if (vals)
   {
      CDM_CUDA(c, cudaStreamSynchronize(c->stream));
      CDM_CUDA(c, cudaMemcpy(vals, m->vals_dev, sizeof(double) * (size_t)m->nnz, cudaMemcpyDeviceToHost));
   }
   return CDM_OK;
}

}  // extern "C"
