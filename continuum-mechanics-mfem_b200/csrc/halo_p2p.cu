// halo_p2p.cu -- shared-dof exchange (P, P^T) over peer memory instead of NCCL send/recv
// (operator option "halo" = 1; experimental in round 1: validated on 2 GPUs, NCCL stays the default).
//
// Stands behind the same reference interfaces as comm.cpp: MFEM's prolongation P (owner -> sharers)
// and its transpose on a ParFiniteElementSpace (linear_convection_diffusion_2D.cpp:300,312).
//
// Each rank exports two receive buffers (one per phase) and a flag array with CUDA IPC.  An exchange is
// two kernels per rank and no library call:
//   k_p2p_pack   : gathers the values to send and stores them STRAIGHT INTO THE NEIGHBOURS' receive buffers
//                  (NVLink stores); every thread fences, the last block to finish raises the epoch flag of
//                  this (rank, phase) in every neighbour's flag array (st.release.sys);
//   k_p2p_unpack : spins (ld.acquire.sys) until the flags of the neighbours it expects data from carry this
//                  epoch, then scatters / adds the received values, read past L1 (ld.global.cv).
// Buffer reuse is safe without acknowledgements because the operator always runs P and P^T as a pair:
// a rank writes its neighbour's P buffer for apply k+1 only after it has received that neighbour's P^T
// data of apply k, which the neighbour sent after it had consumed the P buffer of apply k (and vice
// versa); every pair of neighbours exchanges data in both phases (the own list of one is the ghost list
// of the other).  Exchanges that are not such a pair (Jacobi diagonal, linear form, error norm) keep
// using NCCL on separate buffers.
#include "cdm_internal.hpp"
#include "kernels_common.cuh"
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

constexpr int P2P_MAXP = 26;                 // neighbours of a box in a 3-D block partition

struct p2p_info                              // what every rank tells the others (all-gathered once)
{
   cudaIpcMemHandle_t h_recv[2], h_flags;
   int rank, npeers;
   int peer_rank[P2P_MAXP];
   long long own_off[P2P_MAXP], ghost_off[P2P_MAXP];
};

struct p2p_route                             // kernel parameter: where my packed values go / whom I wait for
{
   int npeers;
   long long off[P2P_MAXP + 1];              // my send segments (own lists for P, ghost lists for P^T)
   double *dst[P2P_MAXP];                    // neighbour's receive buffer + its offset for me
   unsigned long long *flag[P2P_MAXP];       // my slot in the neighbour's flag array (this phase)
   int expect[P2P_MAXP];                     // 1: this neighbour sends me data in this phase
};

struct p2p_state
{
   double *recv[2] = {nullptr, nullptr};     // my receive buffers: phase 0 = P (ghost layout), 1 = P^T (own layout)
   unsigned long long *flags = nullptr;      // [2][P2P_MAXP], written by the neighbours
   unsigned int *done = nullptr;             // [2] block counters of the pack kernels
   unsigned long long epoch[2] = {0, 0};
   p2p_route route[2];
   std::vector<void *> opened;               // IPC mappings to close
};

namespace
{
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
   unsigned long long v;
   asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
   return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
   asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(256)
k_p2p_pack(const p2p_route r, const long long n, const int32_t *__restrict__ idx, const double *__restrict__ x,
           unsigned int *done, const unsigned long long epoch)
{
   const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n)
   {
      int j = 0;
      while (j + 1 < r.npeers && i >= r.off[j + 1]) { j++; }
      r.dst[j][i - r.off[j]] = x[idx[i]];
   }
   __threadfence_system();                                 // my store is performed before the block counts itself
   __syncthreads();
   if (threadIdx.x == 0)
   {
      const unsigned int prev = atomicAdd(done, 1u);
      if (prev == gridDim.x - 1)                           // last block: every store of the grid is visible
      {
         *done = 0u;
         __threadfence_system();
         for (int j = 0; j < r.npeers; j++)
            if (r.off[j + 1] > r.off[j]) { st_release_sys(r.flag[j], epoch); }
      }
   }
}

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
   unsigned long long t;
   asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
   return t;
}

constexpr unsigned long long P2P_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;    // 20 s of wall clock

// Wait until the flag of neighbour slot `slot` carries `epoch`.  Bounded by the global timer: a neighbour that
// never posts (a rank died, mismatched call sequences) must not wedge the GPU.  On timeout the error word (mapped
// pinned host memory) receives slot + 1; the host turns it into CDM_ENCCL at its next synchronisation point, so the
// stale data scattered after a timeout is never reported as a valid result.
__device__ __forceinline__ void p2p_wait_one(const unsigned long long *flag, unsigned long long epoch, int slot, unsigned int *err)
{
   if (ld_acquire_sys(flag) >= epoch) { return; }
   const unsigned long long t0 = globaltimer_ns();
   while (ld_acquire_sys(flag) < epoch)
   {
      __nanosleep(100);
      if (globaltimer_ns() - t0 > P2P_TIMEOUT_NS) { atomicExch_system(err, (unsigned int)(slot + 1)); __threadfence_system(); break; }
   }
}

__device__ __forceinline__ void p2p_wait(const p2p_route &r, const unsigned long long *my_flags, unsigned long long epoch, unsigned int *err)
{
   if (threadIdx.x < r.npeers && r.expect[threadIdx.x]) { p2p_wait_one(my_flags + threadIdx.x, epoch, threadIdx.x, err); }
   __syncthreads();
}

__global__ void __launch_bounds__(256)
k_p2p_unpack_P(const p2p_route r, const unsigned long long *my_flags, const unsigned long long epoch, const long long n,
               const int32_t *__restrict__ idx, const double *buf, double *__restrict__ x, unsigned int *err)
{
   p2p_wait(r, my_flags, epoch, err);
   const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) { x[idx[i]] = __ldcv(buf + i); }
}

// x[dof[k]] += sum_{j in [off[k], off[k+1])} buf[src[j]]  (fixed order, as the NCCL path)
__global__ void __launch_bounds__(256)
k_p2p_unpack_PT(const p2p_route r, const unsigned long long *my_flags, const unsigned long long epoch, const long long n,
                const int32_t *__restrict__ dof, const int32_t *__restrict__ off, const int32_t *__restrict__ src,
                const double *buf, double *__restrict__ x, unsigned int *err)
{
   p2p_wait(r, my_flags, epoch, err);
   const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) { return; }
   double s = x[dof[i]];
   for (int32_t j = off[i]; j < off[i + 1]; j++) { s += __ldcv(buf + src[j]); }
   x[dof[i]] = s;
}
}  // namespace

// the error word of the context: mapped pinned host memory, written by a kernel that timed out
static int ensure_err_word(cdm_ctx *c)
{
   if (c->p2p_err_host) { return CDM_OK; }
   CDM_CUDA(c, cudaHostAlloc((void **)&c->p2p_err_host, sizeof(unsigned int), cudaHostAllocMapped));
   *c->p2p_err_host = 0u;
   CDM_CUDA(c, cudaHostGetDevicePointer((void **)&c->p2p_err_dev, c->p2p_err_host, 0));
   return CDM_OK;
}

void cdm_halo_p2p_destroy(cdm_space *sp)
{
   p2p_state *st = sp->halo.p2p;
   if (!st) { return; }
   for (void *p : st->opened) { cudaIpcCloseMemHandle(p); }
   cudaFree(st->recv[0]); cudaFree(st->recv[1]); cudaFree(st->flags); cudaFree(st->done);
   delete st;
   sp->halo.p2p = nullptr;
}

int cdm_halo_p2p_setup(cdm_space *sp)
{
   cdm_ctx *c = sp->ctx;
   cdm_halo_plan &hp = sp->halo;
   if (hp.p2p) { return CDM_OK; }
   if (c->nranks <= 1 || sp->peers.empty()) { return CDM_OK; }
   if ((int)sp->peers.size() > P2P_MAXP) { return cdm_fail(c, CDM_EUNSUP, "peer-memory halo: more than 26 neighbours"); }
   { const int rc = ensure_err_word(c); if (rc) { return rc; } }
   p2p_state *st = new p2p_state;
   hp.p2p = st;
   auto fail = [&](int rc) { cdm_halo_p2p_destroy(sp); return rc; };
   const size_t nP = std::max<size_t>(hp.ghost_all.size(), 1), nPT = std::max<size_t>(hp.own_all.size(), 1);
   if (cudaMalloc(&st->recv[0], nP * sizeof(double)) != cudaSuccess || cudaMalloc(&st->recv[1], nPT * sizeof(double)) != cudaSuccess ||
       cudaMalloc(&st->flags, 2 * P2P_MAXP * sizeof(unsigned long long)) != cudaSuccess ||
       cudaMalloc(&st->done, 2 * sizeof(unsigned int)) != cudaSuccess)
   { cudaGetLastError(); return fail(cdm_fail(c, CDM_ENOMEM, "peer-memory halo: allocation failed")); }
   cudaMemsetAsync(st->flags, 0, 2 * P2P_MAXP * sizeof(unsigned long long), c->stream);
   cudaMemsetAsync(st->done, 0, 2 * sizeof(unsigned int), c->stream);
   // ---- tell everybody where my buffers are and how my peer lists are laid out
   p2p_info mine;
   memset(&mine, 0, sizeof(mine));
   if (cudaIpcGetMemHandle(&mine.h_recv[0], st->recv[0]) != cudaSuccess || cudaIpcGetMemHandle(&mine.h_recv[1], st->recv[1]) != cudaSuccess ||
       cudaIpcGetMemHandle(&mine.h_flags, st->flags) != cudaSuccess)
   { cudaGetLastError(); return fail(cdm_fail(c, CDM_ECUDA, "peer-memory halo: cudaIpcGetMemHandle failed")); }
   mine.rank = c->rank; mine.npeers = (int)sp->peers.size();
   for (int j = 0; j < mine.npeers; j++)
   {
      mine.peer_rank[j] = sp->peers[j].rank;
      mine.own_off[j] = sp->peers[j].own_off; mine.ghost_off[j] = sp->peers[j].ghost_off;
   }
   std::vector<p2p_info> all(c->nranks);
   {
      p2p_info *sd = nullptr, *rd = nullptr;
      if (cudaMalloc(&sd, sizeof(p2p_info)) != cudaSuccess || cudaMalloc(&rd, sizeof(p2p_info) * c->nranks) != cudaSuccess)
      { cudaGetLastError(); cudaFree(sd); return fail(cdm_fail(c, CDM_ENOMEM, "peer-memory halo: allocation failed")); }
      cudaMemcpyAsync(sd, &mine, sizeof(mine), cudaMemcpyHostToDevice, c->stream);
      int rc = cdm_allgather_bytes(c, sd, rd, sizeof(p2p_info));
      if (!rc && cudaMemcpyAsync(all.data(), rd, sizeof(p2p_info) * c->nranks, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) { rc = CDM_ECUDA; }
      if (cudaStreamSynchronize(c->stream) != cudaSuccess && !rc) { rc = CDM_ECUDA; }
      cudaFree(sd); cudaFree(rd);
      if (rc) { return fail(cdm_fail(c, rc, "peer-memory halo: all-gather of the exchange records failed")); }
   }
   // ---- map the neighbours' buffers and build the two routes
   for (int ph = 0; ph < 2; ph++)
   {
      p2p_route &r = st->route[ph];
      memset(&r, 0, sizeof(r));
      r.npeers = mine.npeers;
   }
   for (int j = 0; j < mine.npeers; j++)
   {
      const cdm_halo_peer &pr = sp->peers[j];
      const p2p_info &theirs = all[pr.rank];
      int m = -1;
      for (int k = 0; k < theirs.npeers; k++) { if (theirs.peer_rank[k] == c->rank) { m = k; } }
      if (theirs.rank != pr.rank || m < 0) { return fail(cdm_fail(c, CDM_EINVAL, "peer-memory halo: inconsistent neighbour lists")); }
      void *pP = nullptr, *pPT = nullptr, *pF = nullptr;
      if (cudaIpcOpenMemHandle(&pP, theirs.h_recv[0], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
          cudaIpcOpenMemHandle(&pPT, theirs.h_recv[1], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
          cudaIpcOpenMemHandle(&pF, theirs.h_flags, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
      {
         const std::string why = cudaGetErrorString(cudaGetLastError());
         if (pP) { st->opened.push_back(pP); } if (pPT) { st->opened.push_back(pPT); }
         return fail(cdm_fail(c, CDM_ECUDA, "peer-memory halo: cudaIpcOpenMemHandle failed: " + why));
      }
      st->opened.push_back(pP); st->opened.push_back(pPT); st->opened.push_back(pF);
      // phase 0 (P): I send my own list for this peer; it lands in the peer's ghost layout at ITS offset for me
      st->route[0].off[j] = pr.own_off; st->route[0].off[j + 1] = pr.own_off + (long long)pr.own_idx.size();
      st->route[0].dst[j] = (double *)pP + theirs.ghost_off[m];
      st->route[0].flag[j] = (unsigned long long *)pF + 0 * P2P_MAXP + m;
      st->route[0].expect[j] = pr.ghost_idx.empty() ? 0 : 1;
      // phase 1 (P^T): I send my ghost list for this peer; it lands in the peer's own layout
      st->route[1].off[j] = pr.ghost_off; st->route[1].off[j + 1] = pr.ghost_off + (long long)pr.ghost_idx.size();
      st->route[1].dst[j] = (double *)pPT + theirs.own_off[m];
      st->route[1].flag[j] = (unsigned long long *)pF + 1 * P2P_MAXP + m;
      st->route[1].expect[j] = pr.own_idx.empty() ? 0 : 1;
   }
   // nobody may write into a neighbour before that neighbour's buffers and flags are initialised
   double *tok = nullptr;
   if (cudaMalloc(&tok, sizeof(double)) != cudaSuccess) { cudaGetLastError(); return fail(cdm_fail(c, CDM_ENOMEM, "peer-memory halo: allocation failed")); }
   cudaMemsetAsync(tok, 0, sizeof(double), c->stream);
   int rc = cdm_allreduce_sum(c, tok, 1);
   if (cudaStreamSynchronize(c->stream) != cudaSuccess && !rc) { rc = CDM_ECUDA; }
   cudaFree(tok);
   if (rc) { return fail(cdm_fail(c, rc, "peer-memory halo: setup barrier failed")); }
   return CDM_OK;
}

static int p2p_exchange(cdm_space *sp, int ph, double *v, cudaStream_t s, cudaEvent_t ev_packed)
{
   cdm_ctx *c = sp->ctx;
   cdm_halo_plan &hp = sp->halo;
   p2p_state *st = hp.p2p;
   const unsigned long long epoch = ++st->epoch[ph];
   const p2p_route &r = st->route[ph];
   // send side
   const long long ns = (long long)(ph == 0 ? hp.own_all.size() : hp.ghost_all.size());
   if (ns > 0)
   {
      const unsigned nb = (unsigned)((ns + 255) / 256);
      k_p2p_pack<<<nb, 256, 0, s>>>(r, ns, ph == 0 ? hp.own_all_dev : hp.ghost_all_dev, v, st->done + ph, epoch);
      c->launches++;
   }
   if (ev_packed) { CDM_CUDA(c, cudaEventRecord(ev_packed, s)); }
   // receive side
   const unsigned long long *my_flags = st->flags + ph * P2P_MAXP;
   if (ph == 0)
   {
      const long long nr = (long long)hp.ghost_all.size();
      if (nr > 0)
      {
         k_p2p_unpack_P<<<(unsigned)((nr + 255) / 256), 256, 0, s>>>(r, my_flags, epoch, nr, hp.ghost_all_dev, st->recv[0], v, c->p2p_err_dev);
         c->launches++;
      }
   }
   else
   {
      const long long nr = (long long)hp.pt_dof.size();
      if (nr > 0)
      {
         k_p2p_unpack_PT<<<(unsigned)((nr + 255) / 256), 256, 0, s>>>(r, my_flags, epoch, nr, hp.pt_dof_dev, hp.pt_off_dev, hp.pt_src_dev,
                                                                    st->recv[1], v, c->p2p_err_dev);
         c->launches++;
      }
   }
   CDM_CUDA(c, cudaGetLastError());
   return CDM_OK;
}

int cdm_halo_p2p_P(cdm_space *sp, double *xL, cudaStream_t s, cudaEvent_t ev_packed) { return p2p_exchange(sp, 0, xL, s, ev_packed); }
int cdm_halo_p2p_PT(cdm_space *sp, double *yL, cudaStream_t s, cudaEvent_t ev_packed) { return p2p_exchange(sp, 1, yL, s, ev_packed); }

// =====================================================================================================
// Symmetric exchange ("halo sum") and the Krylov all-reduce over peer memory -- the default multi-GPU path.
//
//   exchange : ONE kernel pair per apply instead of the P and P^T pairs.  After the element kernel every rank holds
//              partial sums in all its shared dofs (owned or ghost).  k_sym_pack stores them straight into the
//              receive buffers of all ranks of each dof's sharing group (NVLink stores) and raises this rank's epoch
//              flag there; k_sym_unpack waits for the flags and adds, per shared dof, the contributions of the whole
//              group in ascending rank order (its own value at its own position).  Every sharer therefore ends up
//              with the bitwise identical sum: the result vector is consistent on the ghost entries as well, so the
//              NEXT apply needs no P exchange (the Krylov loop keeps its basis vectors ghost-consistent).
//              Receive buffers are double-buffered by epoch parity: a rank can run at most one exchange ahead of a
//              neighbour (its unpack of epoch E+1 needs the neighbour's flag E+1, posted after the neighbour's
//              unpack of epoch E), so buffer E&1 is free again when epoch E+2 is packed.  No acknowledgements.
//   all-reduce: k_red_post writes this rank's k partial scalars into slot [rank] of every rank's array and raises a
//              flag; k_red_sum waits for all flags and adds the slots in rank order (bitwise identical everywhere).
//              Two tiny kernels instead of ncclAllReduce (launch + protocol latency ~ 15-25 us at 8 GPUs).
// Stands behind ParFiniteElementSpace's group communicator (linear_convection_diffusion_2D.cpp:300,312) and
// MPI_Allreduce in InnerProduct (newton_petsc_solver.hpp:82-85).

constexpr int SYM_MAXR = 32;                  // ranks of the peer-memory all-reduce (one NVLink domain)

struct sym_info
{
   cudaIpcMemHandle_t h;                      // one allocation per rank: [recv 2 x total | flags | red 2 x R x MAXK | red flags]
   int rank, npeers;
   long long total;
   int peer_rank[P2P_MAXP];
   long long off[P2P_MAXP];
};

struct sym_route
{
   int npeers;
   long long off[P2P_MAXP + 1];
   double *dst[2][P2P_MAXP];                  // neighbour's receive buffer (parity 0 / 1) at its offset for me
   unsigned long long *flag[P2P_MAXP];        // my slot in the neighbour's flag array
};

struct red_route
{
   int nranks, me;
   double *dst[SYM_MAXR];                     // base of every rank's reduction array
   unsigned long long *flag[SYM_MAXR];        // my slot in every rank's reduction flags
};

struct sym_state
{
   void *block = nullptr;                     // my exported allocation
   double *recv = nullptr;                    // [2][total]
   long long total = 0;
   unsigned long long *flags = nullptr;       // [P2P_MAXP]
   double *red = nullptr;                     // [2][nranks][CDM_RED_MAXK]
   unsigned long long *red_flags = nullptr;   // [SYM_MAXR]
   unsigned int *done = nullptr;              // block counter of the pack kernel (private)
   unsigned long long epoch = 0, red_epoch = 0;
   sym_route route;
   red_route rroute;
   bool red_ok = false;
   std::vector<void *> opened;
};

namespace
{
__global__ void __launch_bounds__(256)
k_sym_pack(const sym_route r, const int par, const long long n, const int32_t *__restrict__ idx, const double *__restrict__ y,
           unsigned int *done, const unsigned long long epoch)
{
   const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n)
   {
      int j = 0;
      while (j + 1 < r.npeers && i >= r.off[j + 1]) { j++; }
      r.dst[par][j][i - r.off[j]] = y[idx[i]];
   }
   __threadfence_system();
   __syncthreads();
   if (threadIdx.x == 0)
   {
      const unsigned int prev = atomicAdd(done, 1u);
      if (prev == gridDim.x - 1)
      {
         *done = 0u;
         __threadfence_system();
         for (int j = 0; j < r.npeers; j++) { st_release_sys(r.flag[j], epoch); }
      }
   }
}

__global__ void __launch_bounds__(256)
k_sym_unpack(const int npeers, const unsigned long long *my_flags, const unsigned long long epoch, const long long n,
             const int32_t *__restrict__ dof, const int32_t *__restrict__ off, const int32_t *__restrict__ src,
             const double *buf, double *__restrict__ y, unsigned int *err)
{
   if (threadIdx.x < npeers) { p2p_wait_one(my_flags + threadIdx.x, epoch, threadIdx.x, err); }
   __syncthreads();
   const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   if (k >= n) { return; }
   const int32_t g = dof[k];
   const double own = y[g];
   double s = 0.0;
   for (int32_t j = off[k]; j < off[k + 1]; j++)
   {
      const int32_t q = src[j];
      s += (q < 0) ? own : __ldcv(buf + q);
   }
   y[g] = s;
}

// one block: my k scalars -> slot [me] of every rank (including my own), then the flags
__global__ void __launch_bounds__(64)
k_red_post(const red_route r, const int par, const int k, const double *__restrict__ vals, const unsigned long long epoch)
{
   const size_t slot = ((size_t)par * r.nranks + r.me) * CDM_RED_MAXK;
   for (int t = threadIdx.x; t < k * r.nranks; t += blockDim.x)
   {
      const int rk = t / k, j = t % k;
      r.dst[rk][slot + j] = vals[j];
   }
   __threadfence_system();
   __syncthreads();
   if (threadIdx.x < r.nranks) { st_release_sys(r.flag[threadIdx.x], epoch); }
}

// one block: wait for every rank's post, then out[j] = sum over ranks (rank order) of slot[r][j]
__global__ void __launch_bounds__(64)
k_red_sum(const int nranks, const int par, const int k, const double *red, const unsigned long long *my_flags,
          const unsigned long long epoch, double *__restrict__ out, unsigned int *err)
{
   if (threadIdx.x < nranks) { p2p_wait_one(my_flags + threadIdx.x, epoch, P2P_MAXP + threadIdx.x, err); }
   __syncthreads();
   for (int j = threadIdx.x; j < k; j += blockDim.x)
   {
      double s = 0.0;
      for (int rk = 0; rk < nranks; rk++) { s += __ldcv(red + ((size_t)par * nranks + rk) * CDM_RED_MAXK + j); }
      out[j] = s;
   }
}
}  // namespace

void cdm_halo_sym_destroy(cdm_space *sp)
{
   sym_state *st = sp->sym.st;
   if (!st) { return; }
   for (void *p : st->opened) { cudaIpcCloseMemHandle(p); }
   cudaFree(st->block); cudaFree(st->done);
   delete st;
   sp->sym.st = nullptr;
   if (sp->ctx && sp->ctx->red_sym == st) { sp->ctx->red_sym = nullptr; }
}

// Collective over the ranks of the communicator.  On any failure every rank falls back to the NCCL P / P^T path
// (sy.ready = -1): the decision is all-reduced so that the ranks never disagree about the protocol.
int cdm_halo_sym_setup(cdm_space *sp)
{
   cdm_ctx *c = sp->ctx;
   cdm_sym_plan &sy = sp->sym;
   if (sy.ready != 0) { return CDM_OK; }
   if (c->nranks <= 1) { sy.ready = -1; return CDM_OK; }
   int rc = ensure_err_word(c);
   if (rc) { return rc; }
   int bad = 0;
   const int npeers = (int)sy.peers.size();
   if (npeers > P2P_MAXP) { bad = 1; }
   sym_state *st = new sym_state;
   sy.st = st;
   st->total = std::max<long long>((long long)sy.all.size(), 1);
   const bool red_ok = c->nranks <= SYM_MAXR;
   const size_t b_recv = sizeof(double) * 2 * (size_t)st->total;
   const size_t b_flags = sizeof(unsigned long long) * P2P_MAXP;
   const size_t b_red = sizeof(double) * 2 * (size_t)SYM_MAXR * CDM_RED_MAXK;
   const size_t b_rflags = sizeof(unsigned long long) * SYM_MAXR;
   const size_t o_flags = (b_recv + 255) & ~(size_t)255, o_red = o_flags + ((b_flags + 255) & ~(size_t)255),
                o_rflags = o_red + b_red, bytes = o_rflags + b_rflags;
   sym_info mine;
   memset(&mine, 0, sizeof(mine));
   if (!bad)
   {
      if (cudaMalloc(&st->block, bytes) != cudaSuccess || cudaMalloc(&st->done, sizeof(unsigned int)) != cudaSuccess) { cudaGetLastError(); bad = 1; }
      else
      {
         cudaMemsetAsync(st->block, 0, bytes, c->stream);
         cudaMemsetAsync(st->done, 0, sizeof(unsigned int), c->stream);
         st->recv = (double *)st->block;
         st->flags = (unsigned long long *)((char *)st->block + o_flags);
         st->red = (double *)((char *)st->block + o_red);
         st->red_flags = (unsigned long long *)((char *)st->block + o_rflags);
         if (cudaIpcGetMemHandle(&mine.h, st->block) != cudaSuccess) { cudaGetLastError(); bad = 1; }
      }
   }
   mine.rank = c->rank; mine.npeers = bad ? 0 : npeers; mine.total = st->total;
   for (int j = 0; j < mine.npeers; j++) { mine.peer_rank[j] = sy.peers[j].rank; mine.off[j] = sy.peers[j].off; }
   std::vector<sym_info> all(c->nranks);
   {
      sym_info *sd = nullptr, *rd = nullptr;
      if (cudaMalloc(&sd, sizeof(sym_info)) != cudaSuccess || cudaMalloc(&rd, sizeof(sym_info) * c->nranks) != cudaSuccess)
      { cudaGetLastError(); cudaFree(sd); return cdm_fail(c, CDM_ENOMEM, "symmetric halo: allocation failed"); }
      cudaMemcpyAsync(sd, &mine, sizeof(mine), cudaMemcpyHostToDevice, c->stream);
      rc = cdm_allgather_bytes(c, sd, rd, sizeof(sym_info));
      if (!rc && cudaMemcpyAsync(all.data(), rd, sizeof(sym_info) * c->nranks, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) { rc = CDM_ECUDA; }
      if (cudaStreamSynchronize(c->stream) != cudaSuccess && !rc) { rc = CDM_ECUDA; }
      cudaFree(sd); cudaFree(rd);
      if (rc) { return cdm_fail(c, rc, "symmetric halo: all-gather of the exchange records failed"); }
   }
   // map every other rank's block (the all-reduce needs all of them, the exchange only the neighbours)
   std::vector<char *> base(c->nranks, nullptr);
   base[c->rank] = (char *)st->block;
   for (int r = 0; r < c->nranks && !bad; r++)
   {
      if (r == c->rank) { continue; }
      bool need = red_ok;
      for (int j = 0; j < npeers; j++) { if (sy.peers[j].rank == r) { need = true; } }
      if (!need) { continue; }
      void *p = nullptr;
      if (cudaIpcOpenMemHandle(&p, all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); bad = 1; break; }
      st->opened.push_back(p);
      base[r] = (char *)p;
   }
   memset(&st->route, 0, sizeof(st->route));
   memset(&st->rroute, 0, sizeof(st->rroute));
   st->route.npeers = npeers;
   for (int j = 0; j < npeers && !bad; j++)
   {
      const cdm_sym_peer &pr = sy.peers[j];
      const sym_info &th = all[pr.rank];
      int m = -1;
      for (int k = 0; k < th.npeers; k++) { if (th.peer_rank[k] == c->rank) { m = k; } }
      if (th.rank != pr.rank || m < 0) { bad = 1; break; }
      st->route.off[j] = pr.off; st->route.off[j + 1] = pr.off + (long long)pr.idx.size();
      double *their_recv = (double *)base[pr.rank];
      st->route.dst[0][j] = their_recv + th.off[m];
      st->route.dst[1][j] = their_recv + th.total + th.off[m];
      const size_t their_o_flags = (sizeof(double) * 2 * (size_t)th.total + 255) & ~(size_t)255;
      st->route.flag[j] = (unsigned long long *)(base[pr.rank] + their_o_flags) + m;
   }
   if (!bad && red_ok)
   {
      st->rroute.nranks = c->nranks; st->rroute.me = c->rank;
      for (int r = 0; r < c->nranks; r++)
      {
         const size_t their_o_flags = (sizeof(double) * 2 * (size_t)all[r].total + 255) & ~(size_t)255;
         const size_t their_o_red = their_o_flags + ((b_flags + 255) & ~(size_t)255);
         st->rroute.dst[r] = (double *)(base[r] + their_o_red);
         st->rroute.flag[r] = (unsigned long long *)(base[r] + their_o_red + b_red) + c->rank;
      }
      st->red_ok = true;
   }
   // agree on the outcome (and make sure nobody writes into a neighbour before its block is zeroed)
   double *tok = nullptr;
   if (cudaMalloc(&tok, sizeof(double)) != cudaSuccess) { cudaGetLastError(); return cdm_fail(c, CDM_ENOMEM, "symmetric halo: allocation failed"); }
   const double mybad = bad ? 1.0 : 0.0;
   cudaMemcpyAsync(tok, &mybad, sizeof(double), cudaMemcpyHostToDevice, c->stream);
   rc = cdm_allreduce_sum(c, tok, 1);
   double anybad = 1.0;
   if (!rc && cudaMemcpyAsync(&anybad, tok, sizeof(double), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) { rc = CDM_ECUDA; }
   if (cudaStreamSynchronize(c->stream) != cudaSuccess && !rc) { rc = CDM_ECUDA; }
   cudaFree(tok);
   if (rc) { return cdm_fail(c, rc, "symmetric halo: setup barrier failed"); }
   if (anybad != 0.0) { cdm_halo_sym_destroy(sp); sy.ready = -1; return CDM_OK; }
   sy.ready = 1;
   if (st->red_ok && !c->red_sym) { c->red_sym = st; }
   return CDM_OK;
}

// y (partial sums in every shared dof) -> consistent sums in every shared dof, owned or ghost, on stream s
int cdm_halo_sym_exchange(cdm_space *sp, double *yL, cudaStream_t s)
{
   cdm_ctx *c = sp->ctx;
   cdm_sym_plan &sy = sp->sym;
   sym_state *st = sy.st;
   if (!st || sy.ready != 1) { return cdm_fail(c, CDM_ENCCL, "symmetric halo exchange used before its setup"); }
   const unsigned long long epoch = ++st->epoch;
   const int par = (int)(epoch & 1ull);
   const long long ns = (long long)sy.all.size(), nr = (long long)sy.sh_dof.size();
   if (ns > 0)
   {
      k_sym_pack<<<(unsigned)((ns + 255) / 256), 256, 0, s>>>(st->route, par, ns, sy.all_dev, yL, st->done, epoch);
      c->launches++;
   }
   if (nr > 0)
   {
      k_sym_unpack<<<(unsigned)((nr + 255) / 256), 256, 0, s>>>(st->route.npeers, st->flags, epoch, nr, sy.sh_dof_dev, sy.sh_off_dev,
                                                               sy.sh_src_dev, st->recv + (size_t)par * st->total, yL, c->p2p_err_dev);
      c->launches++;
   }
   CDM_CUDA(c, cudaGetLastError());
   return CDM_OK;
}

// buf[0..k) <- sum over the ranks, in rank order, through peer memory (ctx->stream); false when unavailable
bool cdm_allreduce_sym(cdm_ctx *c, double *buf_dev, int k, int *rc_out)
{
   sym_state *st = c->red_sym;
   if (!st || !st->red_ok || k < 1 || k > CDM_RED_MAXK) { return false; }
   const unsigned long long epoch = ++st->red_epoch;
   const int par = (int)(epoch & 1ull);
   k_red_post<<<1, 64, 0, c->stream>>>(st->rroute, par, k, buf_dev, epoch);
   k_red_sum<<<1, 64, 0, c->stream>>>(c->nranks, par, k, st->red, st->red_flags, epoch, buf_dev, c->p2p_err_dev);
   c->launches += 2;
   *rc_out = (cudaGetLastError() == cudaSuccess) ? CDM_OK : cdm_fail(c, CDM_ECUDA, "peer-memory all-reduce launch failed");
   return true;
}
