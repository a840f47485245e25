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

__device__ __forceinline__ void p2p_wait(const p2p_route &r, const unsigned long long *my_flags, unsigned long long epoch)
{
   if (threadIdx.x < r.npeers && r.expect[threadIdx.x])
   {
      // bounded: a neighbour that never posts (a rank died, mismatched call sequences) must not wedge the GPU;
      // after ~5 s the kernel gives up, the result is wrong and the caller's checks fail instead of hanging
      unsigned int spins = 0;
      while (ld_acquire_sys(my_flags + threadIdx.x) < epoch)
      {
         __nanosleep(64);
         if (++spins > 40000000u) { break; }
      }
   }
   __syncthreads();
}

__global__ void __launch_bounds__(256)
k_p2p_unpack_P(const p2p_route r, const unsigned long long *my_flags, const unsigned long long epoch, const long long n,
               const int32_t *__restrict__ idx, const double *buf, double *__restrict__ x)
{
   p2p_wait(r, my_flags, epoch);
   const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) { x[idx[i]] = __ldcv(buf + i); }
}

// x[dof[k]] += sum_{j in [off[k], off[k+1])} buf[src[j]]  (fixed order, as the NCCL path)
__global__ void __launch_bounds__(256)
k_p2p_unpack_PT(const p2p_route r, const unsigned long long *my_flags, const unsigned long long epoch, const long long n,
                const int32_t *__restrict__ dof, const int32_t *__restrict__ off, const int32_t *__restrict__ src,
                const double *buf, double *__restrict__ x)
{
   p2p_wait(r, my_flags, epoch);
   const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) { return; }
   double s = x[dof[i]];
   for (int32_t j = off[i]; j < off[i + 1]; j++) { s += __ldcv(buf + src[j]); }
   x[dof[i]] = s;
}
}  // namespace

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
         k_p2p_unpack_P<<<(unsigned)((nr + 255) / 256), 256, 0, s>>>(r, my_flags, epoch, nr, hp.ghost_all_dev, st->recv[0], v);
         c->launches++;
      }
   }
   else
   {
      const long long nr = (long long)hp.pt_dof.size();
      if (nr > 0)
      {
         k_p2p_unpack_PT<<<(unsigned)((nr + 255) / 256), 256, 0, s>>>(r, my_flags, epoch, nr, hp.pt_dof_dev, hp.pt_off_dev, hp.pt_src_dev,
                                                                    st->recv[1], v);
         c->launches++;
      }
   }
   CDM_CUDA(c, cudaGetLastError());
   return CDM_OK;
}

int cdm_halo_p2p_P(cdm_space *sp, double *xL, cudaStream_t s, cudaEvent_t ev_packed) { return p2p_exchange(sp, 0, xL, s, ev_packed); }
int cdm_halo_p2p_PT(cdm_space *sp, double *yL, cudaStream_t s, cudaEvent_t ev_packed) { return p2p_exchange(sp, 1, yL, s, ev_packed); }
