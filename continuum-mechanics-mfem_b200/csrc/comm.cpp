// comm.cpp -- mesh-partitioned multi-GPU mode: shared-dof halo exchange (P, P^T)
// and the Krylov all-reduce, over NCCL.  One rank per GPU.
//
// Stands behind ParMesh(MPI_COMM_WORLD, *mesh) / ParFiniteElementSpace
// (linear_convection_diffusion_2D.cpp:300,312): MFEM's prolongation P (owner ->
// sharers) and its transpose, and MPI_Allreduce inside InnerProduct
// (newton_petsc_solver.hpp:82-85).  NCCL is resolved at run time with dlopen so
// that the single-GPU path has no dependency on it.
#include "cdm_internal.hpp"
#include <dlfcn.h>
#include <cstring>

namespace
{
typedef struct { char internal[128]; } nccl_uid;
typedef int nccl_result;
struct NcclApi
{
   void *h = nullptr;
   nccl_result (*GetUniqueId)(nccl_uid *) = nullptr;
   nccl_result (*CommInitRank)(ncclComm **, int, nccl_uid, int) = nullptr;
   nccl_result (*CommDestroy)(ncclComm *) = nullptr;
   nccl_result (*AllReduce)(const void *, void *, size_t, int, int, ncclComm *, cudaStream_t) = nullptr;
   nccl_result (*AllGather)(const void *, void *, size_t, int, ncclComm *, cudaStream_t) = nullptr;
   nccl_result (*Send)(const void *, size_t, int, int, ncclComm *, cudaStream_t) = nullptr;
   nccl_result (*Recv)(void *, size_t, int, int, ncclComm *, cudaStream_t) = nullptr;
   nccl_result (*GroupStart)() = nullptr;
   nccl_result (*GroupEnd)() = nullptr;
   const char *(*GetErrorString)(nccl_result) = nullptr;
   nccl_result (*CommSplit)(ncclComm *, int, int, ncclComm **, void *) = nullptr;   // optional (NCCL >= 2.18)
   nccl_result (*GetVersion)(int *) = nullptr;                                      // optional (NCCL >= 2.3.4)
   int version = 0;
};
// The prototypes above and these enumerators are the NCCL 2.x ABI (nccl.h: ncclFloat64 = 8, ncclUint8 = 1, ncclSum = 0,
// ncclUniqueId = 128 bytes); api() refuses a library whose ncclGetVersion reports another major version.
const int NCCL_FLOAT64 = 8, NCCL_SUM = 0, NCCL_UINT8 = 1;

NcclApi *api()
{
   static NcclApi a;
   static bool tried = false;
   if (tried) { return a.h ? &a : nullptr; }
   tried = true;
   const char *names[] = {"libnccl.so.2", "libnccl.so"};
   for (const char *nm : names) { a.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (a.h) { break; } }
   if (!a.h) { return nullptr; }
#define SYM(field, name) *(void **)(&a.field) = dlsym(a.h, name); if (!a.field) { a.h = nullptr; return nullptr; }
   SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
   SYM(AllReduce, "ncclAllReduce") SYM(AllGather, "ncclAllGather") SYM(Send, "ncclSend") SYM(Recv, "ncclRecv")
   SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
   *(void **)(&a.CommSplit) = dlsym(a.h, "ncclCommSplit");
   *(void **)(&a.GetVersion) = dlsym(a.h, "ncclGetVersion");
   if (a.GetVersion && a.GetVersion(&a.version) == 0)
   {
      // version code: major * 10000 + minor * 100 + patch since 2.9, major * 1000 + minor * 100 + patch before
      const int major = (a.version >= 10000) ? a.version / 10000 : a.version / 1000;
      if (major != 2) { a.h = nullptr; return nullptr; }
   }
   return &a;
}
}  // namespace

// inside ncclGroupStart ... ncclGroupEnd: close the group before reporting the error
#define NCCL_GROUP_CALL(ctx, call)                                                             \
   do { nccl_result r_ = (call);                                                               \
        if (r_ != 0) { api()->GroupEnd();                                                      \
           return cdm_fail(ctx, CDM_ENCCL, std::string(#call) + ": " + api()->GetErrorString(r_)); } \
   } while (0)
#define NCCL_CALL(ctx, call)                                                                   \
   do { nccl_result r_ = (call);                                                               \
        if (r_ != 0) { return cdm_fail(ctx, CDM_ENCCL, std::string(#call) + ": " + api()->GetErrorString(r_)); } \
   } while (0)

extern "C" int cdm_comm_unique_id(void *uid128)
{
   if (!uid128) { return CDM_EINVAL; }
   NcclApi *a = api();
   if (!a) { return CDM_ENCCL; }
   nccl_uid id;
   if (a->GetUniqueId(&id) != 0) { return CDM_ENCCL; }
   std::memcpy(uid128, &id, sizeof(id));
   return CDM_OK;
}

extern "C" int cdm_comm_init(cdm_ctx *ctx, int rank, int nranks, const void *uid128)
{
   CDM_REQUIRE_GPU(ctx);
   if (nranks < 1 || rank < 0 || rank >= nranks) { return cdm_fail(ctx, CDM_EINVAL, "cdm_comm_init: bad rank"); }
   ctx->rank = rank; ctx->nranks = nranks;
   if (nranks == 1) { return CDM_OK; }
   if (!uid128) { return cdm_fail(ctx, CDM_EINVAL, "cdm_comm_init: missing unique id"); }
   NcclApi *a = api();
   if (!a) { return cdm_fail(ctx, CDM_ENCCL, "cdm_comm_init: libnccl.so.2 not found"); }
   nccl_uid id;
   std::memcpy(&id, uid128, sizeof(id));
   CDM_CUDA(ctx, cudaSetDevice(ctx->device));
   NCCL_CALL(ctx, a->CommInitRank(&ctx->comm, nranks, id, rank));
   // second communicator + high-priority stream for the halo exchange, so that it can run beside the
   // element kernels of the compute stream (and beside the Krylov all-reduces of the main communicator)
   {
      int lo = 0, hi = 0;
      cudaDeviceGetStreamPriorityRange(&lo, &hi);
      if (cudaStreamCreateWithPriority(&ctx->stream_halo, cudaStreamNonBlocking, hi) != cudaSuccess)
      { cudaGetLastError(); ctx->stream_halo = nullptr; }
      for (int i = 0; i < 6 && ctx->stream_halo; i++)
         if (cudaEventCreateWithFlags(&ctx->ev_h[i], cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); ctx->stream_halo = nullptr; }
   }
   // the NCCL exchanges that overlap compute need their own communicator; the peer-memory exchange only the stream
   if (!(a->CommSplit && a->CommSplit(ctx->comm, 0, rank, &ctx->comm_halo, nullptr) == 0 && ctx->comm_halo)) { ctx->comm_halo = nullptr; }
   return CDM_OK;
}

extern "C" int cdm_comm_rank(const cdm_ctx *ctx, int *rank, int *nranks)
{
   if (!ctx) { return CDM_EINVAL; }
   if (rank) { *rank = ctx->rank; }
   if (nranks) { *nranks = ctx->nranks; }
   return CDM_OK;
}

int cdm_allreduce_sum(cdm_ctx *c, double *buf_dev, int k)
{
   if (c->nranks <= 1 || !c->comm) { return CDM_OK; }
   // peer-memory all-reduce (two tiny kernels, rank-ordered sum) once a partitioned space has set it up
   int rc = CDM_OK;
   if (c->allreduce_mode != 0 && cdm_allreduce_sym(c, buf_dev, k, &rc)) { return rc; }
   NCCL_CALL(c, api()->AllReduce(buf_dev, buf_dev, (size_t)k, NCCL_FLOAT64, NCCL_SUM, c->comm, c->stream));
   return CDM_OK;
}

// setup-time exchange of small records between the ranks (peer-memory halo: IPC handles and offsets)
int cdm_allgather_bytes(cdm_ctx *c, const void *send_dev, void *recv_dev, size_t bytes_per_rank)
{
   if (c->nranks <= 1 || !c->comm) { return cdm_fail(c, CDM_ENCCL, "cdm_allgather_bytes: no communicator"); }
   NCCL_CALL(c, api()->AllGather(send_dev, recv_dev, bytes_per_rank, NCCL_UINT8, c->comm, c->stream));
   return CDM_OK;
}

// Run `body` with the context's stream temporarily replaced (the kernel launchers read c->stream).
struct StreamSwap
{
   cdm_ctx *c; cudaStream_t saved;
   StreamSwap(cdm_ctx *ctx, cudaStream_t s) : c(ctx), saved(ctx->stream) { c->stream = s; }
   ~StreamSwap() { c->stream = saved; }
};

// x_L ghost entries <- owner values.  halo_stream: run on the halo stream / communicator and record
// ev_packed right after the pack kernel (the compute stream waits on it before filling the GPU).
static int halo_P_impl(cdm_space *sp, double *xL, bool halo_stream, cudaEvent_t ev_packed)
{
   cdm_ctx *c = sp->ctx;
   if (!c->comm) { return cdm_fail(c, CDM_ENCCL, "partitioned space used without cdm_comm_init"); }
   if (sp->halo.p2p && sp->halo.p2p_active) { return cdm_halo_p2p_P(sp, xL, halo_stream ? c->stream_halo : c->stream, ev_packed); }
   NcclApi *a = api();
   cdm_halo_plan &hp = sp->halo;
   StreamSwap sw(c, halo_stream ? c->stream_halo : c->stream);
   ncclComm *comm = halo_stream ? c->comm_halo : c->comm;
   int rc = cdm_k_pack(c, (int64_t)hp.own_all.size(), hp.own_all_dev, xL, hp.send_dev);
   if (rc) { return rc; }
   if (ev_packed) { CDM_CUDA(c, cudaEventRecord(ev_packed, c->stream)); }
   NCCL_CALL(c, a->GroupStart());
   for (auto &pr : sp->peers)
   {
      if (!pr.own_idx.empty()) { NCCL_GROUP_CALL(c, a->Send(hp.send_dev + pr.own_off, pr.own_idx.size(), NCCL_FLOAT64, pr.rank, comm, c->stream)); }
      if (!pr.ghost_idx.empty()) { NCCL_GROUP_CALL(c, a->Recv(hp.recv_dev + pr.ghost_off, pr.ghost_idx.size(), NCCL_FLOAT64, pr.rank, comm, c->stream)); }
   }
   NCCL_CALL(c, a->GroupEnd());
   return cdm_k_unpack(c, (int64_t)hp.ghost_all.size(), hp.ghost_all_dev, hp.recv_dev, xL, 0);
}

int cdm_halo_P(cdm_op *op, double *xL) { return halo_P_impl(op->sp, xL, false, nullptr); }
int cdm_halo_P_async(cdm_op *op, double *xL, cudaEvent_t ev_packed) { return halo_P_impl(op->sp, xL, true, ev_packed); }
int cdm_halo_P_space(cdm_space *sp, double *xL) { return halo_P_impl(sp, xL, false, nullptr); }

// owner entries of y_L += partial sums held in the sharers' ghost entries (fixed peer order)
static int halo_PT_impl(cdm_space *sp, double *yL, bool halo_stream, cudaEvent_t ev_packed)
{
   cdm_ctx *c = sp->ctx;
   if (!c->comm) { return cdm_fail(c, CDM_ENCCL, "partitioned space used without cdm_comm_init"); }
   if (sp->halo.p2p && sp->halo.p2p_active) { return cdm_halo_p2p_PT(sp, yL, halo_stream ? c->stream_halo : c->stream, ev_packed); }
   NcclApi *a = api();
   cdm_halo_plan &hp = sp->halo;
   StreamSwap sw(c, halo_stream ? c->stream_halo : c->stream);
   ncclComm *comm = halo_stream ? c->comm_halo : c->comm;
   int rc = cdm_k_pack(c, (int64_t)hp.ghost_all.size(), hp.ghost_all_dev, yL, hp.send_dev);
   if (rc) { return rc; }
   if (ev_packed) { CDM_CUDA(c, cudaEventRecord(ev_packed, c->stream)); }
   NCCL_CALL(c, a->GroupStart());
   for (auto &pr : sp->peers)
   {
      if (!pr.ghost_idx.empty()) { NCCL_GROUP_CALL(c, a->Send(hp.send_dev + pr.ghost_off, pr.ghost_idx.size(), NCCL_FLOAT64, pr.rank, comm, c->stream)); }
      if (!pr.own_idx.empty()) { NCCL_GROUP_CALL(c, a->Recv(hp.recv_dev + pr.own_off, pr.own_idx.size(), NCCL_FLOAT64, pr.rank, comm, c->stream)); }
   }
   NCCL_CALL(c, a->GroupEnd());
   return cdm_k_unpack_add_csr(c, (int64_t)hp.pt_dof.size(), hp.pt_dof_dev, hp.pt_off_dev, hp.pt_src_dev, hp.recv_dev, yL);
}

int cdm_halo_PT(cdm_op *op, double *yL) { return halo_PT_impl(op->sp, yL, false, nullptr); }
int cdm_halo_PT_async(cdm_op *op, double *yL, cudaEvent_t ev_packed) { return halo_PT_impl(op->sp, yL, true, ev_packed); }
int cdm_halo_PT_space(cdm_space *sp, double *yL) { return halo_PT_impl(sp, yL, false, nullptr); }
