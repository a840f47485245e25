// capi.cu -- C ABI entry points: context, space, operator, vectors.
// Every exported symbol is declared in include/cdm_b200.h together with the
// reference interface it stands behind.
#include "cdm_internal.hpp"
#include <thread>
#include "kernels_common.cuh"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#define CDM_VERSION_STR "cdm_b200 0.1 (sm_100a)"

namespace
{
template <typename T>
int upload(cdm_ctx *ctx, const std::vector<T> &h, T **d)
{
   if (h.empty()) { *d = nullptr; return CDM_OK; }
   CDM_CUDA(ctx, cudaMalloc((void **)d, h.size() * sizeof(T)));
   CDM_CUDA(ctx, cudaMemcpyAsync(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
   return CDM_OK;
}
inline int ipow(int b, int e) { int r = 1; while (e-- > 0) { r *= b; } return r; }
}

extern "C" {

const char *cdm_version(void) { return CDM_VERSION_STR; }

int cdm_init(int device, void *stream, cdm_ctx **out)
{
   if (!out) { return CDM_EINVAL; }
   *out = nullptr;
   int ndev = 0;
   if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev)
   {
      cudaGetLastError();
      return CDM_ENOGPU;
   }
   cdm_ctx *c = new (std::nothrow) cdm_ctx;
   if (!c) { return CDM_ENOMEM; }
   c->device = device;
   if (cudaSetDevice(device) != cudaSuccess) { delete c; return CDM_ENOGPU; }
   if (stream) { c->stream = (cudaStream_t)stream; }
   else
   {
      if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return CDM_ENOGPU; }
      c->own_stream = true;
   }
   cudaDeviceProp prop;
   if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) { c->sm_count = prop.multiProcessorCount; }
   const size_t nred = (size_t)CDM_RED_MAXK * CDM_RED_BLOCKS + 4 * CDM_RED_MAXK;
   if (cudaMalloc(&c->red_dev, nred * sizeof(double)) != cudaSuccess ||
       cudaMallocHost(&c->red_host, 4 * CDM_RED_MAXK * sizeof(double)) != cudaSuccess ||
       cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess)
   {
      cdm_finalize(c);
      return CDM_ENOGPU;
   }
   *out = c;
   return CDM_OK;
}

int cdm_init_host(cdm_ctx **out)
{
   if (!out) { return CDM_EINVAL; }
   cdm_ctx *c = new (std::nothrow) cdm_ctx;
   if (!c) { return CDM_ENOMEM; }
   c->device = -1;
   *out = c;
   return CDM_OK;
}

int cdm_finalize(cdm_ctx *c)
{
   if (!c) { return CDM_OK; }
   if (c->device >= 0)
   {
      cudaSetDevice(c->device);
      if (c->stream) { cudaStreamSynchronize(c->stream); }
      if (c->red_dev) { cudaFree(c->red_dev); }
      if (c->red_host) { cudaFreeHost(c->red_host); }
      if (c->p2p_err_host) { cudaFreeHost(c->p2p_err_host); }
      if (c->ev0) { cudaEventDestroy(c->ev0); }
      if (c->ev1) { cudaEventDestroy(c->ev1); }
      if (c->ev_kry) { cudaEventDestroy(c->ev_kry); }
      if (c->evk0) { cudaEventDestroy(c->evk0); }
      if (c->evk1) { cudaEventDestroy(c->evk1); }
      // the NCCL communicators are left to process teardown: ncclCommDestroy from an arbitrary point of the
      // host language's shutdown sequence can block on peers that are already gone
      for (int i = 0; i < 6; i++) { if (c->ev_h[i]) { cudaEventDestroy(c->ev_h[i]); } }
      if (c->stream_halo) { cudaStreamSynchronize(c->stream_halo); cudaStreamDestroy(c->stream_halo); }
      if (c->own_stream && c->stream) { cudaStreamDestroy(c->stream); }
   }
   delete c;
   return CDM_OK;
}

const char *cdm_last_error(const cdm_ctx *c) { return c ? c->err.c_str() : "null context"; }
void *cdm_stream(cdm_ctx *c) { return c ? (void *)c->stream : nullptr; }
int64_t cdm_launch_count(const cdm_ctx *c) { return c ? c->launches : 0; }

int cdm_sync(cdm_ctx *c)
{
   CDM_REQUIRE_GPU(c);
   CDM_CUDA(c, cudaStreamSynchronize(c->stream));
   return cdm_check_p2p(c);
}

// ------------------------------------------------------------------ space

static int space_finish(cdm_ctx *ctx, const cdm_mesh *mesh, cdm_space *sp)
{
   // per-element vertex coordinates (order-1 mesh nodes)
   const int nvpe = (sp->geom == 1) ? sp->dim + 1 : ((sp->dim == 2) ? 4 : 8);
   sp->elem_x.resize((size_t)sp->ne * nvpe * sp->dim);
   auto fill_x = [&](int64_t eb, int64_t ee)
   {
      for (int64_t e = eb; e < ee; e++)
      {
         const int64_t em = sp->elem_perm.empty() ? e : sp->elem_perm[e];      // element index in the mesh
         for (int k = 0; k < nvpe; k++)
            for (int c = 0; c < sp->dim; c++)
               sp->elem_x[((size_t)e * nvpe + k) * sp->dim + c] = mesh->vx[(size_t)mesh->ev[(size_t)em * nvpe + k] * sp->dim + c];
      }
   };
   {
      unsigned nt = std::thread::hardware_concurrency();      // independent elements: slices on the host threads
      if (nt > 16) { nt = 16; }
      if (nt < 2 || sp->ne < 65536) { fill_x(0, sp->ne); }
      else
      {
         std::vector<std::thread> th;
         for (unsigned t = 0; t < nt; t++) { th.emplace_back(fill_x, sp->ne * t / nt, sp->ne * (t + 1) / nt); }
         for (auto &x : th) { x.join(); }
      }
   }
   cdm_host_restriction(sp->ne, sp->nd, sp->ndof, sp->gather, sp->offsets, sp->indices);
   if (ctx->device >= 0)
   {
      int rc;
      if ((rc = upload(ctx, sp->gather, &sp->gather_dev))) { return rc; }
      if ((rc = upload(ctx, sp->offsets, &sp->offsets_dev))) { return rc; }
      if ((rc = upload(ctx, sp->indices, &sp->indices_dev))) { return rc; }
      if ((rc = upload(ctx, sp->elem_x, &sp->elem_x_dev))) { return rc; }
      if (sp->geom == 1)
      {
         if ((rc = upload(ctx, sp->sB, &sp->sB_dev))) { return rc; }
         if ((rc = upload(ctx, sp->sG, &sp->sG_dev))) { return rc; }
         if ((rc = upload(ctx, sp->sqw, &sp->sqw_dev))) { return rc; }
         if ((rc = upload(ctx, sp->sqx, &sp->sqx_dev))) { return rc; }
      }
      if (!sp->peers.empty())
      {
         // fused exchange plan: one pack and one unpack kernel per phase, whatever the peer count.
         // own_all / ghost_all = per-peer lists concatenated in peer order; P^T adds are done per
         // owned-shared dof in peer order (CSR), so the summation order is fixed.
         cdm_halo_plan &hp = sp->halo;
         for (auto &pr : sp->peers)
         {
            pr.own_off = (int64_t)hp.own_all.size(); pr.ghost_off = (int64_t)hp.ghost_all.size();
            hp.own_all.insert(hp.own_all.end(), pr.own_idx.begin(), pr.own_idx.end());
            hp.ghost_all.insert(hp.ghost_all.end(), pr.ghost_idx.begin(), pr.ghost_idx.end());
         }
         std::vector<int32_t> order(hp.own_all.size());
         for (size_t i = 0; i < order.size(); i++) { order[i] = (int32_t)i; }
         std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return hp.own_all[a] < hp.own_all[b]; });
         for (size_t i = 0; i < order.size(); i++)
         {
            if (i == 0 || hp.own_all[order[i]] != hp.own_all[order[i - 1]])
            {
               hp.pt_dof.push_back(hp.own_all[order[i]]);
               hp.pt_off.push_back((int32_t)i);
            }
         }
         hp.pt_off.push_back((int32_t)order.size());
         hp.pt_src = order;
         if ((rc = upload(ctx, hp.own_all, &hp.own_all_dev))) { return rc; }
         if ((rc = upload(ctx, hp.ghost_all, &hp.ghost_all_dev))) { return rc; }
         if ((rc = upload(ctx, hp.pt_dof, &hp.pt_dof_dev))) { return rc; }
         if ((rc = upload(ctx, hp.pt_off, &hp.pt_off_dev))) { return rc; }
         if ((rc = upload(ctx, hp.pt_src, &hp.pt_src_dev))) { return rc; }
         const size_t nb = (std::max(hp.own_all.size(), hp.ghost_all.size()) + 1) * sizeof(double);
         CDM_CUDA(ctx, cudaMalloc(&hp.send_dev, nb));
         CDM_CUDA(ctx, cudaMalloc(&hp.recv_dev, nb));
         cdm_sym_plan &sy = sp->sym;
         if ((rc = upload(ctx, sy.all, &sy.all_dev))) { return rc; }
         if ((rc = upload(ctx, sy.sh_dof, &sy.sh_dof_dev))) { return rc; }
         if ((rc = upload(ctx, sy.sh_off, &sy.sh_off_dev))) { return rc; }
         if ((rc = upload(ctx, sy.sh_src, &sy.sh_src_dev))) { return rc; }
      }
      CDM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
   }
   return CDM_OK;
}

static cdm_space *space_new(cdm_ctx *ctx, const cdm_mesh *mesh, int order)
{
   cdm_space *sp = new (std::nothrow) cdm_space;
   if (!sp) { return nullptr; }
   sp->ctx = ctx; sp->dim = mesh->dim; sp->p = order; sp->d1d = order + 1;
   sp->geom = mesh->geom;
   sp->q1d = cdm_host_q1d(mesh->dim, order);
   sp->nd = ipow(sp->d1d, sp->dim); sp->nq = ipow(sp->q1d, sp->dim);
   if (sp->geom == 1)
   {
      // order-p triangle: (p+1)(p+2)/2 nodes; collapsed Gauss-Legendre rule with p+1 points per direction (exact to
      // degree 2p: the constant-coefficient integrands of all three integrators on affine triangles)
      sp->q1d = order + 1;
      sp->nd = (order + 1) * (order + 2) / 2; sp->nq = sp->q1d * sp->q1d;
      sp->snodes.resize(2 * sp->nd); sp->sqx.resize(2 * sp->nq); sp->sqw.resize(sp->nq);
      sp->sB.resize((size_t)sp->nq * sp->nd); sp->sG.resize((size_t)2 * sp->nq * sp->nd);
      cdm_host_tri_nodes(order, sp->snodes.data());
      cdm_host_tri_rule(sp->q1d, sp->sqx.data(), sp->sqw.data());
      cdm_host_tri_basis(order, sp->nq, sp->sqx.data(), sp->sB.data(), sp->sG.data());
   }
   sp->ne = mesh->ne; sp->nv = mesh->nv; sp->nbe = mesh->nbe;
   sp->B.resize(sp->q1d * sp->d1d); sp->G.resize(sp->q1d * sp->d1d);
   sp->qw.resize(sp->q1d); sp->qx.resize(sp->q1d); sp->nodes.resize(sp->d1d);
   cdm_host_basis(order, sp->q1d, sp->B.data(), sp->G.data(), sp->qw.data(), sp->nodes.data(), sp->qx.data());
   sp->bdr_attr = mesh->battr;
   return sp;
}

// shared-dof plan of a box-partitioned Cartesian space + owned-first renumbering
// one shared dof as seen from this rank: P / P^T plan (owner <-> each ghost holder) and symmetric plan (every other sharer)
struct PShare { int64_t key; int32_t dof; int peer; bool mine; };
struct PSymShare { int64_t key; int32_t dof; int peer; };
static void finish_partition(cdm_space *sp, int me, const std::vector<int64_t> &key, const std::vector<int> &owner,
                             std::vector<PShare> &shares, std::vector<PSymShare> &sym);

static void build_partition(const cdm_mesh *m, cdm_space *sp)
{
   const int dim = sp->dim, p = sp->p, d1d = sp->d1d;
   const int64_t N[3] = {p * m->gn[0] + 1, p * m->gn[1] + 1, dim == 3 ? p * m->gn[2] + 1 : 1};
   // element index -> rank coordinate along each axis
   auto axis_rank = [&](int d, int64_t el) -> int
   {
      const int64_t q = m->gn[d] / m->parts[d], r = m->gn[d] % m->parts[d];
      // boxes 0..r-1 have q+1 elements, the rest q
      if (el < r * (q + 1)) { return (int)(el / (q + 1)); }
      return (int)(r + (el - r * (q + 1)) / q);
   };
   const int64_t ln[3] = {m->n[0], m->n[1], dim == 3 ? m->n[2] : 1};
   std::vector<int64_t> key(sp->ndof, -1);
   for (int64_t e = 0; e < sp->ne; e++)
   {
      const int64_t ei = e % ln[0], ej = (e / ln[0]) % ln[1], ek = e / (ln[0] * ln[1]);
      for (int l = 0; l < sp->nd; l++)
      {
         const int lx = l % d1d, ly = (l / d1d) % d1d, lz = (dim == 3) ? l / (d1d * d1d) : 0;
         const int64_t I = p * (m->lo[0] + ei) + lx, J = p * (m->lo[1] + ej) + ly,
                       K = (dim == 3) ? p * (m->lo[2] + ek) + lz : 0;
         key[sp->gather[(size_t)e * sp->nd + l]] = I + N[0] * (J + N[1] * K);
      }
   }
   const int me = m->rank;
   std::vector<int> owner(sp->ndof, me);
   std::vector<PShare> shares;
   std::vector<PSymShare> sym;                // (dof, every other rank of its sharing group)
   for (int64_t g = 0; g < sp->ndof; g++)
   {
      const int64_t k = key[g];
      const int64_t c[3] = {k % N[0], (k / N[0]) % N[1], k / (N[0] * N[1])};
      int rs[3][2], nr[3];
      for (int d = 0; d < 3; d++)
      {
         nr[d] = 1; rs[d][0] = 0;
         if (d >= dim) { continue; }
         const int64_t el_hi = std::min<int64_t>(c[d] / p, m->gn[d] - 1);
         rs[d][0] = axis_rank(d, el_hi);
         if (c[d] % p == 0 && c[d] > 0 && c[d] < p * m->gn[d])
         {
            const int r2 = axis_rank(d, c[d] / p - 1);
            if (r2 != rs[d][0]) { rs[d][1] = r2; nr[d] = 2; }
         }
      }
      int ranks[8], cnt = 0, own = 1 << 30;
      for (int a = 0; a < nr[0]; a++)
         for (int b = 0; b < nr[1]; b++)
            for (int cc = 0; cc < nr[2]; cc++)
            {
               const int r = rs[0][a] + m->parts[0] * (rs[1][b] + m->parts[1] * rs[2][cc]);
               ranks[cnt++] = r; own = std::min(own, r);
            }
      owner[g] = own;
      if (cnt == 1) { continue; }
      for (int i = 0; i < cnt; i++) if (ranks[i] != me) { sym.push_back({k, (int32_t)g, ranks[i]}); }
      if (own == me) { for (int i = 0; i < cnt; i++) if (ranks[i] != me) { shares.push_back({k, (int32_t)g, ranks[i], true}); } }
      else { shares.push_back({k, (int32_t)g, own, false}); }
   }
   finish_partition(sp, me, key, owner, shares, sym);
}

// The part of the plan that does not depend on how the sharing ranks were found: owned-first renumbering, boundary-first
// element order, the P / P^T peer lists and the symmetric exchange plan (both in key order, so that the two sides of a pair
// build the same lists without talking to each other).
static void finish_partition(cdm_space *sp, int me, const std::vector<int64_t> &key, const std::vector<int> &owner,
                             std::vector<PShare> &shares, std::vector<PSymShare> &sym)
{
   // owned-first renumbering
   std::vector<int32_t> newid(sp->ndof);
   int64_t nown = 0;
   for (int64_t g = 0; g < sp->ndof; g++) if (owner[g] == me) { newid[g] = (int32_t)nown++; }
   int64_t ng = nown;
   for (int64_t g = 0; g < sp->ndof; g++) if (owner[g] != me) { newid[g] = (int32_t)ng++; }
   sp->ntrue = nown;
   for (auto &g : sp->gather) { g = newid[g]; }
   for (auto &g : sp->bdr_dofs_flat) { g = newid[g]; }
   // boundary-first element order: elements touching a shared dof come first so that the halo
   // exchange can overlap the interior elements
   {
      std::vector<uint8_t> shared(sp->ndof, 0);
      for (const PShare &s : shares) { shared[newid[s.dof]] = 1; }
      std::vector<uint8_t> isb(sp->ne, 0);
      for (int64_t e = 0; e < sp->ne; e++)
         for (int l = 0; l < sp->nd; l++)
            if (shared[sp->gather[(size_t)e * sp->nd + l]]) { isb[e] = 1; break; }
      sp->elem_perm.clear();
      for (int64_t e = 0; e < sp->ne; e++) if (isb[e]) { sp->elem_perm.push_back(e); }
      sp->n_bdr_elems = (int64_t)sp->elem_perm.size();
      for (int64_t e = 0; e < sp->ne; e++) if (!isb[e]) { sp->elem_perm.push_back(e); }
      std::vector<int32_t> g2(sp->gather.size());
      for (int64_t e = 0; e < sp->ne; e++)
         std::memcpy(&g2[(size_t)e * sp->nd], &sp->gather[(size_t)sp->elem_perm[e] * sp->nd], sizeof(int32_t) * sp->nd);
      sp->gather.swap(g2);
   }
   sp->dof_global.assign(sp->ndof, 0);
   for (int64_t g = 0; g < sp->ndof; g++) { sp->dof_global[newid[g]] = key[g]; }
   std::sort(shares.begin(), shares.end(), [](const PShare &a, const PShare &b)
   { return a.peer != b.peer ? a.peer < b.peer : a.key < b.key; });
   for (const PShare &s : shares)
   {
      if (sp->peers.empty() || sp->peers.back().rank != s.peer) { sp->peers.emplace_back(); sp->peers.back().rank = s.peer; }
      (s.mine ? sp->peers.back().own_idx : sp->peers.back().ghost_idx).push_back(newid[s.dof]);
   }
   // symmetric plan: per peer the dofs shared with it in key order (both sides of a pair build the same list);
   // per shared dof the contributions in ascending rank order, the own value at this rank's position
   std::sort(sym.begin(), sym.end(), [](const PSymShare &a, const PSymShare &b)
   { return a.peer != b.peer ? a.peer < b.peer : a.key < b.key; });
   cdm_sym_plan &sy = sp->sym;
   for (const PSymShare &s : sym)
   {
      if (sy.peers.empty() || sy.peers.back().rank != s.peer)
      { sy.peers.emplace_back(); sy.peers.back().rank = s.peer; sy.peers.back().off = (int64_t)sy.all.size(); }
      sy.peers.back().idx.push_back(newid[s.dof]);
      sy.all.push_back(newid[s.dof]);
   }
   {
      // entries of `all` grouped by dof; within a dof they are already in ascending peer rank (peer-major order)
      std::vector<int32_t> order(sy.all.size());
      for (size_t i = 0; i < order.size(); i++) { order[i] = (int32_t)i; }
      std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return sy.all[a] < sy.all[b]; });
      std::vector<int> peer_of(sy.all.size());
      for (const cdm_sym_peer &pr : sy.peers) for (size_t i = 0; i < pr.idx.size(); i++) { peer_of[pr.off + i] = pr.rank; }
      size_t i = 0;
      while (i < order.size())
      {
         size_t j = i;
         while (j < order.size() && sy.all[order[j]] == sy.all[order[i]]) { j++; }
         sy.sh_dof.push_back(sy.all[order[i]]);
         sy.sh_off.push_back((int32_t)sy.sh_src.size());
         bool own_done = false;
         for (size_t k = i; k < j; k++)
         {
            if (!own_done && peer_of[order[k]] > me) { sy.sh_src.push_back(-1); own_done = true; }
            sy.sh_src.push_back(order[k]);
         }
         if (!own_done) { sy.sh_src.push_back(-1); }
         i = j;
      }
      sy.sh_off.push_back((int32_t)sy.sh_src.size());
   }
}

// Element-wise partition (cdm_mesh_partition_elements): the parent mesh is replicated, so every rank numbers the GLOBAL space
// itself (same routine, same ids everywhere) and reads off, for each of its dofs, the global id (the key both sides of a pair
// sort by) and the set of ranks whose elements contain it.  No setup communication, like the box partition.
static int build_partition_general(const cdm_mesh *m, cdm_space *sp)
{
   const cdm_part_info &pi = *m->pinfo;
   const int nd = sp->nd, me = m->rank;
   std::vector<int32_t> gg, off, flat;
   int64_t ndof_g;
   {
      cdm_mesh parent;
      parent.geom = 0; parent.dim = pi.dim; parent.nv = pi.nv; parent.ne = pi.ne; parent.nbe = 0; parent.ev = pi.ev;
      ndof_g = cdm_host_h1_numbering(parent, sp->p, gg, off, flat);
   }
   if (ndof_g < 0 || ndof_g > 2147483000LL) { return -1; }
   std::vector<int64_t> key(sp->ndof, -1);
   for (int64_t e = 0; e < sp->ne; e++)
      for (int l = 0; l < nd; l++)
      {
         const int64_t k = gg[(size_t)m->eglobal[e] * nd + l];
         int64_t &slot = key[sp->gather[(size_t)e * nd + l]];
         if (slot >= 0 && slot != k) { return -1; }            // a local dof must be one global dof
         slot = k;
      }
   std::vector<int32_t> g2l((size_t)ndof_g, -1);
   for (int64_t d = 0; d < sp->ndof; d++)
   {
      if (key[d] < 0 || g2l[key[d]] >= 0) { return -1; }       // ... and the other way round
      g2l[key[d]] = (int32_t)d;
   }
   std::vector<uint64_t> mask(sp->ndof, 0);
   for (int64_t e = 0; e < pi.ne; e++)
   {
      const uint64_t bit = 1ull << pi.elem_rank[e];
      for (int l = 0; l < nd; l++) { const int32_t d = g2l[gg[(size_t)e * nd + l]]; if (d >= 0) { mask[d] |= bit; } }
   }
   std::vector<int> owner(sp->ndof, me);
   std::vector<PShare> shares;
   std::vector<PSymShare> sym;
   for (int64_t d = 0; d < sp->ndof; d++)
   {
      const uint64_t mk = mask[d];
      if (!(mk >> me & 1ull)) { return -1; }
      const int own = __builtin_ctzll(mk);
      owner[d] = own;
      if ((mk & (mk - 1)) == 0) { continue; }                  // a single rank
      for (int r = 0; r < pi.nranks; r++)
      {
         if (r == me || !(mk >> r & 1ull)) { continue; }
         sym.push_back({key[d], (int32_t)d, r});
         if (own == me) { shares.push_back({key[d], (int32_t)d, r, true}); }
      }
      if (own != me) { shares.push_back({key[d], (int32_t)d, own, false}); }
   }
   finish_partition(sp, me, key, owner, shares, sym);
   return 0;
}

int cdm_space_create_h1(cdm_ctx *ctx, const cdm_mesh *mesh, int order, cdm_space **space)
{
   if (!ctx || !mesh || !space) { return cdm_fail(ctx, CDM_EINVAL, "cdm_space_create_h1: bad arguments"); }
   if (order < 1 || order > 6) { return cdm_fail(ctx, CDM_EUNSUP, "cdm_space_create_h1: order must be 1..6"); }
   cdm_space *sp = space_new(ctx, mesh, order);
   if (!sp) { return cdm_fail(ctx, CDM_ENOMEM, "out of memory"); }
   sp->class_off.assign(5, 0);
   if (mesh->geom == 1)
   {
      if (mesh->dim != 2) { delete sp; return cdm_fail(ctx, CDM_EUNSUP, "cdm_space_create_h1: simplex meshes are triangles only"); }
      sp->ndof = cdm_host_h1_numbering_tri(*mesh, order, sp->gather, sp->bdr_dofs_off, sp->bdr_dofs_flat);
      sp->class_off.clear();
   }
   else
   sp->ndof = cdm_host_h1_numbering(*mesh, order, sp->gather, sp->bdr_dofs_off, sp->bdr_dofs_flat, sp->class_off.data());
   if (sp->ndof < 0) { delete sp; return cdm_fail(ctx, CDM_EINVAL, "cdm_space_create_h1: boundary element not found in mesh"); }
   if (sp->ndof > 2147483000LL) { delete sp; return cdm_fail(ctx, CDM_EUNSUP, "cdm_space_create_h1: more than 2^31 dofs on one rank"); }
   sp->ntrue = sp->ndof;
   if (mesh->is_part && mesh->pinfo)
   {
      if (mesh->pinfo->nranks > 1)
      {
         if (build_partition_general(mesh, sp)) { delete sp; return cdm_fail(ctx, CDM_EINVAL, "cdm_space_create_h1: inconsistent element-wise partition"); }
         sp->class_off.clear();
      }
   }
   else if (mesh->is_part && mesh->parts[0] * mesh->parts[1] * mesh->parts[2] > 1) { build_partition(mesh, sp); sp->class_off.clear(); }
   int rc = space_finish(ctx, mesh, sp);
   if (rc) { cdm_space_destroy(sp); return rc; }
   *space = sp;
   return CDM_OK;
}

int cdm_space_create_from_table(cdm_ctx *ctx, const cdm_mesh *mesh, int order, int64_t ndof,
                                const int32_t *elem_dof, cdm_space **space)
{
   if (!ctx || !mesh || !space || !elem_dof || ndof < 1) { return cdm_fail(ctx, CDM_EINVAL, "cdm_space_create_from_table: bad arguments"); }
   if (order < 1 || order > 6) { return cdm_fail(ctx, CDM_EUNSUP, "cdm_space_create_from_table: order must be 1..6"); }
   if (mesh->geom != 0) { return cdm_fail(ctx, CDM_EUNSUP, "cdm_space_create_from_table: tensor-product meshes only"); }
   cdm_space *sp = space_new(ctx, mesh, order);
   if (!sp) { return cdm_fail(ctx, CDM_ENOMEM, "out of memory"); }
   sp->ndof = sp->ntrue = ndof;
   sp->gather.assign(elem_dof, elem_dof + sp->ne * sp->nd);
   for (int32_t g : sp->gather)
      if (g < 0 || g >= ndof) { delete sp; return cdm_fail(ctx, CDM_EINVAL, "cdm_space_create_from_table: dof id out of range"); }
   // boundary dofs: use our own numbering only to locate which lexicographic
   // nodes lie on each boundary element, then translate through the table
   {
      std::vector<int32_t> own_gather, off, flat;
      const int64_t nd_own = cdm_host_h1_numbering(*mesh, order, own_gather, off, flat);
      if (nd_own < 0) { delete sp; return cdm_fail(ctx, CDM_EINVAL, "cdm_space_create_from_table: boundary element not found in mesh"); }
      std::vector<int32_t> tr(nd_own, -1);
      for (size_t i = 0; i < own_gather.size(); i++) { tr[own_gather[i]] = sp->gather[i]; }
      for (auto &g : flat) { g = tr[g]; }
      sp->bdr_dofs_off.swap(off); sp->bdr_dofs_flat.swap(flat);
   }
   int rc = space_finish(ctx, mesh, sp);
   if (rc) { cdm_space_destroy(sp); return rc; }
   *space = sp;
   return CDM_OK;
}

int cdm_space_sizes(const cdm_space *sp, int *dim, int *order, int64_t *ne, int64_t *ndof, int *d1d,
                    int *q1d, int64_t *ntrue)
{
   if (!sp) { return CDM_EINVAL; }
   if (dim) { *dim = sp->dim; }
   if (order) { *order = sp->p; }
   if (ne) { *ne = sp->ne; }
   if (ndof) { *ndof = sp->ndof; }
   if (d1d) { *d1d = sp->d1d; }
   if (q1d) { *q1d = sp->q1d; }
   if (ntrue) { *ntrue = sp->ntrue; }
   return CDM_OK;
}

int cdm_space_get_maps(const cdm_space *sp, int32_t *gather_map, int32_t *offsets, int32_t *indices)
{
   if (!sp) { return CDM_EINVAL; }
   if (gather_map) { std::memcpy(gather_map, sp->gather.data(), sp->gather.size() * sizeof(int32_t)); }
   if (offsets) { std::memcpy(offsets, sp->offsets.data(), sp->offsets.size() * sizeof(int32_t)); }
   if (indices) { std::memcpy(indices, sp->indices.data(), sp->indices.size() * sizeof(int32_t)); }
   return CDM_OK;
}

int cdm_space_essential_dofs(const cdm_space *sp, const int32_t *bdr_marker, int nattr, int32_t *list, int64_t *count)
{
   if (!sp || !bdr_marker || !count) { return CDM_EINVAL; }
   std::vector<uint8_t> mark(sp->ndof, 0);
   for (int64_t b = 0; b < sp->nbe; b++)
   {
      const int a = sp->bdr_attr[b];
      if (a < 1 || a > nattr || !bdr_marker[a - 1]) { continue; }
      for (int32_t j = sp->bdr_dofs_off[b]; j < sp->bdr_dofs_off[b + 1]; j++) { mark[sp->bdr_dofs_flat[j]] = 1; }
   }
   int64_t n = 0;
   for (int64_t g = 0; g < sp->ndof; g++) if (mark[g]) { if (list) { list[n] = (int32_t)g; } n++; }
   *count = n;
   return CDM_OK;
}

int cdm_space_get_basis(const cdm_space *sp, double *B, double *G, double *qw, double *nodes)
{
   if (!sp) { return CDM_EINVAL; }
   if (B) { std::memcpy(B, sp->B.data(), sp->B.size() * sizeof(double)); }
   if (G) { std::memcpy(G, sp->G.data(), sp->G.size() * sizeof(double)); }
   if (qw) { std::memcpy(qw, sp->qw.data(), sp->qw.size() * sizeof(double)); }
   if (nodes) { std::memcpy(nodes, sp->nodes.data(), sp->nodes.size() * sizeof(double)); }
   return CDM_OK;
}

// trilinear / bilinear map of reference point xi through element e
static void map_point(const cdm_space *sp, int64_t e, const double *xi, double *X)
{
   const int dim = sp->dim, nvpe = (sp->geom == 1) ? 3 : ((dim == 2) ? 4 : 8);
   const double x = xi[0], y = xi[1], z = (dim == 3) ? xi[2] : 0.0;
   double N[8];
   if (sp->geom == 1) { N[0] = 1.0 - x - y; N[1] = x; N[2] = y; }
   else if (dim == 2) { N[0] = (1 - x) * (1 - y); N[1] = x * (1 - y); N[2] = x * y; N[3] = (1 - x) * y; }
   else
   {
      N[0] = (1 - x) * (1 - y) * (1 - z); N[1] = x * (1 - y) * (1 - z); N[2] = x * y * (1 - z); N[3] = (1 - x) * y * (1 - z);
      N[4] = (1 - x) * (1 - y) * z; N[5] = x * (1 - y) * z; N[6] = x * y * z; N[7] = (1 - x) * y * z;
   }
   for (int c = 0; c < dim; c++)
   {
      double s = 0.0;
      for (int k = 0; k < nvpe; k++) { s += N[k] * sp->elem_x[((size_t)e * nvpe + k) * dim + c]; }
      X[c] = s;
   }
}

int cdm_space_dof_coords(const cdm_space *sp, double *xyz)
{
   if (!sp || !xyz) { return CDM_EINVAL; }
   const int d1d = sp->d1d, dim = sp->dim;
   for (int64_t e = 0; e < sp->ne; e++)
      for (int l = 0; l < sp->nd; l++)
      {
         double xi[3] = {0.0, 0.0, 0.0};
         if (sp->geom == 1) { xi[0] = sp->snodes[2 * l]; xi[1] = sp->snodes[2 * l + 1]; }
         else { xi[0] = sp->nodes[l % d1d]; xi[1] = sp->nodes[(l / d1d) % d1d]; xi[2] = dim == 3 ? sp->nodes[l / (d1d * d1d)] : 0.0; }
         map_point(sp, e, xi, xyz + (size_t)sp->gather[(size_t)e * sp->nd + l] * dim);
      }
   return CDM_OK;
}

int cdm_space_qpt_coords(const cdm_space *sp, double *xyz)
{
   if (!sp || !xyz) { return CDM_EINVAL; }
   const int q1d = sp->q1d, dim = sp->dim;
   for (int64_t e = 0; e < sp->ne; e++)
      for (int q = 0; q < sp->nq; q++)
      {
         double xi[3] = {0.0, 0.0, 0.0};
         if (sp->geom == 1) { xi[0] = sp->sqx[2 * q]; xi[1] = sp->sqx[2 * q + 1]; }
         else { xi[0] = sp->qx[q % q1d]; xi[1] = sp->qx[(q / q1d) % q1d]; xi[2] = dim == 3 ? sp->qx[q / (q1d * q1d)] : 0.0; }
         map_point(sp, e, xi, xyz + ((size_t)e * sp->nq + q) * dim);
      }
   return CDM_OK;
}

int cdm_space_halo_peers(const cdm_space *sp, int *npeers)
{
   if (!sp || !npeers) { return CDM_EINVAL; }
   *npeers = (int)sp->peers.size();
   return CDM_OK;
}

int cdm_space_halo_peer(const cdm_space *sp, int i, int *rank, int64_t *n_own, int64_t *n_ghost,
                        int32_t *own_idx, int32_t *ghost_idx)
{
   if (!sp || i < 0 || i >= (int)sp->peers.size()) { return CDM_EINVAL; }
   const cdm_halo_peer &pr = sp->peers[i];
   if (rank) { *rank = pr.rank; }
   if (n_own) { *n_own = (int64_t)pr.own_idx.size(); }
   if (n_ghost) { *n_ghost = (int64_t)pr.ghost_idx.size(); }
   if (own_idx && !pr.own_idx.empty()) { std::memcpy(own_idx, pr.own_idx.data(), pr.own_idx.size() * sizeof(int32_t)); }
   if (ghost_idx && !pr.ghost_idx.empty()) { std::memcpy(ghost_idx, pr.ghost_idx.data(), pr.ghost_idx.size() * sizeof(int32_t)); }
   return CDM_OK;
}

int cdm_space_sym_peers(const cdm_space *sp, int *npeers, int64_t *n_shared, int64_t *n_contrib)
{
   if (!sp) { return CDM_EINVAL; }
   if (npeers) { *npeers = (int)sp->sym.peers.size(); }
   if (n_shared) { *n_shared = (int64_t)sp->sym.sh_dof.size(); }
   if (n_contrib) { *n_contrib = (int64_t)sp->sym.sh_src.size(); }
   return CDM_OK;
}

int cdm_space_sym_peer(const cdm_space *sp, int i, int *rank, int64_t *n, int64_t *offset, int32_t *idx)
{
   if (!sp || i < 0 || i >= (int)sp->sym.peers.size()) { return CDM_EINVAL; }
   const cdm_sym_peer &pr = sp->sym.peers[i];
   if (rank) { *rank = pr.rank; }
   if (n) { *n = (int64_t)pr.idx.size(); }
   if (offset) { *offset = pr.off; }
   if (idx && !pr.idx.empty()) { std::memcpy(idx, pr.idx.data(), pr.idx.size() * sizeof(int32_t)); }
   return CDM_OK;
}

int cdm_space_sym_sum_plan(const cdm_space *sp, int32_t *dof, int32_t *off, int32_t *src)
{
   if (!sp) { return CDM_EINVAL; }
   const cdm_sym_plan &sy = sp->sym;
   if (dof && !sy.sh_dof.empty()) { std::memcpy(dof, sy.sh_dof.data(), sy.sh_dof.size() * sizeof(int32_t)); }
   if (off && !sy.sh_off.empty()) { std::memcpy(off, sy.sh_off.data(), sy.sh_off.size() * sizeof(int32_t)); }
   if (src && !sy.sh_src.empty()) { std::memcpy(src, sy.sh_src.data(), sy.sh_src.size() * sizeof(int32_t)); }
   return CDM_OK;
}

int cdm_space_elem_perm(const cdm_space *sp, int64_t *perm, int64_t *n_boundary)
{
   if (!sp) { return CDM_EINVAL; }
   if (perm) { for (int64_t e = 0; e < sp->ne; e++) { perm[e] = sp->elem_perm.empty() ? e : sp->elem_perm[e]; } }
   if (n_boundary) { *n_boundary = sp->n_bdr_elems; }
   return CDM_OK;
}

int cdm_space_dof_global(const cdm_space *sp, int64_t *keys)
{
   if (!sp || !keys) { return CDM_EINVAL; }
   if (sp->dof_global.empty()) { for (int64_t g = 0; g < sp->ndof; g++) { keys[g] = g; } }
   else { std::memcpy(keys, sp->dof_global.data(), sp->dof_global.size() * sizeof(int64_t)); }
   return CDM_OK;
}

int cdm_space_destroy(cdm_space *sp)
{
   if (!sp) { return CDM_OK; }
   if (sp->ctx && sp->ctx->device >= 0)
   {
      cudaFree(sp->gather_dev); cudaFree(sp->offsets_dev); cudaFree(sp->indices_dev); cudaFree(sp->elem_x_dev);
      cudaFree(sp->work_dev); cudaFree(sp->elem_part_dev); cudaFree(sp->iota_dev);
      cudaFree(sp->sB_dev); cudaFree(sp->sG_dev); cudaFree(sp->sqw_dev); cudaFree(sp->sqx_dev);
      cdm_halo_p2p_destroy(sp);
      cdm_halo_sym_destroy(sp);
      cudaFree(sp->sym.all_dev); cudaFree(sp->sym.sh_dof_dev); cudaFree(sp->sym.sh_off_dev); cudaFree(sp->sym.sh_src_dev);
      cdm_halo_plan &hp = sp->halo;
      cudaFree(hp.own_all_dev); cudaFree(hp.ghost_all_dev); cudaFree(hp.pt_dof_dev); cudaFree(hp.pt_off_dev);
      cudaFree(hp.pt_src_dev); cudaFree(hp.send_dev); cudaFree(hp.recv_dev);
   }
   delete sp;
   return CDM_OK;
}

// --------------------------------------------------------------- operator

static int check_coeff(const cdm_coeff *c, int dim, int which)
{
   if (!c || c->kind == CDM_COEFF_NONE) { return 0; }
   if (c->kind != CDM_COEFF_CONST && c->kind != CDM_COEFF_QPT) { return -1; }
   if (!c->data) { return -1; }
   if (which == 0 && c->ncomp != 1 && c->ncomp != dim * (dim + 1) / 2) { return -1; }
   if (which == 1 && c->ncomp != dim) { return -1; }
   if (which == 2 && c->ncomp != 1) { return -1; }
   return 1;
}

int cdm_operator_create(cdm_space *sp, const cdm_coeff *kappa, const cdm_coeff *vel, double conv_alpha,
                        const cdm_coeff *mass, const int32_t *ess_dofs, int64_t n_ess, cdm_op **out)
{
   if (!sp || !out) { return CDM_EINVAL; }
   cdm_ctx *ctx = sp->ctx;
   CDM_REQUIRE_GPU(ctx);
   const int hk = check_coeff(kappa, sp->dim, 0), hv = check_coeff(vel, sp->dim, 1), hm = check_coeff(mass, sp->dim, 2);
   if (hk < 0 || hv < 0 || hm < 0) { return cdm_fail(ctx, CDM_EINVAL, "cdm_operator_create: malformed coefficient"); }
   if (!hk && !hv && !hm) { return cdm_fail(ctx, CDM_EINVAL, "cdm_operator_create: no integrator"); }
   if (n_ess < 0 || (n_ess > 0 && !ess_dofs)) { return cdm_fail(ctx, CDM_EINVAL, "cdm_operator_create: bad essential dof list"); }
   cdm_op *op = new (std::nothrow) cdm_op;
   if (!op) { return cdm_fail(ctx, CDM_ENOMEM, "out of memory"); }
   op->sp = sp;
   op->has_diff = hk; op->has_conv = hv; op->has_mass = hm;
   const int nsym = sp->dim * (sp->dim + 1) / 2, q2 = sp->q1d * sp->q1d;
   op->ncomp = (hk ? nsym : 0) + (hv ? sp->dim : 0) + (hm ? 1 : 0);
   op->slab = (op->ncomp * q2 + 1) & ~1;
   op->D_len = sp->ne * (int64_t)op->slab * (sp->dim == 3 ? sp->q1d : 1);
   if (sp->geom == 1) { op->slab = (op->ncomp * sp->nq + 1) & ~1; op->D_len = sp->ne * (int64_t)op->slab; }
   int rc = CDM_OK;
   do
   {
      if (cudaMalloc(&op->D_dev, (size_t)op->D_len * sizeof(double)) != cudaSuccess)
      { cudaGetLastError(); rc = cdm_fail(ctx, CDM_ENOMEM, "cdm_operator_create: cannot allocate quadrature data"); break; }
      if (cudaMemsetAsync(op->D_dev, 0, (size_t)op->D_len * sizeof(double), ctx->stream) != cudaSuccess) { rc = CDM_ECUDA; break; }
      if (n_ess > 0)
      {
         std::vector<uint8_t> mark(sp->ndof, 0);
         for (int64_t i = 0; i < n_ess; i++)
         {
            if (ess_dofs[i] < 0 || ess_dofs[i] >= sp->ndof) { rc = cdm_fail(ctx, CDM_EINVAL, "cdm_operator_create: essential dof out of range"); break; }
            mark[ess_dofs[i]] = 1;
         }
         if (rc) { break; }
         std::vector<int32_t> gc(sp->gather);
         for (auto &g : gc) if (mark[g]) { g = -1 - g; }
         for (int64_t g = 0; g < sp->ntrue; g++) if (mark[g]) { op->ess_host.push_back((int32_t)g); }
         op->n_ess = (int64_t)op->ess_host.size();
         for (int64_t g = sp->ntrue; g < sp->ndof; g++) if (mark[g]) { op->ess_host.push_back((int32_t)g); }
         op->n_ess_all = (int64_t)op->ess_host.size();
         if ((rc = upload(ctx, gc, &op->gather_c_dev))) { break; }
         if ((rc = upload(ctx, op->ess_host, &op->ess_dev))) { break; }
         if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { rc = CDM_ECUDA; break; }
      }
      rc = (sp->geom == 1) ? cdm_k_setup_qdata_simplex(op, kappa, vel, conv_alpha, mass) : cdm_k_setup_qdata(op, kappa, vel, conv_alpha, mass);
      // triangles have no sum-factorised apply: the operator IS the assembled sparse matrix (the reference's own path)
      if (!rc && sp->geom == 1) { rc = cdm_operator_assemble_csr(op); if (!rc) { op->assembly = 1; } }
   } while (0);
   if (rc) { cdm_operator_destroy(op); return rc; }
   // default kernel per order (measured, profiles/): p=3 hand-specialised register-z kernel,
   // p>=4 the generic group kernel, p<=2 and 2D the block kernel
   // 3D defaults: sub-warp kernel (5) for orders 1-2, warp kernel (3) for order 3, group kernel (4) above
   op->kernel_variant = (sp->dim != 3) ? (sp->p <= 4 ? 6 : 0) : (sp->p <= 2 ? 5 : (sp->p == 3 ? 3 : 4));
   *out = op;
   return CDM_OK;
}

int cdm_operator_update(cdm_op *op, const cdm_coeff *kappa, const cdm_coeff *vel, double conv_alpha, const cdm_coeff *mass)
{
   if (!op) { return CDM_EINVAL; }
   cdm_ctx *ctx = op->sp->ctx;
   CDM_REQUIRE_GPU(ctx);
   const int hk = check_coeff(kappa, op->sp->dim, 0), hv = check_coeff(vel, op->sp->dim, 1), hm = check_coeff(mass, op->sp->dim, 2);
   if (hk < 0 || hv < 0 || hm < 0 || (hk > 0) != op->has_diff || (hv > 0) != op->has_conv || (hm > 0) != op->has_mass)
      return cdm_fail(ctx, CDM_EINVAL, "cdm_operator_update: the set of integrators must not change");
   if (op->dinv_dev) { cudaFree(op->dinv_dev); op->dinv_dev = nullptr; }
   const int rc = (op->sp->geom == 1) ? cdm_k_setup_qdata_simplex(op, kappa, vel, conv_alpha, mass)
                                      : cdm_k_setup_qdata(op, kappa, vel, conv_alpha, mass);
   return rc ? rc : cdm_csr_refill_if_present(op);        // an assembled matrix follows the new coefficients
}

static void destroy_host_pipe(cdm_op *op);

int cdm_operator_destroy(cdm_op *op)
{
   if (!op) { return CDM_OK; }
   if (op->sp && op->sp->ctx && op->sp->ctx->stream) { cudaStreamSynchronize(op->sp->ctx->stream); }
   cudaFree(op->D_dev); cudaFree(op->gather_c_dev); cudaFree(op->ess_dev); cudaFree(op->yE_dev);
   cudaFree(op->xL_dev); cudaFree(op->yL_dev); cudaFree(op->dinv_dev); cudaFree(op->kry_dev);
   for (int i = 0; i < 3; i++) { cudaFree(op->coef_dev[i]); }
   destroy_host_pipe(op);
   cdm_csr_destroy(op);
   delete op;
   return CDM_OK;
}

int64_t cdm_operator_size(const cdm_op *op) { return op ? op->sp->ntrue : 0; }
int64_t cdm_operator_local_size(const cdm_op *op) { return op ? op->sp->ndof : 0; }

int cdm_operator_set_option(cdm_op *op, const char *name, int value)
{
   if (!op || !name) { return CDM_EINVAL; }
   if (!std::strcmp(name, "scatter")) { if (value != 0 && value != 1) { return CDM_EINVAL; } op->scatter_mode = value; return CDM_OK; }
   if (!std::strcmp(name, "kernel")) { op->kernel_variant = value; return CDM_OK; }
   if (!std::strcmp(name, "assembly"))
   {
      // 1: the reference's literal path -- apply = SpMV with the fully assembled CSR matrix (csr_path.cu)
      if (value != 0 && value != 1) { return CDM_EINVAL; }
      if (value == 0 && op->sp->geom == 1) { return cdm_fail(op->sp->ctx, CDM_EUNSUP, "triangle spaces have no matrix-free apply (assembly must stay 1)"); }
      if (value == 1 && !op->csr) { const int rc = cdm_operator_assemble_csr(op); if (rc) { return rc; } }
      op->assembly = value;
      return CDM_OK;
   }
   if (!std::strcmp(name, "halo"))
   {
      // 1: exchange the shared dofs through peer memory (collective: every rank must make this call)
      // 2: symmetric exchange (its setup is collective and happens at the first apply)
      if (value < 0 || value > 2) { return CDM_EINVAL; }
      if (value == 1) { const int rc = cdm_halo_p2p_setup(op->sp); if (rc) { return rc; } }
      op->halo_mode = value;
      return CDM_OK;
   }
   if (!std::strcmp(name, "tail")) { op->tail = value != 0; return CDM_OK; }
   if (!std::strcmp(name, "ghost_in")) { op->ghost_in = value != 0; return CDM_OK; }
   if (!std::strcmp(name, "allreduce")) { if (value != 0 && value != 1) { return CDM_EINVAL; } op->sp->ctx->allreduce_mode = value; return CDM_OK; }
   if (!std::strcmp(name, "grid_cap")) { if (value < 0) { return CDM_EINVAL; } op->grid_cap = value; return CDM_OK; }
   if (!std::strcmp(name, "host_pipeline")) { op->host_pipeline = value; destroy_host_pipe(op); return CDM_OK; }
   if (!std::strcmp(name, "host_pipeline_shape")) { if (value != 0 && value != 1) { return CDM_EINVAL; } op->host_pipeline_shape = value; destroy_host_pipe(op); return CDM_OK; }
   if (!std::strcmp(name, "overlap")) { if (value < 0 || value > 2) { return CDM_EINVAL; } op->overlap = value; return CDM_OK; }
   if (!std::strcmp(name, "ilu_sweep")) { if (value != 0 && value != 1) { return CDM_EINVAL; } op->ilu_sweep = value; return CDM_OK; }
   return cdm_fail(op->sp->ctx, CDM_EINVAL, std::string("unknown option ") + name);
}

int cdm_operator_get_option(const cdm_op *op, const char *name, int *value)
{
   if (!op || !name || !value) { return CDM_EINVAL; }
   if (!std::strcmp(name, "scatter")) { *value = op->scatter_mode; return CDM_OK; }
   if (!std::strcmp(name, "kernel")) { *value = op->kernel_variant; return CDM_OK; }
   if (!std::strcmp(name, "assembly")) { *value = op->assembly; return CDM_OK; }
   if (!std::strcmp(name, "overlap")) { *value = op->overlap; return CDM_OK; }
   if (!std::strcmp(name, "ilu_sweep")) { *value = op->ilu_sweep; return CDM_OK; }
   if (!std::strcmp(name, "tail")) { *value = op->tail ? 1 : 0; return CDM_OK; }
   if (!std::strcmp(name, "ghost_in")) { *value = op->ghost_in ? 1 : 0; return CDM_OK; }
   // the protocol actually in use: 2 falls back to 0 when the peer-memory setup failed on any rank
   if (!std::strcmp(name, "halo")) { *value = (op->halo_mode == 2 && op->sp->sym.ready == -1) ? 0 : op->halo_mode; return CDM_OK; }
   if (!std::strcmp(name, "allreduce")) { *value = (op->sp->ctx->allreduce_mode && op->sp->ctx->red_sym) ? 1 : 0; return CDM_OK; }
   return cdm_fail(op->sp->ctx, CDM_EINVAL, std::string("unknown option ") + name);
}

static int ensure_L(cdm_op *op)
{
   cdm_ctx *ctx = op->sp->ctx;
   if (!op->xL_dev) { CDM_CUDA(ctx, cudaMalloc(&op->xL_dev, sizeof(double) * (size_t)op->sp->ndof)); }
   if (!op->yL_dev) { CDM_CUDA(ctx, cudaMalloc(&op->yL_dev, sizeof(double) * (size_t)op->sp->ndof)); }
   return CDM_OK;
}

// Apply on buffers that have room for the ghost tail (length >= ndof): used by the
// Krylov drivers so that no T<->L copies are needed.
//   x_ghost_valid : the ghost tail of x already holds the owners' values (a vector produced by a previous
//                   apply in the symmetric mode, or by cdm_prolongate): the P exchange is skipped.
//                   Otherwise x's tail is overwritten by P.
// Multi-GPU protocols (option "halo"): 2 (default) symmetric peer-memory exchange: ONE exchange per apply, y comes
// back consistent on its ghost tail as well; 0 / 1: P before and P^T after the element kernel, over NCCL
// send/recv or peer-memory stores (y's ghost tail is then NOT valid).
int cdm_apply_tail(cdm_op *op, double *x_buf, double *y_buf, bool constrained, bool x_ghost_valid)
{
   cdm_space *sp = op->sp;
   cdm_ctx *ctx = sp->ctx;
   int rc;
   const bool par = ctx->nranks > 1 && !sp->peers.empty();
   if (par && op->halo_mode == 2 && op->assembly == 0)
   {
      if (sp->sym.ready == 0 && (rc = cdm_halo_sym_setup(sp))) { return rc; }      // collective, first apply only
      if (sp->sym.ready == 1)
      {
         if (!x_ghost_valid && (rc = cdm_halo_P(op, x_buf))) { return rc; }
         cudaStream_t C = ctx->stream, H = ctx->stream_halo;
         const int64_t nb = sp->n_bdr_elems;
         // overlapped schedule: boundary elements (the only ones that touch shared dofs) first, then the exchange
         // on the high-priority stream beside the interior elements:
         //    C: memset y | boundary | interior ...................| wait H | y[ess] = x[ess]
         //    H:                     | pack -> NVLink -> wait, add |
         if (op->overlap != 0 && H && op->scatter_mode == 1 && cdm_k_range_capable(op) && nb > 0 && nb < sp->ne)
         {
            cudaEvent_t *ev = ctx->ev_h;
            CDM_CUDA(ctx, cudaMemsetAsync(y_buf, 0, sizeof(double) * (size_t)sp->ndof, C));
            op->range_on = true;
            op->e_begin = 0; op->e_end = nb;
            rc = cdm_k_apply(op, x_buf, y_buf, constrained);
            if (!rc)
            {
               cudaEventRecord(ev[0], C);
               cudaStreamWaitEvent(H, ev[0], 0);
               rc = cdm_halo_sym_exchange(sp, y_buf, H);
               cudaEventRecord(ev[1], H);
            }
            if (!rc)
            {
               op->e_begin = nb; op->e_end = sp->ne;
               rc = cdm_k_apply(op, x_buf, y_buf, constrained);
               cudaStreamWaitEvent(C, ev[1], 0);
            }
            op->range_on = false;
            if (rc) { return rc; }
         }
         else
         {
            if ((rc = cdm_k_apply(op, x_buf, y_buf, constrained))) { return rc; }
            if ((rc = cdm_halo_sym_exchange(sp, y_buf, C))) { return rc; }
         }
         if (constrained && op->n_ess_all > 0) { rc = cdm_k_copy_idx(ctx, op->n_ess_all, op->ess_dev, x_buf, y_buf); }
         return rc;
      }
   }
   // the P / P^T pair of an apply may go through peer memory (option "halo" = 1); any other exchange uses NCCL
   struct P2PScope
   {
      cdm_halo_plan &hp;
      P2PScope(cdm_halo_plan &h, bool on) : hp(h) { hp.p2p_active = on; }
      ~P2PScope() { hp.p2p_active = false; }
   } p2p_scope(sp->halo, par && op->halo_mode == 1 && sp->halo.p2p != nullptr);
   // Overlapped schedule (atomic scatter, range-capable kernels): the halo exchange runs on a
   // high-priority stream with its own communicator while the compute stream works on interior
   // elements:   H: pack, P exchange, unpack      | C: memset y, interior A
   //             C: boundary elements             |
   //             H: pack, P^T exchange, unpack-add | C: interior B
   // Each interior launch waits for the pack kernel of its phase, so the NCCL kernel (ready at the
   // same moment, higher priority) is placed before the persistent element kernel fills the SMs.
   // Measured (profiles/r01_scaling.md): faster with 1 neighbour, slower with 7 (the NCCL kernel then competes
   // with the persistent element kernel for SM slots), so by default only used up to 3 neighbours;
   // option "overlap" = 2 forces it.
   if (par && (op->overlap == 2 || (op->overlap == 1 && sp->peers.size() <= 3)) && ctx->stream_halo &&
       (op->halo_mode == 1 || ctx->comm_halo) &&
       op->scatter_mode == 1 && cdm_k_range_capable(op) && sp->n_bdr_elems > 0 && sp->n_bdr_elems < sp->ne)
   {
      cudaStream_t C = ctx->stream, H = ctx->stream_halo;
      cudaEvent_t *ev = ctx->ev_h;
      const int64_t nb = sp->n_bdr_elems, mid = nb + (sp->ne - nb) / 2;
      CDM_CUDA(ctx, cudaEventRecord(ev[0], C));                       // x is ready
      CDM_CUDA(ctx, cudaStreamWaitEvent(H, ev[0], 0));
      if ((rc = cdm_halo_P_async(op, x_buf, ev[1]))) { return rc; }   // H: pack (ev[1]), exchange, unpack
      CDM_CUDA(ctx, cudaEventRecord(ev[2], H));
      CDM_CUDA(ctx, cudaMemsetAsync(y_buf, 0, sizeof(double) * (size_t)sp->ndof, C));
      CDM_CUDA(ctx, cudaStreamWaitEvent(C, ev[1], 0));
      op->range_on = true;
      op->e_begin = nb; op->e_end = mid;
      rc = cdm_k_apply(op, x_buf, y_buf, constrained);                // C: interior A
      if (!rc)
      {
         cudaStreamWaitEvent(C, ev[2], 0);                            // ghosts of x have arrived
         op->e_begin = 0; op->e_end = nb;
         rc = cdm_k_apply(op, x_buf, y_buf, constrained);             // C: boundary elements
      }
      if (!rc)
      {
         cudaEventRecord(ev[3], C);
         cudaStreamWaitEvent(H, ev[3], 0);
         rc = cdm_halo_PT_async(op, y_buf, ev[4]);                    // H: pack (ev[4]), exchange, unpack-add
         cudaEventRecord(ev[5], H);
      }
      if (!rc)
      {
         cudaStreamWaitEvent(C, ev[4], 0);
         op->e_begin = mid; op->e_end = sp->ne;
         rc = cdm_k_apply(op, x_buf, y_buf, constrained);             // C: interior B
         cudaStreamWaitEvent(C, ev[5], 0);
      }
      op->range_on = false;
      if (rc) { return rc; }
      if (constrained && op->n_ess > 0) { rc = cdm_k_copy_idx(ctx, op->n_ess, op->ess_dev, x_buf, y_buf); }
      return rc;
   }
   if (par && (rc = cdm_halo_P(op, x_buf))) { return rc; }
   if ((rc = cdm_k_apply(op, x_buf, y_buf, constrained))) { return rc; }
   if (par && (rc = cdm_halo_PT(op, y_buf))) { return rc; }
   if (constrained && op->n_ess > 0) { rc = cdm_k_copy_idx(ctx, op->n_ess, op->ess_dev, x_buf, y_buf); }
   return rc;
}

// does an apply on this operator leave y consistent on its ghost tail (so that y may be the next x without P)?
bool cdm_apply_keeps_ghosts(const cdm_op *op)
{
   const cdm_space *sp = op->sp;
   return sp->ctx->nranks > 1 && !sp->peers.empty() && op->halo_mode == 2 && op->assembly == 0 && sp->sym.ready == 1;
}

static int apply_T(cdm_op *op, const double *x, double *y, bool constrained)
{
   cdm_space *sp = op->sp;
   cdm_ctx *ctx = sp->ctx;
   if (sp->ntrue == sp->ndof || op->tail) { return cdm_apply_tail(op, const_cast<double *>(x), y, constrained, op->ghost_in != 0); }
   int rc = ensure_L(op); if (rc) { return rc; }
   CDM_CUDA(ctx, cudaMemcpyAsync(op->xL_dev, x, sizeof(double) * (size_t)sp->ntrue, cudaMemcpyDeviceToDevice, ctx->stream));
   if ((rc = cdm_apply_tail(op, op->xL_dev, op->yL_dev, constrained, false))) { return rc; }
   CDM_CUDA(ctx, cudaMemcpyAsync(y, op->yL_dev, sizeof(double) * (size_t)sp->ntrue, cudaMemcpyDeviceToDevice, ctx->stream));
   return CDM_OK;
}

int cdm_operator_apply(cdm_op *op, const double *x_dev, double *y_dev)
{
   if (!op || !x_dev || !y_dev) { return CDM_EINVAL; }
   if (x_dev == y_dev) { return cdm_fail(op->sp->ctx, CDM_EINVAL, "cdm_operator_apply: x and y must not alias"); }
   CDM_REQUIRE_GPU(op->sp->ctx);
   return apply_T(op, x_dev, y_dev, true);
}

int cdm_operator_apply_unconstrained(cdm_op *op, const double *x_dev, double *y_dev)
{
   if (!op || !x_dev || !y_dev) { return CDM_EINVAL; }
   if (x_dev == y_dev) { return cdm_fail(op->sp->ctx, CDM_EINVAL, "cdm_operator_apply_unconstrained: x and y must not alias"); }
   CDM_REQUIRE_GPU(op->sp->ctx);
   return apply_T(op, x_dev, y_dev, false);
}

// Plan of the pipelined host-vector apply.  Elements are cut into K chunks (in element order).  Because the
// library's numbering hands out ids in first-encounter order inside each entity class (vertices, edges, faces,
// interiors), the dofs first touched by chunk c form (nearly) a contiguous block at the front of the not yet
// uploaded part of each class, and the dofs whose last element lies in chunk c a block at the front of the not
// yet downloaded part.  ub[c+1][k] = one past the largest class-k dof touched by chunks <= c (upload bound),
// db[c+1][k] = the smallest class-k dof still touched by a chunk > c (download bound).
static void destroy_host_pipe(cdm_op *op)
{
   cdm_op::host_pipe &hp = op->pipe;
   if (hp.su) { cudaStreamDestroy(hp.su); hp.su = nullptr; }
   if (hp.sd) { cudaStreamDestroy(hp.sd); hp.sd = nullptr; }
   for (cudaEvent_t e : hp.ev_up) { if (e) { cudaEventDestroy(e); } }
   for (cudaEvent_t e : hp.ev_k) { if (e) { cudaEventDestroy(e); } }
   hp.ev_up.clear(); hp.ev_k.clear();
   cudaFree(hp.ess_by_chunk_dev); hp.ess_by_chunk_dev = nullptr;
   hp.K = 0;
}

static int build_host_pipe(cdm_op *op)
{
   cdm_space *sp = op->sp;
   cdm_ctx *ctx = sp->ctx;
   cdm_op::host_pipe &hp = op->pipe;
   // option "host_pipeline" > 1 selects the chunk count.  Round 1: 6 equal chunks, four copies (one per entity class) per
   // chunk and direction: 2.05 ms per step at config 2 (un-pipelined 2.83 ms).  The device timeline (CDM_PIPE_DEBUG)
   // showed both links busy end to end at 37 GB/s each -- a single large copy pair reaches 47 GB/s each -- plus 0.37 ms
   // before the first download and 0.22 ms after the last kernel.  So (option "host_pipeline_shape" = 1, default):
   // * tapered chunks (the first and the last are half as long): the head and the tail of the pipeline shrink;
   // * fewer, larger copies: vertices (4 % of the dofs) go up in one piece before the first chunk and come down in one
   //   piece after the last, edges (22 %) move every second chunk; faces and interiors every chunk.
   const int Kwant = op->host_pipeline > 1 ? std::min(op->host_pipeline, 64) : (op->host_pipeline_shape ? 8 : 6);
   const int K = (int)std::min<int64_t>(Kwant, std::max<int64_t>(1, sp->ne / 4096));
   const bool shaped = op->host_pipeline_shape != 0 && K >= 3;
   hp.eb.resize(K + 1);
   {
      std::vector<double> w(K, 2.0);
      if (shaped) { w[0] = w[K - 1] = 1.0; }
      double tot = 0.0, acc = 0.0;
      for (double v : w) { tot += v; }
      hp.eb[0] = 0;
      for (int c = 0; c < K; c++) { acc += w[c]; hp.eb[c + 1] = (c + 1 == K) ? sp->ne : (int64_t)((double)sp->ne * acc / tot); }
   }
   std::vector<int32_t> fc(sp->ndof, K), lc(sp->ndof, -1);
   for (int c = 0; c < K; c++)
      for (int64_t i = hp.eb[c] * sp->nd; i < hp.eb[c + 1] * sp->nd; i++)
      {
         const int32_t g = sp->gather[i];
         if (fc[g] > c) { fc[g] = c; }
         lc[g] = c;
      }
   hp.ub.assign((size_t)(K + 1) * 4, 0);
   hp.db.assign((size_t)(K + 1) * 4, 0);
   for (int k = 0; k < 4; k++)
   {
      const int64_t lo = sp->class_off[k], hi = sp->class_off[k + 1];
      std::vector<int64_t> umax(K, lo), dmin(K, hi);       // per chunk: max first-touched + 1 ; min dof with lc == c
      for (int64_t g = lo; g < hi; g++)
      {
         if (fc[g] < K) { umax[fc[g]] = std::max(umax[fc[g]], g + 1); }
         if (lc[g] >= 0) { dmin[lc[g]] = std::min(dmin[lc[g]], g); }
      }
      // needed[c+1]: everything chunks 0..c read ; final[c+1]: everything below min{g : lc[g] > c} is final after chunk c
      std::vector<int64_t> needed(K + 1, lo), fin(K + 1, lo), suffix(K + 1, hi);
      for (int c = 0; c < K; c++) { needed[c + 1] = std::max(needed[c], umax[c]); }
      needed[K] = hi;
      for (int c = K - 1; c >= 0; c--) { suffix[c] = std::min(suffix[c + 1], dmin[c]); }
      for (int c = 0; c < K; c++) { fin[c + 1] = std::max(fin[c], (c + 1 < K) ? suffix[c + 1] : hi); }
      // copies of class k happen every m-th chunk: uploads run ahead, downloads lag behind
      const int m = !shaped ? 1 : (k == 0 ? K : (k == 1 ? 2 : 1));
      hp.ub[k] = lo; hp.db[k] = lo;
      for (int c = 0; c < K; c++)
      {
         const bool up_now = (c % m) == 0, down_now = ((c + 1) % m) == 0 || c == K - 1;
         hp.ub[(size_t)(c + 1) * 4 + k] = up_now ? needed[std::min(K, (c / m + 1) * m)] : hp.ub[(size_t)c * 4 + k];
         hp.db[(size_t)(c + 1) * 4 + k] = down_now ? fin[c + 1] : hp.db[(size_t)c * 4 + k];
      }
   }
   CDM_CUDA(ctx, cudaStreamCreateWithFlags(&hp.su, cudaStreamNonBlocking));
   CDM_CUDA(ctx, cudaStreamCreateWithFlags(&hp.sd, cudaStreamNonBlocking));
   hp.ev_up.assign(K, nullptr); hp.ev_k.assign(K, nullptr);
   for (int c = 0; c < K; c++)
   {
      CDM_CUDA(ctx, cudaEventCreateWithFlags(&hp.ev_up[c], cudaEventDisableTiming));
      CDM_CUDA(ctx, cudaEventCreateWithFlags(&hp.ev_k[c], cudaEventDisableTiming));
   }
   // essential dofs grouped by the chunk whose download carries them: y[ess] = x[ess] is applied on the
   // device right before that download (a host loop over 235 k scattered entries costs 0.75 ms)
   {
      std::vector<std::vector<int32_t>> by_chunk(K);
      for (int32_t g : op->ess_host)
      {
         int k = 0;
         while (k < 3 && g >= sp->class_off[k + 1]) { k++; }
         int c = 0;
         while (c < K - 1 && g >= hp.db[(size_t)(c + 1) * 4 + k]) { c++; }
         by_chunk[c].push_back(g);
      }
      std::vector<int32_t> flat;
      hp.ess_off.assign(K + 1, 0);
      for (int c = 0; c < K; c++) { flat.insert(flat.end(), by_chunk[c].begin(), by_chunk[c].end()); hp.ess_off[c + 1] = (int64_t)flat.size(); }
      int rc = upload(ctx, flat, &hp.ess_by_chunk_dev);
      if (rc) { return rc; }
      CDM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
   }
   hp.K = K;
   return CDM_OK;
}

// x_host -> device in K pieces, element chunks as their inputs land, finished parts of y back to the host
// while later chunks still compute: the step costs about max(H2D, D2H) instead of H2D + apply + D2H.
static int mult_host_pipelined(cdm_op *op, const double *x_host, double *y_host, bool constrained)
{
   cdm_space *sp = op->sp;
   cdm_ctx *ctx = sp->ctx;
   int rc;
   if (op->pipe.K == 0 && (rc = build_host_pipe(op))) { destroy_host_pipe(op); return rc; }
   cdm_op::host_pipe &hp = op->pipe;
   const int K = hp.K;
   cudaStream_t C = ctx->stream;
   static const bool debug = getenv("CDM_PIPE_DEBUG") != nullptr;
   // CDM_PIPE_DEBUG: device timeline of the three streams (timing events, created on first use and kept)
   static std::vector<cudaEvent_t> dbg;                  // [0] start, then per chunk: upload done, kernel done, download done
   if (debug && dbg.size() != (size_t)(3 * K + 1))
   {
      for (cudaEvent_t e : dbg) { cudaEventDestroy(e); }
      dbg.assign((size_t)(3 * K + 1), nullptr);
      for (cudaEvent_t &e : dbg) { cudaEventCreate(&e); }
   }
   const auto t0 = std::chrono::steady_clock::now();
   // work already queued on the compute stream may still read xL / yL (cdm_eliminate_rhs returns without a sync):
   // the upload and download streams start after it
   CDM_CUDA(ctx, cudaEventRecord(ctx->ev0, C));
   if (debug) { cudaEventRecord(dbg[0], C); }
   CDM_CUDA(ctx, cudaStreamWaitEvent(hp.su, ctx->ev0, 0));
   CDM_CUDA(ctx, cudaStreamWaitEvent(hp.sd, ctx->ev0, 0));
   CDM_CUDA(ctx, cudaMemsetAsync(op->yL_dev, 0, sizeof(double) * (size_t)sp->ndof, C));
   // all uploads first (they depend on nothing else), then per chunk: kernel, download
   for (int c = 0; c < K; c++)
   {
      for (int k = 0; k < 4; k++)
      {
         const int64_t a = hp.ub[(size_t)c * 4 + k], b = hp.ub[(size_t)(c + 1) * 4 + k];
         if (b > a) { cudaMemcpyAsync(op->xL_dev + a, x_host + a, sizeof(double) * (size_t)(b - a), cudaMemcpyHostToDevice, hp.su); }
      }
      cudaEventRecord(hp.ev_up[c], hp.su);
      if (debug) { cudaEventRecord(dbg[1 + 3 * c], hp.su); }
   }
   const auto t1 = std::chrono::steady_clock::now();
   op->range_on = true;
   rc = CDM_OK;
   for (int c = 0; c < K && !rc; c++)
   {
      cudaStreamWaitEvent(C, hp.ev_up[c], 0);
      op->e_begin = hp.eb[c]; op->e_end = hp.eb[c + 1];
      rc = cdm_k_apply(op, op->xL_dev, op->yL_dev, constrained);
      if (!rc && constrained && hp.ess_off[c + 1] > hp.ess_off[c])
      {
         op->range_on = false;
         rc = cdm_k_copy_idx(ctx, hp.ess_off[c + 1] - hp.ess_off[c], hp.ess_by_chunk_dev + hp.ess_off[c], op->xL_dev, op->yL_dev);
         op->range_on = true;
      }
      cudaEventRecord(hp.ev_k[c], C);
      if (debug) { cudaEventRecord(dbg[2 + 3 * c], C); }
      cudaStreamWaitEvent(hp.sd, hp.ev_k[c], 0);
      for (int k = 0; k < 4; k++)
      {
         const int64_t a = hp.db[(size_t)c * 4 + k], b = hp.db[(size_t)(c + 1) * 4 + k];
         if (b > a) { cudaMemcpyAsync(y_host + a, op->yL_dev + a, sizeof(double) * (size_t)(b - a), cudaMemcpyDeviceToHost, hp.sd); }
      }
      if (debug) { cudaEventRecord(dbg[3 + 3 * c], hp.sd); }
   }
   op->range_on = false;
   const auto t2 = std::chrono::steady_clock::now();
   cudaError_t e1 = cudaStreamSynchronize(hp.sd), e2 = cudaStreamSynchronize(C), e3 = cudaStreamSynchronize(hp.su);
   const auto t3 = std::chrono::steady_clock::now();
   if (rc) { return rc; }
   if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) { return cdm_fail(ctx, CDM_ECUDA, "pipelined host apply failed"); }
   if (debug)
   {
      const auto t4 = std::chrono::steady_clock::now();
      auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b)
      { return std::chrono::duration<double, std::milli>(b - a).count(); };
      fprintf(stderr, "[cdm pipe] enqueue uploads %.3f ms, enqueue kernels+downloads %.3f ms, wait %.3f ms, ess fix %.3f ms\n",
              ms(t0, t1), ms(t1, t2), ms(t2, t3), ms(t3, t4));
      fprintf(stderr, "[cdm pipe] timeline (ms after start; chunk: upload done / kernel done / download done):");
      for (int c = 0; c < K; c++)
      {
         float a = 0.f, b = 0.f, d = 0.f;
         cudaEventElapsedTime(&a, dbg[0], dbg[1 + 3 * c]); cudaEventElapsedTime(&b, dbg[0], dbg[2 + 3 * c]); cudaEventElapsedTime(&d, dbg[0], dbg[3 + 3 * c]);
         fprintf(stderr, "  %d: %.3f / %.3f / %.3f", c, a, b, d);
      }
      fprintf(stderr, "\n");
   }
   return CDM_OK;
}

int cdm_operator_mult_host(cdm_op *op, const double *x_host, double *y_host, int constrained)
{
   if (!op || !x_host || !y_host) { return CDM_EINVAL; }
   cdm_space *sp = op->sp;
   cdm_ctx *ctx = sp->ctx;
   CDM_REQUIRE_GPU(ctx);
   int rc = ensure_L(op); if (rc) { return rc; }
   if (op->host_pipeline && ctx->nranks == 1 && sp->dim == 3 && sp->class_off.size() == 5 && op->scatter_mode == 1 &&
       cdm_k_range_capable(op) && sp->ne >= 8192)
   {
      return mult_host_pipelined(op, x_host, y_host, constrained != 0);
   }
   CDM_CUDA(ctx, cudaMemcpyAsync(op->xL_dev, x_host, sizeof(double) * (size_t)sp->ntrue, cudaMemcpyHostToDevice, ctx->stream));
   if ((rc = cdm_apply_tail(op, op->xL_dev, op->yL_dev, constrained != 0, false))) { return rc; }
   CDM_CUDA(ctx, cudaMemcpyAsync(y_host, op->yL_dev, sizeof(double) * (size_t)sp->ntrue, cudaMemcpyDeviceToHost, ctx->stream));
   CDM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
   return CDM_OK;
}

int cdm_operator_diag(cdm_op *op, double *d_dev)
{
   if (!op || !d_dev) { return CDM_EINVAL; }
   cdm_space *sp = op->sp;
   cdm_ctx *ctx = sp->ctx;
   CDM_REQUIRE_GPU(ctx);
   int rc;
   const bool par = ctx->nranks > 1 && !sp->peers.empty();
   double *dst = d_dev;
   if (sp->ntrue != sp->ndof) { if ((rc = ensure_L(op))) { return rc; } dst = op->yL_dev; }
   if (sp->geom == 1) { if ((rc = cdm_k_csr_diag(op, dst))) { return rc; } }
   else if ((rc = cdm_k_diag(op, dst))) { return rc; }
   if (par && (rc = cdm_halo_PT(op, dst))) { return rc; }
   if (dst != d_dev) { CDM_CUDA(ctx, cudaMemcpyAsync(d_dev, dst, sizeof(double) * (size_t)sp->ntrue, cudaMemcpyDeviceToDevice, ctx->stream)); }
   if (op->n_ess > 0) { rc = cdm_k_set_idx(ctx, op->n_ess, op->ess_dev, 1.0, d_dev); }
   return rc;
}

int cdm_eliminate_rhs(cdm_op *op, const double *x_dev, double *b_dev)
{
   if (!op || !x_dev || !b_dev) { return CDM_EINVAL; }
   cdm_space *sp = op->sp;
   cdm_ctx *ctx = sp->ctx;
   CDM_REQUIRE_GPU(ctx);
   int rc = ensure_L(op); if (rc) { return rc; }
   // w = 0 ; w[ess] = x[ess]
   CDM_CUDA(ctx, cudaMemsetAsync(op->xL_dev, 0, sizeof(double) * (size_t)sp->ndof, ctx->stream));
   if (op->n_ess > 0 && (rc = cdm_k_copy_idx(ctx, op->n_ess, op->ess_dev, x_dev, op->xL_dev))) { return rc; }
   if ((rc = cdm_apply_tail(op, op->xL_dev, op->yL_dev, false, false))) { return rc; }
   if ((rc = cdm_k_axpy(ctx, sp->ntrue, -1.0, op->yL_dev, b_dev))) { return rc; }
   if (op->n_ess > 0) { rc = cdm_k_copy_idx(ctx, op->n_ess, op->ess_dev, x_dev, b_dev); }
   return rc;
}

int cdm_operator_time_kernel(cdm_op *op, const double *x_dev, double *y_dev, int reps, int constrained, double *mean_ms)
{
   if (!op || !x_dev || !y_dev || reps < 1 || !mean_ms) { return CDM_EINVAL; }
   cdm_ctx *ctx = op->sp->ctx;
   CDM_REQUIRE_GPU(ctx);
   if (!ctx->evk0) { CDM_CUDA(ctx, cudaEventCreate(&ctx->evk0)); CDM_CUDA(ctx, cudaEventCreate(&ctx->evk1)); }
   double tot = 0.0;
   ctx->time_main = true;
   for (int i = 0; i < reps; i++)
   {
      int rc = cdm_k_apply(op, x_dev, y_dev, constrained != 0);
      if (rc) { ctx->time_main = false; return rc; }
      cudaError_t e = cudaEventSynchronize(ctx->evk1);
      if (e != cudaSuccess) { ctx->time_main = false; return cdm_fail(ctx, CDM_ECUDA, cudaGetErrorString(e)); }
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ctx->evk0, ctx->evk1);
      tot += ms;
   }
   ctx->time_main = false;
   CDM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
   *mean_ms = tot / reps;
   return CDM_OK;
}

int cdm_operator_get_qdata(const cdm_op *op, double *Ddiff, double *Dconv, double *Dmass)
{
   if (!op) { return CDM_EINVAL; }
   return cdm_k_get_qdata(op, Ddiff, Dconv, Dmass);
}

// ---------------------------------------------------------------- vectors

int cdm_vec_alloc(cdm_ctx *c, int64_t n, double **v)
{
   CDM_REQUIRE_GPU(c);
   if (n < 0 || !v) { return CDM_EINVAL; }
   if (cudaMalloc((void **)v, sizeof(double) * (size_t)(n > 0 ? n : 1)) != cudaSuccess)
   { cudaGetLastError(); return cdm_fail(c, CDM_ENOMEM, "cdm_vec_alloc: out of device memory"); }
   return CDM_OK;
}
int cdm_vec_free(cdm_ctx *c, double *v) { CDM_REQUIRE_GPU(c); CDM_CUDA(c, cudaFree(v)); return CDM_OK; }
int cdm_vec_set(cdm_ctx *c, int64_t n, double value, double *v) { CDM_REQUIRE_GPU(c); return cdm_k_set(c, n, value, v); }
int cdm_vec_upload(cdm_ctx *c, int64_t n, const double *host, double *v)
{
   CDM_REQUIRE_GPU(c);
   CDM_CUDA(c, cudaMemcpyAsync(v, host, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
   CDM_CUDA(c, cudaStreamSynchronize(c->stream));
   return CDM_OK;
}
int cdm_vec_download(cdm_ctx *c, int64_t n, const double *v, double *host)
{
   CDM_REQUIRE_GPU(c);
   CDM_CUDA(c, cudaMemcpyAsync(host, v, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
   CDM_CUDA(c, cudaStreamSynchronize(c->stream));
   return CDM_OK;
}
int cdm_axpy(cdm_ctx *c, int64_t n, double a, const double *x, double *y) { CDM_REQUIRE_GPU(c); return cdm_k_axpy(c, n, a, x, y); }
int cdm_add(cdm_ctx *c, int64_t n, const double *x, double a, const double *y, double *z) { CDM_REQUIRE_GPU(c); return cdm_k_add(c, n, x, a, y, z); }
int cdm_pointwise_mult(cdm_ctx *c, int64_t n, const double *d, const double *x, double *y) { CDM_REQUIRE_GPU(c); return cdm_k_pmult(c, n, d, x, y); }

static double *red_out(cdm_ctx *c) { return c->red_dev + (size_t)CDM_RED_MAXK * CDM_RED_BLOCKS; }

int cdm_mdot(cdm_ctx *c, int64_t n, int k, const double *w, const double *V, int64_t ldv, double *result_host)
{
   CDM_REQUIRE_GPU(c);
   if (!result_host) { return CDM_EINVAL; }
   int rc = cdm_k_mdot_dev(c, n, k, w, V, ldv, red_out(c)); if (rc) { return rc; }
   if ((rc = cdm_allreduce_sum(c, red_out(c), k))) { return rc; }
   CDM_CUDA(c, cudaMemcpyAsync(c->red_host, red_out(c), sizeof(double) * k, cudaMemcpyDeviceToHost, c->stream));
   CDM_CUDA(c, cudaStreamSynchronize(c->stream));
   for (int i = 0; i < k; i++) { result_host[i] = c->red_host[i]; }
   return CDM_OK;
}
int cdm_dot(cdm_ctx *c, int64_t n, const double *x, const double *y, double *result_host) { return cdm_mdot(c, n, 1, x, y, n, result_host); }
int cdm_norm2(cdm_ctx *c, int64_t n, const double *x, double *result_host)
{
   int rc = cdm_mdot(c, n, 1, x, x, n, result_host);
   if (!rc) { *result_host = std::sqrt(*result_host); }
   return rc;
}
int cdm_maxpy(cdm_ctx *c, int64_t n, int k, const double *h_host, const double *V, int64_t ldv, double *w)
{
   CDM_REQUIRE_GPU(c);
   if (k < 0 || k > CDM_RED_MAXK || (k > 0 && !h_host)) { return CDM_EINVAL; }
   double *hd = red_out(c) + CDM_RED_MAXK;
   for (int i = 0; i < k; i++) { c->red_host[CDM_RED_MAXK + i] = h_host[i]; }
   CDM_CUDA(c, cudaMemcpyAsync(hd, c->red_host + CDM_RED_MAXK, sizeof(double) * k, cudaMemcpyHostToDevice, c->stream));
   int rc = cdm_k_maxpy_dev(c, n, k, hd, V, ldv, w, nullptr);
   if (!rc) { CDM_CUDA(c, cudaStreamSynchronize(c->stream)); }
   return rc;
}

}  // extern "C"
