// cdm_internal.hpp -- internal types shared by the host and CUDA translation units.
#pragma once
#include "cdm_b200.h"
#include <cuda_runtime.h>
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#define CDM_MAX_D1D 7
#define CDM_MAX_Q1D 8

struct ncclComm;

struct cdm_ctx
{
   int device = -1;                 // -1: host-only context
   cudaStream_t stream = nullptr;
   bool own_stream = false;
   mutable std::string err;
   int64_t launches = 0;
   // multi-GPU
   ncclComm *comm = nullptr;
   int rank = 0, nranks = 1;
   // reduction scratch (device) + pinned host mirror
   double *red_dev = nullptr;       // [RED_BLOCKS * RED_MAXK] partials + [RED_MAXK] results
   double *red_host = nullptr;      // pinned, RED_MAXK
   int sm_count = 148;
   cudaEvent_t ev0 = nullptr, ev1 = nullptr;
   cudaEvent_t ev_kry = nullptr;    // "scalars of this Krylov iteration are on the host"
   // halo / compute overlap: a high-priority stream with its own communicator (ncclCommSplit)
   cudaStream_t stream_halo = nullptr;
   ncclComm *comm_halo = nullptr;
   cudaEvent_t ev_h[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
   // kernel-only timing (cdm_operator_time_kernel)
   bool time_main = false;
   cudaEvent_t evk0 = nullptr, evk1 = nullptr;
   // per-context launch configuration of every kernel instantiation that needs the dynamic shared-memory
   // opt-in: (kernel, shared bytes) -> resident blocks per SM on THIS device (cdm_kernel_cfg)
   std::map<std::pair<const void *, size_t>, int> kernel_cfg;
   // device-side error word of the peer-memory halo exchange (mapped pinned host memory): a kernel that gave up
   // waiting for a neighbour sets it, the next host synchronisation point turns it into CDM_ENCCL
   unsigned int *p2p_err_host = nullptr, *p2p_err_dev = nullptr;
   // peer-memory all-reduce state (owned by the first partitioned space that set up the symmetric exchange)
   struct sym_state *red_sym = nullptr;
   int allreduce_mode = 1;         // 1: peer-memory all-reduce when available, 0: always ncclAllReduce
};

// what a submesh of an element-wise partition (cdm_mesh_partition_elements) remembers of its parent: enough to number the
// global H1 space and to find, for every local dof, the ranks that share it
struct cdm_part_info
{
   int dim = 0, nranks = 1;
   int64_t nv = 0, ne = 0;
   std::vector<int32_t> ev;          // parent element -> vertices
   std::vector<int32_t> elem_rank;   // parent element -> rank
};

struct cdm_mesh
{
   int geom = 0;                    // 0: tensor elements (quads / hexes), 1: simplices (triangles)
   int dim = 0;
   int64_t nv = 0, ne = 0, nbe = 0;
   std::vector<double> vx;          // nv*dim, vertex-major
   std::vector<int32_t> ev;         // ne*2^dim
   std::vector<int32_t> bv;         // nbe*2^(dim-1)
   std::vector<int32_t> battr;      // nbe
   // Cartesian provenance (for box partitioning)
   bool cartesian = false;
   int64_t n[3] = {0, 0, 0};
   // set on submeshes produced by cdm_mesh_partition_box
   bool is_part = false;
   int parts[3] = {1, 1, 1};
   int rank = 0;
   std::vector<int64_t> vglobal;    // local vertex -> global vertex id of the parent mesh
   std::vector<int64_t> eglobal;    // local element -> parent element (element-wise partitions)
   std::shared_ptr<const cdm_part_info> pinfo;   // set on submeshes produced by cdm_mesh_partition_elements
   int64_t gn[3] = {0, 0, 0};       // parent mesh element counts
   int64_t lo[3] = {0, 0, 0};       // first parent element of this box along each axis
};

// shared-dof exchange plan with one neighbour rank
struct cdm_halo_peer
{
   int rank = -1;
   // dofs I own that the peer holds as ghosts (peer sends me partial sums / I send values)
   std::vector<int32_t> own_idx;    // local dof ids (owned, < ntrue)
   // dofs the peer owns that I hold as ghosts
   std::vector<int32_t> ghost_idx;  // local dof ids (>= ntrue)
   int64_t own_off = 0, ghost_off = 0;   // offsets of this peer's lists in the fused plan
};

// symmetric shared-dof exchange ("halo sum"): every rank of a dof's sharing group sends its partial sum to every
// other rank of the group and all of them add the contributions in ascending rank order, so that after ONE exchange
// the whole L-vector (ghost entries included) is consistent and bitwise identical on all sharers
struct cdm_sym_peer
{
   int rank = -1;
   std::vector<int32_t> idx;        // local dofs shared with this rank, ordered by global key (same order on both sides)
   int64_t off = 0;                 // offset of this peer's segment in the concatenated list / receive buffer
};
struct cdm_sym_plan
{
   std::vector<cdm_sym_peer> peers;
   std::vector<int32_t> all;                          // per-peer lists concatenated in peer order
   std::vector<int32_t> sh_dof, sh_off, sh_src;       // distinct shared dofs; CSR of contributions in rank order, -1 = own value
   std::vector<int32_t> bdr_dofs;                     // (unused on the device) distinct shared dofs, same as sh_dof
   int32_t *all_dev = nullptr, *sh_dof_dev = nullptr, *sh_off_dev = nullptr, *sh_src_dev = nullptr;
   struct sym_state *st = nullptr;                    // peer-memory buffers / flags (halo_p2p.cu)
   int ready = 0;                                     // 0 not tried, 1 usable, -1 unavailable (fall back to P / P^T over NCCL)
};

// fused exchange plan over all peers (one pack + one unpack kernel per phase)
struct cdm_halo_plan
{
   std::vector<int32_t> own_all, ghost_all;          // per-peer lists concatenated in peer order
   std::vector<int32_t> pt_dof, pt_off, pt_src;      // P^T: distinct owned-shared dofs, CSR into the receive buffer
   int32_t *own_all_dev = nullptr, *ghost_all_dev = nullptr;
   int32_t *pt_dof_dev = nullptr, *pt_off_dev = nullptr, *pt_src_dev = nullptr;
   double *send_dev = nullptr, *recv_dev = nullptr;
   // peer-memory exchange (halo_p2p.cu, option "halo" = 1): the pack kernel stores straight into the
   // neighbours' receive buffers over NVLink and raises a flag there; the unpack kernel waits on its flags
   struct p2p_state *p2p = nullptr;
   bool p2p_active = false;         // set by the operator apply around its P / P^T pair
};

struct cdm_space
{
   cdm_ctx *ctx = nullptr;
   int dim = 0, p = 0, d1d = 0, q1d = 0, nd = 0, nq = 0;
   int64_t ne = 0, ndof = 0, ntrue = 0, nv = 0;
   // host
   std::vector<int32_t> gather, offsets, indices;     // ElementRestriction arrays
   std::vector<double> B, G, qw, nodes, qx;           // 1-D tables (q1d x d1d)
   std::vector<double> elem_x;                        // ne * nvpe * dim vertex coordinates
   // simplex (triangle) spaces: dense reference tables of the operator's rule -- sB[q*nd + i] = phi_i(x_q),
   // sG[(c*nq + q)*nd + i] = d phi_i / d xi_c (x_q), sqw[q], sqx[q*dim + c], snodes[i*dim + c]; nvpe = 3
   int geom = 0;
   std::vector<double> sB, sG, sqw, sqx, snodes;
   double *sB_dev = nullptr, *sG_dev = nullptr, *sqw_dev = nullptr, *sqx_dev = nullptr;
   // entity tables kept for essential-dof marking
   std::vector<int32_t> bdr_vtx, bdr_attr;
   std::vector<int32_t> bdr_dofs_flat, bdr_dofs_off;  // per boundary element: its dofs
   int64_t nbe = 0;
   // device
   int32_t *gather_dev = nullptr, *offsets_dev = nullptr, *indices_dev = nullptr;
   double *elem_x_dev = nullptr;
   int32_t *iota_dev = nullptr;                       // identity gather map [ne*nd] (E-vector input, lazy)
   double *work_dev = nullptr;                        // L-vector scratch of the form kernels (lazy, partitioned spaces)
   double *elem_part_dev = nullptr;                   // per-block partial sums of the error-norm kernel (lazy)
   // multi-GPU
   std::vector<cdm_halo_peer> peers;
   cdm_halo_plan halo;
   cdm_sym_plan sym;
   // partitioned spaces order their elements "boundary first": elements [0, n_bdr_elems) touch a
   // shared dof, the rest are interior; elem_perm[new] = element index in the mesh
   int64_t n_bdr_elems = 0;
   std::vector<int64_t> elem_perm;
   std::vector<int64_t> class_off;                     // entity-class dof ranges (spaces numbered by this library)
   std::vector<int64_t> dof_global;                    // local dof -> global dof id (partitioned spaces)
};

// assembled matrix of an operator (csr_path.cu)
struct cdm_csr
{
   int64_t n = 0, nnz = 0;
   std::vector<int64_t> rowptr;
   std::vector<int32_t> colind;
   int64_t *rowptr_dev = nullptr;
   int32_t *colind_dev = nullptr;
   double *vals_dev = nullptr;
   unsigned char *ess_mark_dev = nullptr;
};

struct cdm_op
{
   cdm_space *sp = nullptr;
   bool has_diff = false, has_conv = false, has_mass = false;
   int ncomp = 0;                  // stored D components per point
   int slab = 0;                   // doubles per (element, z-slab) [3D] or per element [2D]
   double *D_dev = nullptr;
   int64_t D_len = 0;
   int32_t *gather_c_dev = nullptr; // gather map with essential dofs encoded as -1-g
   int32_t *ess_dev = nullptr;      // owned essential dofs (y[ess] = x[ess]), followed by the ghost ones
   int64_t n_ess = 0;               // owned essential dofs
   int64_t n_ess_all = 0;           // owned + ghost essential dofs (ghost-consistent apply)
   std::vector<int32_t> ess_host;
   double *yE_dev = nullptr;       // E-vector scratch (scatter mode 0)
   double *xL_dev = nullptr, *yL_dev = nullptr;   // L-vector scratch (multi-GPU / host mult)
   double *dinv_dev = nullptr;     // cached Jacobi inverse diagonal
   bool range_on = false;          // launch only elements [e_begin, e_end) (overlapped multi-GPU apply)
   int64_t e_begin = 0, e_end = 0;
   int overlap = 1;                // 1: overlap the halo exchange with interior elements when possible
   bool tail = false;             // caller vectors have room for the ghost tail (length >= ndof)
   bool ghost_in = false;         // with tail: the caller guarantees that x's ghost tail is already consistent
   int scatter_mode = 1;           // 0: E-vector + gather transpose, 1: FP64 red.add
   int halo_mode = 2;              // 0: P / P^T over NCCL send/recv, 1: P / P^T over peer memory, 2: one symmetric peer-memory exchange
   int assembly = 0;               // 0: partial assembly (matrix-free), 1: apply = SpMV with the assembled CSR matrix
   struct cdm_csr *csr = nullptr;  // csr_path.cu (built on demand)
   struct cdm_ilu *ilu = nullptr;  // precond.cu: ILU(0) of the assembled matrix (built on demand)
   int ilu_sweep = 1;              // 1: single-launch triangular sweeps (dependency flags), 0: one launch per level
   int kernel_variant = 0;
   int64_t grid_cap = 0;           // > 0: upper bound on the persistent grids (tests: forces many elements per warp)
   double *e_out = nullptr;        // != null: E-vector output of the element kernels goes here and is not transposed
   const int32_t *gmap_override = nullptr;   // != null: gather map of the next launch (identity map: E-vector input)
   // pipelined host-vector apply (cdm_operator_mult_host): element chunks, per-class upload / download bounds
   struct host_pipe
   {
      int K = 0;
      std::vector<int64_t> eb, ub, db;      // eb[K+1]; ub, db[(K+1)*4]
      cudaStream_t su = nullptr, sd = nullptr;
      std::vector<cudaEvent_t> ev_up, ev_k;
      std::vector<int64_t> ess_off;         // essential dofs grouped by download chunk
      int32_t *ess_by_chunk_dev = nullptr;
   } pipe;
   int host_pipeline = 1;
   int host_pipeline_shape = 1;    // 1: tapered chunks + merged copies of the small entity classes, 0: round-1 schedule
   // krylov workspace (lazy)
   double *kry_dev = nullptr; int64_t kry_len = 0;
   std::vector<double> coef_scratch;
   double *coef_dev[3] = {nullptr, nullptr, nullptr};   // staging of host per-point coefficient arrays
   size_t coef_bytes[3] = {0, 0, 0};
};

// ---- error helpers
int cdm_fail(const cdm_ctx *ctx, int code, const std::string &msg);
// dynamic shared-memory opt-in + occupancy of one kernel instantiation, cached per context (= per device)
int cdm_kernel_cfg(cdm_ctx *ctx, const void *kern, int threads, size_t smem, const char *name, int *blocks_per_sm);
// CDM_ENCCL if a peer-memory halo kernel of this context timed out since the last check (clears the word)
int cdm_check_p2p(cdm_ctx *ctx);
#define CDM_CUDA(ctx, call)                                                              \
   do { cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess)                                                           \
           return cdm_fail(ctx, CDM_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
   } while (0)
// every compute entry point starts here: fail without a device, and make the context's device the calling
// thread's current one (several contexts on different devices may live in one process)
#define CDM_REQUIRE_GPU(ctx)                                                             \
   do { if (!(ctx) || (ctx)->device < 0)                                                 \
           return cdm_fail(ctx, CDM_ENOGPU, "no CUDA device bound to this context (no CPU fallback)"); \
        int cur_ = -1;                                                                   \
        if (cudaGetDevice(&cur_) != cudaSuccess || cur_ != (ctx)->device)                \
           CDM_CUDA(ctx, cudaSetDevice((ctx)->device));                                  \
   } while (0)

// ---- simplex elements (host_simplex.cpp, simplex.cu)
// nodes of the order-p Lagrange triangle (MFEM H1_TriangleElement node set), nd = (p+1)(p+2)/2, xy[i*2 + c]
void cdm_host_tri_nodes(int p, double *xy);
// collapsed Gauss-Legendre rule with n points per direction on the reference triangle (exact to degree 2n-2): n*n points
void cdm_host_tri_rule(int n, double *xy, double *w);
// values and reference gradients of the nodal basis at arbitrary reference points: B[q*nd + i], G[(c*npts + q)*nd + i]
void cdm_host_tri_basis(int p, int npts, const double *xy, double *B, double *G);
int64_t cdm_host_h1_numbering_tri(const cdm_mesh &m, int p, std::vector<int32_t> &elem_dof,
                                  std::vector<int32_t> &bdr_off, std::vector<int32_t> &bdr_flat);
int cdm_k_setup_qdata_simplex(cdm_op *op, const cdm_coeff *kappa, const cdm_coeff *vel, double alpha, const cdm_coeff *mass);
int cdm_simplex_rule_coords(const cdm_space *sp, int nq1d, double *xyz);           // host or device output
int cdm_simplex_domain_lf(cdm_space *sp, int nq1d, const double *f_q, double scale, int accumulate, double *b_dev);
int cdm_simplex_l2_error(cdm_space *sp, int nq1d, const double *u_dev, const double *uex_q, double *result_host);
// ---- host-side builders (host_*.cpp)
void cdm_host_gauss_legendre(int n, double *x, double *w);
void cdm_host_gauss_lobatto(int n, double *x);
void cdm_host_basis(int p, int q1d, double *B, double *G, double *qw, double *nodes, double *qx);
int  cdm_host_q1d(int dim, int p);
int64_t cdm_host_h1_numbering(const cdm_mesh &m, int p, std::vector<int32_t> &elem_dof,
                              std::vector<int32_t> &bdr_off, std::vector<int32_t> &bdr_flat,
                              int64_t *class_off = nullptr);   // [vertices | edges | faces | interiors] dof ranges
void cdm_host_restriction(int64_t ne, int nd, int64_t ndof, const std::vector<int32_t> &gather,
                          std::vector<int32_t> &offsets, std::vector<int32_t> &indices);

// ---- kernel launchers (kernels_*.cu); all asynchronous on ctx->stream
int cdm_k_setup_qdata(cdm_op *op, const cdm_coeff *kappa, const cdm_coeff *vel, double alpha,
                      const cdm_coeff *mass);
int cdm_k_apply(cdm_op *op, const double *xL, double *yL, bool constrained);
bool cdm_k_range_capable(const cdm_op *op);   // does the selected kernel honour op->range_on?
int cdm_k_diag(cdm_op *op, double *dL);
int cdm_k_diag_evec(cdm_op *op, double *dE);     // element-wise diagonal, E-vector layout
int cdm_k_ensure_yE(cdm_op *op);
int cdm_k_restrict_transpose(cdm_space *sp, const double *yE, double *yL);
int cdm_k_get_qdata(const cdm_op *op, double *Ddiff, double *Dconv, double *Dmass);
int cdm_k_upload_basis(cdm_space *sp);

int cdm_k_set(cdm_ctx *c, int64_t n, double v, double *x);
int cdm_k_axpy(cdm_ctx *c, int64_t n, double a, const double *x, double *y);
int cdm_k_add(cdm_ctx *c, int64_t n, const double *x, double a, const double *y, double *z);
int cdm_k_pmult(cdm_ctx *c, int64_t n, const double *d, const double *x, double *y);
int cdm_k_copy_idx(cdm_ctx *c, int64_t n, const int32_t *idx, const double *x, double *y);   // y[idx]=x[idx]
int cdm_k_zero_idx(cdm_ctx *c, int64_t n, const int32_t *idx, double *y);                    // y[idx]=0
int cdm_k_set_idx(cdm_ctx *c, int64_t n, const int32_t *idx, double v, double *y);           // y[idx]=v
int cdm_k_pack(cdm_ctx *c, int64_t n, const int32_t *idx, const double *x, double *buf);     // buf[i]=x[idx[i]]
int cdm_k_unpack(cdm_ctx *c, int64_t n, const int32_t *idx, const double *buf, double *x, int add);
int cdm_k_recip(cdm_ctx *c, int64_t n, const double *d, double *dinv);
// x[dof[k]] += sum_{j in [off[k],off[k+1])} buf[src[j]]  (fixed order)
int cdm_k_unpack_add_csr(cdm_ctx *c, int64_t n, const int32_t *dof, const int32_t *off, const int32_t *src,
                         const double *buf, double *x);
// k dots of w against V columns -> ctx->red_dev results [k] (device), deterministic
int cdm_k_mdot_dev(cdm_ctx *c, int64_t n, int k, const double *w, const double *V, int64_t ldv,
                   double *out_dev);
// w_out = dinv .* t fused with the k dots of w_out against V
int cdm_k_mdot_pc_dev(cdm_ctx *c, int64_t n, int k, const double *t, const double *dinv, double *w_out,
                      const double *V, int64_t ldv, double *out_dev);
// lazily normalised basis (V holds u_i, nrm2[i] = ||u_i||^2): w_out = dinv .* t / sqrt(*sc2), out[i] = (w_out, u_i) / ||u_i||
int cdm_k_mdot_lazy_dev(cdm_ctx *c, int64_t n, int k, const double *t, const double *dinv, const double *sc2,
                        double *w_out, const double *V, int64_t ldv, const double *nrm2, double *out_dev);
int cdm_k_pmult_scaled(cdm_ctx *c, int64_t n, const double *dinv, const double *sc2, const double *x, double *y);
int cdm_k_maxpy_lazy_dev(cdm_ctx *c, int64_t n, int k, const double *h_dev, const double *nrm2, const double *V, int64_t ldv,
                         double *w, double *norm2_out_dev);
// w -= sum_i h_dev[i] V_i ; optionally also out_dev[0] = ||w_new||^2 partial-reduced
int cdm_k_maxpy_dev(cdm_ctx *c, int64_t n, int k, const double *h_dev, const double *V, int64_t ldv,
                    double *w, double *norm2_out_dev);
// v = w * (1/sqrt(*norm2_dev))
int cdm_k_scale_by_rnorm(cdm_ctx *c, int64_t n, const double *norm2_dev, const double *w, double *v);
int cdm_k_scale(cdm_ctx *c, int64_t n, double a, const double *w, double *v);
// h[i] /= sqrt(nrm2[i]), i < k (dots against un-normalised basis vectors)
int cdm_k_scale_dots(cdm_ctx *c, int k, const double *nrm2, double *h);
// fused CG update: x += a d ; r -= a z ; out = (r,r)
int cdm_k_cg_update(cdm_ctx *c, int64_t n, double a, const double *d, const double *z, double *x,
                    double *r, double *rr_out_dev);

// ---- halo exchange (comm.cpp)
int cdm_halo_P(cdm_op *op, double *xL);            // owner -> ghost values
int cdm_halo_PT(cdm_op *op, double *yL);           // ghost partial sums -> owner (add)
int cdm_halo_P_space(cdm_space *sp, double *xL);   // the same exchanges for callers without an operator
int cdm_halo_PT_space(cdm_space *sp, double *yL);
int cdm_allreduce_sum(cdm_ctx *c, double *buf_dev, int k);
int cdm_allgather_bytes(cdm_ctx *c, const void *send_dev, void *recv_dev, size_t bytes_per_rank);
// ---- full-assembly path (csr_path.cu)
void cdm_csr_destroy(cdm_op *op);
int cdm_csr_refill_if_present(cdm_op *op);          // new coefficient values into an existing pattern
int cdm_k_csr_spmv(cdm_op *op, const double *x, double *y, bool constrained);
int cdm_k_csr_diag(cdm_op *op, double *d);           // diagonal of the assembled matrix
// ---- ILU(0) of the assembled matrix (precond.cu): -pc_type bjacobi -sub_pc_type ilu on one rank
int cdm_ilu_setup(cdm_op *op);                      // symbolic (levels, host) + numeric factorisation (device)
int cdm_ilu_refactor(cdm_op *op);                   // numeric factorisation only (new matrix values)
int cdm_ilu_apply(cdm_op *op, const double *r, double *z);   // z = U^{-1} L^{-1} r; essential rows pass through
void cdm_ilu_destroy(cdm_op *op);
int cdm_ilu_check(cdm_op *op);                      // blocking: CDM_ECUDA if a single-launch sweep gave up waiting for a row
// ---- peer-memory halo exchange (halo_p2p.cu)
int cdm_halo_p2p_setup(cdm_space *sp);              // collective over the ranks of the communicator
void cdm_halo_p2p_destroy(cdm_space *sp);
int cdm_halo_p2p_P(cdm_space *sp, double *xL, cudaStream_t s, cudaEvent_t ev_packed);
int cdm_halo_p2p_PT(cdm_space *sp, double *yL, cudaStream_t s, cudaEvent_t ev_packed);
// symmetric exchange + Krylov all-reduce over peer memory (halo_p2p.cu)
int cdm_halo_sym_setup(cdm_space *sp);              // collective; sets sp->sym.ready to 1 (usable) or -1 (fall back to NCCL)
void cdm_halo_sym_destroy(cdm_space *sp);
int cdm_halo_sym_exchange(cdm_space *sp, double *yL, cudaStream_t s);
bool cdm_allreduce_sym(cdm_ctx *c, double *buf_dev, int k, int *rc_out);
// same exchanges on the halo stream / communicator; ev_packed is recorded right after the pack kernel
int cdm_halo_P_async(cdm_op *op, double *xL, cudaEvent_t ev_packed);
int cdm_halo_PT_async(cdm_op *op, double *yL, cudaEvent_t ev_packed);
