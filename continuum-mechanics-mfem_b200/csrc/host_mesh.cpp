// host_mesh.cpp -- mesh entry points of the C ABI (host side, setup only).
//
// Stands behind Mesh(file) + UniformRefinement + ParMesh(MPI_COMM_WORLD, *mesh)
// (linear_convection_diffusion_2D.cpp:290-305) for the Cartesian configurations
// of BASELINE.json; numbering follows MFEM Mesh::MakeCartesian2D/3D with
// sfc_ordering=false.
#include "cdm_internal.hpp"
#include <algorithm>
#include <memory>
#include <cmath>
#include <cstring>
#include <new>
#include <unordered_map>

int cdm_fail(const cdm_ctx *ctx, int code, const std::string &msg)
{
   if (ctx) { ctx->err = msg; }
   return code;
}

namespace
{
// smooth displacement vanishing on the boundary of the box; a in units of h
void perturb_point(int dim, const double *s, const double *h, double a, double *X)
{
   const double tp = 2.0 * M_PI;
   const double u = X[0] / s[0], v = X[1] / s[1], w = (dim == 3) ? X[2] / s[2] : 0.0;
   const double bump = std::sin(M_PI * u) * std::sin(M_PI * v) * ((dim == 3) ? std::sin(M_PI * w) : 1.0);
   const double d0 = std::sin(tp * (u + v) + 0.3);
   const double d1 = std::cos(tp * (v + w) + 0.5);
   const double d2 = std::sin(tp * (w + u) + 1.0);
   X[0] += a * h[0] * bump * d0;
   X[1] += a * h[1] * bump * d1;
   if (dim == 3) { X[2] += a * h[2] * bump * d2; }
}

// vertex coordinates of global lattice point (i,j,k)
void lattice_point(int dim, const int64_t *n, const double *s, double a, int64_t i, int64_t j,
                   int64_t k, double *X)
{
   const double h[3] = {s[0] / n[0], s[1] / n[1], dim == 3 ? s[2] / n[2] : 0.0};
   X[0] = (i == n[0]) ? s[0] : i * h[0];
   X[1] = (j == n[1]) ? s[1] : j * h[1];
   if (dim == 3) { X[2] = (k == n[2]) ? s[2] : k * h[2]; }
   const bool interior = i > 0 && i < n[0] && j > 0 && j < n[1] && (dim == 2 || (k > 0 && k < n[2]));
   if (a != 0.0 && interior) { perturb_point(dim, s, h, a, X); }
}

// Fill a (sub)box [lo,hi) of elements of the global lattice as a mesh.
void fill_box(cdm_mesh &m, int dim, const int64_t *gn, const double *s, double a,
              const int64_t *lo, const int64_t *hi)
{
   const int64_t nx = hi[0] - lo[0], ny = hi[1] - lo[1], nz = (dim == 3) ? hi[2] - lo[2] : 0;
   m.dim = dim;
   m.nv = (nx + 1) * (ny + 1) * ((dim == 3) ? nz + 1 : 1);
   m.ne = nx * ny * ((dim == 3) ? nz : 1);
   m.vx.resize((size_t)m.nv * dim);
   m.vglobal.resize(m.nv);
   const int nvpe = (dim == 2) ? 4 : 8;
   m.ev.resize((size_t)m.ne * nvpe);
   auto V = [&](int64_t i, int64_t j, int64_t k) { return (int32_t)(i + (nx + 1) * (j + (ny + 1) * k)); };
   for (int64_t k = 0; k <= nz; k++)
      for (int64_t j = 0; j <= ny; j++)
         for (int64_t i = 0; i <= nx; i++)
         {
            const int64_t gi = lo[0] + i, gj = lo[1] + j, gk = (dim == 3) ? lo[2] + k : 0;
            const int32_t v = V(i, j, k);
            lattice_point(dim, gn, s, a, gi, gj, gk, &m.vx[(size_t)v * dim]);
            m.vglobal[v] = gi + (gn[0] + 1) * (gj + (gn[1] + 1) * gk);
         }
   int64_t e = 0;
   if (dim == 2)
   {
      for (int64_t j = 0; j < ny; j++)
         for (int64_t i = 0; i < nx; i++, e++)
         {
            int32_t *v = &m.ev[(size_t)e * 4];
            v[0] = V(i, j, 0); v[1] = V(i + 1, j, 0); v[2] = V(i + 1, j + 1, 0); v[3] = V(i, j + 1, 0);
         }
   }
   else
   {
      for (int64_t k = 0; k < nz; k++)
         for (int64_t j = 0; j < ny; j++)
            for (int64_t i = 0; i < nx; i++, e++)
            {
               int32_t *v = &m.ev[(size_t)e * 8];
               v[0] = V(i, j, k); v[1] = V(i + 1, j, k); v[2] = V(i + 1, j + 1, k); v[3] = V(i, j + 1, k);
               v[4] = V(i, j, k + 1); v[5] = V(i + 1, j, k + 1);
               v[6] = V(i + 1, j + 1, k + 1); v[7] = V(i, j + 1, k + 1);
            }
   }
   // boundary elements: only faces on the physical boundary of the global box
   m.bv.clear(); m.battr.clear();
   auto quad = [&](int32_t a0, int32_t a1, int32_t a2, int32_t a3, int attr)
   { m.bv.push_back(a0); m.bv.push_back(a1); m.bv.push_back(a2); m.bv.push_back(a3); m.battr.push_back(attr); };
   auto seg = [&](int32_t a0, int32_t a1, int attr)
   { m.bv.push_back(a0); m.bv.push_back(a1); m.battr.push_back(attr); };
   if (dim == 2)
   {
      // bottom 1, right 2, top 3, left 4
      if (lo[1] == 0)     for (int64_t i = 0; i < nx; i++) { seg(V(i, 0, 0), V(i + 1, 0, 0), 1); }
      if (hi[0] == gn[0]) for (int64_t j = 0; j < ny; j++) { seg(V(nx, j, 0), V(nx, j + 1, 0), 2); }
      if (hi[1] == gn[1]) for (int64_t i = 0; i < nx; i++) { seg(V(i + 1, ny, 0), V(i, ny, 0), 3); }
      if (lo[0] == 0)     for (int64_t j = 0; j < ny; j++) { seg(V(0, j + 1, 0), V(0, j, 0), 4); }
   }
   else
   {
      // bottom(z=0)=1 front(y=0)=2 right(x=max)=3 back(y=max)=4 left(x=0)=5 top(z=max)=6
      if (lo[2] == 0)
         for (int64_t j = 0; j < ny; j++) for (int64_t i = 0; i < nx; i++)
            quad(V(i, j + 1, 0), V(i + 1, j + 1, 0), V(i + 1, j, 0), V(i, j, 0), 1);
      if (hi[2] == gn[2])
         for (int64_t j = 0; j < ny; j++) for (int64_t i = 0; i < nx; i++)
            quad(V(i, j, nz), V(i + 1, j, nz), V(i + 1, j + 1, nz), V(i, j + 1, nz), 6);
      if (lo[0] == 0)
         for (int64_t k = 0; k < nz; k++) for (int64_t j = 0; j < ny; j++)
            quad(V(0, j + 1, k), V(0, j, k), V(0, j, k + 1), V(0, j + 1, k + 1), 5);
      if (hi[0] == gn[0])
         for (int64_t k = 0; k < nz; k++) for (int64_t j = 0; j < ny; j++)
            quad(V(nx, j, k), V(nx, j + 1, k), V(nx, j + 1, k + 1), V(nx, j, k + 1), 3);
      if (lo[1] == 0)
         for (int64_t k = 0; k < nz; k++) for (int64_t i = 0; i < nx; i++)
            quad(V(i, 0, k), V(i + 1, 0, k), V(i + 1, 0, k + 1), V(i, 0, k + 1), 2);
      if (hi[1] == gn[1])
         for (int64_t k = 0; k < nz; k++) for (int64_t i = 0; i < nx; i++)
            quad(V(i + 1, ny, k), V(i, ny, k), V(i, ny, k + 1), V(i + 1, ny, k + 1), 4);
   }
   m.nbe = (int64_t)m.battr.size();
}
}  // namespace

// geometry of a Cartesian mesh is remembered so that partitioning can regenerate boxes
struct cart_info { double s[3]; double perturb; };
static std::unordered_map<const cdm_mesh *, cart_info> &cart_registry()
{
   static std::unordered_map<const cdm_mesh *, cart_info> r;
   return r;
}

extern "C" {

int cdm_mesh_cartesian(cdm_ctx *ctx, int dim, const int64_t n[3], const double size[3],
                       double perturb, cdm_mesh **mesh)
{
   if (!mesh || !n || (dim != 2 && dim != 3)) { return cdm_fail(ctx, CDM_EINVAL, "cdm_mesh_cartesian: bad arguments"); }
   for (int d = 0; d < dim; d++) if (n[d] < 1) { return cdm_fail(ctx, CDM_EINVAL, "cdm_mesh_cartesian: n < 1"); }
   double s[3] = {1.0, 1.0, 1.0};
   if (size) { for (int d = 0; d < dim; d++) { s[d] = size[d]; } }
   int64_t nn[3] = {n[0], n[1], dim == 3 ? n[2] : 0};
   {
      double tot = (double)(nn[0] + 1) * (nn[1] + 1) * (dim == 3 ? nn[2] + 1 : 1);
      if (tot > 2.0e9) { return cdm_fail(ctx, CDM_EINVAL, "cdm_mesh_cartesian: vertex count exceeds int32"); }
   }
   cdm_mesh *m = new (std::nothrow) cdm_mesh;
   if (!m) { return cdm_fail(ctx, CDM_ENOMEM, "out of memory"); }
   const int64_t lo[3] = {0, 0, 0};
   fill_box(*m, dim, nn, s, perturb, lo, nn);
   m->cartesian = true;
   for (int d = 0; d < 3; d++) { m->n[d] = nn[d]; m->gn[d] = nn[d]; }
   cart_info ci; ci.perturb = perturb; for (int d = 0; d < 3; d++) { ci.s[d] = s[d]; }
   cart_registry()[m] = ci;
   *mesh = m;
   return CDM_OK;
}

int cdm_mesh_from_arrays(cdm_ctx *ctx, int dim, int64_t nv, const double *vertices,
                         int64_t ne, const int32_t *elem_vtx,
                         int64_t nbe, const int32_t *bdr_vtx, const int32_t *bdr_attr,
                         cdm_mesh **mesh)
{
   if (!mesh || !vertices || !elem_vtx || (dim != 2 && dim != 3) || nv < 1 || ne < 1 || nbe < 0)
      return cdm_fail(ctx, CDM_EINVAL, "cdm_mesh_from_arrays: bad arguments");
   const int nvpe = (dim == 2) ? 4 : 8, nvpf = (dim == 2) ? 2 : 4;
   for (int64_t i = 0; i < ne * nvpe; i++)
      if (elem_vtx[i] < 0 || elem_vtx[i] >= nv) { return cdm_fail(ctx, CDM_EINVAL, "cdm_mesh_from_arrays: vertex id out of range"); }
   cdm_mesh *m = new (std::nothrow) cdm_mesh;
   if (!m) { return cdm_fail(ctx, CDM_ENOMEM, "out of memory"); }
   m->dim = dim; m->nv = nv; m->ne = ne; m->nbe = nbe;
   m->vx.assign(vertices, vertices + nv * dim);
   m->ev.assign(elem_vtx, elem_vtx + ne * nvpe);
   if (nbe > 0)
   {
      if (!bdr_vtx || !bdr_attr) { delete m; return cdm_fail(ctx, CDM_EINVAL, "cdm_mesh_from_arrays: missing boundary arrays"); }
      m->bv.assign(bdr_vtx, bdr_vtx + nbe * nvpf);
      m->battr.assign(bdr_attr, bdr_attr + nbe);
   }
   *mesh = m;
   return CDM_OK;
}

int cdm_mesh_sizes(const cdm_mesh *m, int *dim, int64_t *nv, int64_t *ne, int64_t *nbe)
{
   if (!m) { return CDM_EINVAL; }
   if (dim) { *dim = m->dim; }
   if (nv) { *nv = m->nv; }
   if (ne) { *ne = m->ne; }
   if (nbe) { *nbe = m->nbe; }
   return CDM_OK;
}

int cdm_mesh_get(const cdm_mesh *m, double *vertices, int32_t *elem_vtx, int32_t *bdr_vtx, int32_t *bdr_attr)
{
   if (!m) { return CDM_EINVAL; }
   if (vertices) { std::memcpy(vertices, m->vx.data(), m->vx.size() * sizeof(double)); }
   if (elem_vtx) { std::memcpy(elem_vtx, m->ev.data(), m->ev.size() * sizeof(int32_t)); }
   if (bdr_vtx && !m->bv.empty()) { std::memcpy(bdr_vtx, m->bv.data(), m->bv.size() * sizeof(int32_t)); }
   if (bdr_attr && !m->battr.empty()) { std::memcpy(bdr_attr, m->battr.data(), m->battr.size() * sizeof(int32_t)); }
   return CDM_OK;
}

int cdm_mesh_destroy(cdm_mesh *m)
{
   if (m) { cart_registry().erase(m); delete m; }
   return CDM_OK;
}

int cdm_mesh_partition_box(cdm_ctx *ctx, const cdm_mesh *g, const int parts[3], int rank, cdm_mesh **local)
{
   if (!g || !parts || !local) { return cdm_fail(ctx, CDM_EINVAL, "cdm_mesh_partition_box: bad arguments"); }
   auto it = cart_registry().find(g);
   if (!g->cartesian || it == cart_registry().end())
      return cdm_fail(ctx, CDM_EUNSUP, "cdm_mesh_partition_box: only meshes from cdm_mesh_cartesian can be box-partitioned");
   const int dim = g->dim;
   const int px = parts[0], py = parts[1], pz = (dim == 3) ? parts[2] : 1;
   if (px < 1 || py < 1 || pz < 1 || rank < 0 || rank >= px * py * pz)
      return cdm_fail(ctx, CDM_EINVAL, "cdm_mesh_partition_box: bad partition / rank");
   if (px > g->n[0] || py > g->n[1] || (dim == 3 && pz > g->n[2]))
      return cdm_fail(ctx, CDM_EINVAL, "cdm_mesh_partition_box: more parts than elements along an axis");
   const int pr[3] = {rank % px, (rank / px) % py, rank / (px * py)};
   const int pp[3] = {px, py, pz};
   int64_t lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
   for (int d = 0; d < dim; d++)
   {
      const int64_t q = g->n[d] / pp[d], r = g->n[d] % pp[d];
      lo[d] = pr[d] * q + std::min<int64_t>(pr[d], r);
      hi[d] = lo[d] + q + (pr[d] < r ? 1 : 0);
   }
   cdm_mesh *m = new (std::nothrow) cdm_mesh;
   if (!m) { return cdm_fail(ctx, CDM_ENOMEM, "out of memory"); }
   fill_box(*m, dim, g->n, it->second.s, it->second.perturb, lo, hi);
   m->is_part = true; m->rank = rank;
   for (int d = 0; d < 3; d++) { m->parts[d] = pp[d]; m->gn[d] = g->n[d]; m->n[d] = hi[d] - lo[d]; m->lo[d] = lo[d]; }
   *local = m;
   return CDM_OK;
}

// Element-wise partition of a quadrilateral / hexahedral mesh (what ParMesh(comm, mesh) does with the METIS array,
// linear_convection_diffusion_2D.cpp:300): this rank's elements in parent order, their vertices renumbered in ascending parent
// id (edge and face orientations are then the parent's), the parent boundary elements that are faces of local elements.
int cdm_mesh_partition_elements(cdm_ctx *ctx, const cdm_mesh *g, const int32_t *elem_rank, int nranks, int rank, cdm_mesh **local)
{
   if (!g || !elem_rank || !local) { return cdm_fail(ctx, CDM_EINVAL, "cdm_mesh_partition_elements: bad arguments"); }
   if (g->geom != 0) { return cdm_fail(ctx, CDM_EUNSUP, "cdm_mesh_partition_elements: quadrilateral / hexahedral meshes only"); }
   if (g->is_part) { return cdm_fail(ctx, CDM_EINVAL, "cdm_mesh_partition_elements: the mesh is already a part"); }
   if (nranks < 1 || nranks > 64 || rank < 0 || rank >= nranks)
      return cdm_fail(ctx, CDM_EINVAL, "cdm_mesh_partition_elements: 1 <= nranks <= 64 and 0 <= rank < nranks");
   const int dim = g->dim, nvpe = 1 << dim, nvpf = 1 << (dim - 1);
   for (int64_t e = 0; e < g->ne; e++)
      if (elem_rank[e] < 0 || elem_rank[e] >= nranks) { return cdm_fail(ctx, CDM_EINVAL, "cdm_mesh_partition_elements: rank id out of range"); }
   cdm_mesh *m = new (std::nothrow) cdm_mesh;
   if (!m) { return cdm_fail(ctx, CDM_ENOMEM, "out of memory"); }
   m->geom = 0; m->dim = dim;
   std::vector<int32_t> g2l((size_t)g->nv, -1);
   for (int64_t e = 0; e < g->ne; e++)
      if (elem_rank[e] == rank)
      {
         m->eglobal.push_back(e);
         for (int k = 0; k < nvpe; k++) { g2l[g->ev[(size_t)e * nvpe + k]] = 0; }
      }
   if (m->eglobal.empty()) { delete m; return cdm_fail(ctx, CDM_EINVAL, "cdm_mesh_partition_elements: this rank owns no element"); }
   for (int64_t v = 0; v < g->nv; v++)
      if (g2l[v] == 0)
      {
         g2l[v] = (int32_t)m->vglobal.size();
         m->vglobal.push_back(v);
         for (int c = 0; c < dim; c++) { m->vx.push_back(g->vx[(size_t)v * dim + c]); }
      }
   m->nv = (int64_t)m->vglobal.size(); m->ne = (int64_t)m->eglobal.size();
   m->ev.resize((size_t)m->ne * nvpe);
   for (int64_t e = 0; e < m->ne; e++)
      for (int k = 0; k < nvpe; k++) { m->ev[(size_t)e * nvpe + k] = g2l[g->ev[(size_t)m->eglobal[e] * nvpe + k]]; }
   // faces of the local elements, keyed by their sorted parent vertex ids
   static const int QE[4][2] = {{0,1},{1,2},{2,3},{3,0}};
   static const int HF[6][4] = {{3,2,1,0},{0,1,5,4},{1,2,6,5},{2,3,7,6},{3,0,4,7},{4,5,6,7}};
   struct FKey { int32_t v[4]; bool operator<(const FKey &o) const { return std::lexicographical_compare(v, v + 4, o.v, o.v + 4); }
                 bool operator==(const FKey &o) const { return std::equal(v, v + 4, o.v); } };
   auto make_key = [&](const int32_t *fv)
   {
      FKey k = {{-1, -1, -1, -1}};
      for (int i = 0; i < nvpf; i++)                              // insertion sort of 2 or 4 ids
      {
         int j = i;
         while (j > 0 && k.v[j - 1] > fv[i]) { k.v[j] = k.v[j - 1]; j--; }
         k.v[j] = fv[i];
      }
      return k;
   };
   std::vector<FKey> faces;
   faces.reserve((size_t)m->ne * (dim == 2 ? 4 : 6));
   for (int64_t e = 0; e < m->ne; e++)
   {
      const int32_t *v = &g->ev[(size_t)m->eglobal[e] * nvpe];
      for (int k = 0; k < (dim == 2 ? 4 : 6); k++)
      {
         int32_t fv[4];
         for (int i = 0; i < nvpf; i++) { fv[i] = (dim == 2) ? v[QE[k][i]] : v[HF[k][i]]; }
         faces.push_back(make_key(fv));
      }
   }
   std::sort(faces.begin(), faces.end());
   for (int64_t b = 0; b < g->nbe; b++)
   {
      const int32_t *bv = &g->bv[(size_t)b * nvpf];
      if (!std::binary_search(faces.begin(), faces.end(), make_key(bv))) { continue; }
      for (int i = 0; i < nvpf; i++) { m->bv.push_back(g2l[bv[i]]); }
      m->battr.push_back(g->battr[(size_t)b]);
   }
   m->nbe = (int64_t)m->battr.size();
   auto info = std::make_shared<cdm_part_info>();
   info->dim = dim; info->nranks = nranks; info->nv = g->nv; info->ne = g->ne;
   info->ev = g->ev; info->elem_rank.assign(elem_rank, elem_rank + g->ne);
   m->pinfo = info;
   m->is_part = true; m->rank = rank;
   *local = m;
   return CDM_OK;
}

}  // extern "C"
