// host_io.cpp -- the data formats either side of the path, host side only:
//   * Gmsh 2.2 ASCII reader                      Mesh(mesh_file, 1, 1)      linear_convection_diffusion_2D.cpp:290
//                                                (Mesh/unit_square.msh, unit_circle.msh, square_0p01.msh, ablation_strip*.msh)
//   * Hilbert-curve element order of inline      Mesh::MakeCartesian2D/3D(..., sfc_ordering = true), the default of MFEM's
//     Cartesian meshes                           inline-quad / inline-hex meshes (BASELINE configs 1, 2)
//   * YAML driver input + PETSc options file     LoadParams :62-127, MFEMInitializePetsc(..., petsc_file, ...) :268-282,
//                                                Input/input_2d.yaml, Input/petsc.opts, Input/petsc_circle.opts
// (the ParaView writer lives in host_vtu.cpp)
// [MFEM-upstream, restated from memory]: vertex / element order of the Gmsh reader (file order, unused vertices
// removed), MarkTriMeshForRefinement (longest edge first), the generalised Hilbert curve of NCMesh::GridSfcOrdering.
#include "cdm_internal.hpp"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <new>
#include <sstream>
#include <string>

// ------------------------------------------------------------------------------------------- Hilbert curve
namespace
{
inline int sgn(int64_t x) { return (x < 0) ? -1 : ((x > 0) ? 1 : 0); }

void hilbert2d(int64_t x, int64_t y, int64_t ax, int64_t ay, int64_t bx, int64_t by, std::vector<int64_t> &out)
{
   const int64_t w = std::llabs(ax + ay), h = std::llabs(bx + by);
   const int dax = sgn(ax), day = sgn(ay), dbx = sgn(bx), dby = sgn(by);
   if (h == 1) { for (int64_t i = 0; i < w; i++, x += dax, y += day) { out.push_back(x); out.push_back(y); } return; }
   if (w == 1) { for (int64_t i = 0; i < h; i++, x += dbx, y += dby) { out.push_back(x); out.push_back(y); } return; }
   int64_t ax2 = ax / 2, ay2 = ay / 2, bx2 = bx / 2, by2 = by / 2;
   const int64_t w2 = std::llabs(ax2 + ay2), h2 = std::llabs(bx2 + by2);
   if (2 * w > 3 * h)                                     // long case: split in two parts only
   {
      if ((w2 & 1) && w > 2) { ax2 += dax; ay2 += day; }   // prefer even steps
      hilbert2d(x, y, ax2, ay2, bx, by, out);
      hilbert2d(x + ax2, y + ay2, ax - ax2, ay - ay2, bx, by, out);
   }
   else                                                   // one step up, one long horizontal step, one step down
   {
      if ((h2 & 1) && h > 2) { bx2 += dbx; by2 += dby; }
      hilbert2d(x, y, bx2, by2, ax2, ay2, out);
      hilbert2d(x + bx2, y + by2, ax, ay, bx - bx2, by - by2, out);
      hilbert2d(x + (ax - dax) + (bx2 - dbx), y + (ay - day) + (by2 - dby), -bx2, -by2, -(ax - ax2), -(ay - ay2), out);
   }
}

void hilbert3d(int64_t x, int64_t y, int64_t z, int64_t ax, int64_t ay, int64_t az, int64_t bx, int64_t by, int64_t bz,
               int64_t cx, int64_t cy, int64_t cz, std::vector<int64_t> &out)
{
   const int64_t w = std::llabs(ax + ay + az), h = std::llabs(bx + by + bz), d = std::llabs(cx + cy + cz);
   const int dax = sgn(ax), day = sgn(ay), daz = sgn(az), dbx = sgn(bx), dby = sgn(by), dbz = sgn(bz),
             dcx = sgn(cx), dcy = sgn(cy), dcz = sgn(cz);
   auto push = [&]() { out.push_back(x); out.push_back(y); out.push_back(z); };
   if (h == 1 && d == 1) { for (int64_t i = 0; i < w; i++, x += dax, y += day, z += daz) { push(); } return; }
   if (w == 1 && d == 1) { for (int64_t i = 0; i < h; i++, x += dbx, y += dby, z += dbz) { push(); } return; }
   if (w == 1 && h == 1) { for (int64_t i = 0; i < d; i++, x += dcx, y += dcy, z += dcz) { push(); } return; }
   int64_t ax2 = ax / 2, ay2 = ay / 2, az2 = az / 2, bx2 = bx / 2, by2 = by / 2, bz2 = bz / 2, cx2 = cx / 2, cy2 = cy / 2, cz2 = cz / 2;
   const int64_t w2 = std::llabs(ax2 + ay2 + az2), h2 = std::llabs(bx2 + by2 + bz2), d2 = std::llabs(cx2 + cy2 + cz2);
   if ((w2 & 1) && w > 2) { ax2 += dax; ay2 += day; az2 += daz; }
   if ((h2 & 1) && h > 2) { bx2 += dbx; by2 += dby; bz2 += dbz; }
   if ((d2 & 1) && d > 2) { cx2 += dcx; cy2 += dcy; cz2 += dcz; }
   if (2 * w > 3 * h && 2 * w > 3 * d)                    // wide case: split in w only
   {
      hilbert3d(x, y, z, ax2, ay2, az2, bx, by, bz, cx, cy, cz, out);
      hilbert3d(x + ax2, y + ay2, z + az2, ax - ax2, ay - ay2, az - az2, bx, by, bz, cx, cy, cz, out);
   }
   else if (3 * h > 4 * d)                                // do not split in d
   {
      hilbert3d(x, y, z, bx2, by2, bz2, cx, cy, cz, ax2, ay2, az2, out);
      hilbert3d(x + bx2, y + by2, z + bz2, ax, ay, az, bx - bx2, by - by2, bz - bz2, cx, cy, cz, out);
      hilbert3d(x + (ax - dax) + (bx2 - dbx), y + (ay - day) + (by2 - dby), z + (az - daz) + (bz2 - dbz),
                -bx2, -by2, -bz2, cx, cy, cz, -(ax - ax2), -(ay - ay2), -(az - az2), out);
   }
   else if (3 * d > 4 * h)                                // do not split in h
   {
      hilbert3d(x, y, z, cx2, cy2, cz2, ax2, ay2, az2, bx, by, bz, out);
      hilbert3d(x + cx2, y + cy2, z + cz2, ax, ay, az, bx, by, bz, cx - cx2, cy - cy2, cz - cz2, out);
      hilbert3d(x + (ax - dax) + (cx2 - dcx), y + (ay - day) + (cy2 - dcy), z + (az - daz) + (cz2 - dcz),
                -cx2, -cy2, -cz2, -(ax - ax2), -(ay - ay2), -(az - az2), bx, by, bz, out);
   }
   else                                                   // regular case: split in all of w, h, d
   {
      hilbert3d(x, y, z, bx2, by2, bz2, cx2, cy2, cz2, ax2, ay2, az2, out);
      hilbert3d(x + bx2, y + by2, z + bz2, cx, cy, cz, ax2, ay2, az2, bx - bx2, by - by2, bz - bz2, out);
      hilbert3d(x + (bx2 - dbx) + (cx - dcx), y + (by2 - dby) + (cy - dcy), z + (bz2 - dbz) + (cz - dcz),
                ax, ay, az, -bx2, -by2, -bz2, -(cx - cx2), -(cy - cy2), -(cz - cz2), out);
      hilbert3d(x + (ax - dax) + bx2 + (cx - dcx), y + (ay - day) + by2 + (cy - dcy), z + (az - daz) + bz2 + (cz - dcz),
                -cx, -cy, -cz, -(ax - ax2), -(ay - ay2), -(az - az2), bx - bx2, by - by2, bz - bz2, out);
      hilbert3d(x + (ax - dax) + (bx2 - dbx), y + (ay - day) + (by2 - dby), z + (az - daz) + (bz2 - dbz),
                -bx2, -by2, -bz2, cx2, cy2, cz2, -(ax - ax2), -(ay - ay2), -(az - az2), out);
   }
}
}  // namespace

extern "C" {

// Cell coordinates along the generalised Hilbert curve of an nx x ny (x nz) grid: coords[k*dim + c]
// (NCMesh::GridSfcOrdering2D / 3D).  Every cell appears exactly once.
int cdm_grid_sfc_ordering(int dim, const int64_t n[3], int64_t *coords)
{
   if (!n || !coords || (dim != 2 && dim != 3)) { return CDM_EINVAL; }
   std::vector<int64_t> out;
   if (dim == 2)
   {
      out.reserve((size_t)2 * n[0] * n[1]);
      if (n[0] >= n[1]) { hilbert2d(0, 0, n[0], 0, 0, n[1], out); }
      else { hilbert2d(0, 0, 0, n[1], n[0], 0, out); }
   }
   else
   {
      out.reserve((size_t)3 * n[0] * n[1] * n[2]);
      const int64_t w = n[0], h = n[1], d = n[2];
      if (w >= h && w >= d) { hilbert3d(0, 0, 0, w, 0, 0, 0, h, 0, 0, 0, d, out); }
      else if (h >= w && h >= d) { hilbert3d(0, 0, 0, 0, h, 0, w, 0, 0, 0, 0, d, out); }
      else { hilbert3d(0, 0, 0, 0, 0, d, w, 0, 0, 0, h, 0, out); }
   }
   std::memcpy(coords, out.data(), out.size() * sizeof(int64_t));
   return CDM_OK;
}

// Mesh::MakeCartesian2D / 3D with sfc_ordering = true: the mesh of cdm_mesh_cartesian with its elements listed along
// the Hilbert curve (vertices and boundary elements are unchanged).  The result cannot be box-partitioned.
int cdm_mesh_cartesian_sfc(cdm_ctx *ctx, int dim, const int64_t n[3], const double size[3], double perturb, cdm_mesh **mesh)
{
   cdm_mesh *m = nullptr;
   int rc = cdm_mesh_cartesian(ctx, dim, n, size, perturb, &m);
   if (rc) { return rc; }
   std::vector<int64_t> co((size_t)m->ne * dim);
   const int64_t nn[3] = {n[0], n[1], dim == 3 ? n[2] : 1};
   if ((rc = cdm_grid_sfc_ordering(dim, nn, co.data()))) { cdm_mesh_destroy(m); return rc; }
   const int nvpe = dim == 2 ? 4 : 8;
   std::vector<int32_t> ev((size_t)m->ne * nvpe);
   for (int64_t k = 0; k < m->ne; k++)
   {
      const int64_t e = co[k * dim] + nn[0] * (co[k * dim + 1] + (dim == 3 ? nn[1] * co[k * dim + 2] : 0));
      std::memcpy(&ev[(size_t)k * nvpe], &m->ev[(size_t)e * nvpe], sizeof(int32_t) * nvpe);
   }
   m->ev.swap(ev);
   cdm_mesh *copy = new (std::nothrow) cdm_mesh(*m);      // a plain mesh: no Cartesian provenance (no box partition)
   cdm_mesh_destroy(m);
   if (!copy) { return cdm_fail(ctx, CDM_ENOMEM, "out of memory"); }
   copy->cartesian = false;
   *mesh = copy;
   return CDM_OK;
}

int cdm_mesh_from_arrays_simplex(cdm_ctx *ctx, int dim, int64_t nv, const double *vertices, int64_t ne, const int32_t *elem_vtx,
                                 int64_t nbe, const int32_t *bdr_vtx, const int32_t *bdr_attr, cdm_mesh **mesh)
{
   if (!mesh || !vertices || !elem_vtx || nv < 1 || ne < 1 || nbe < 0) { return cdm_fail(ctx, CDM_EINVAL, "cdm_mesh_from_arrays_simplex: bad arguments"); }
   if (dim != 2) { return cdm_fail(ctx, CDM_EUNSUP, "cdm_mesh_from_arrays_simplex: triangles only (dim = 2)"); }
   for (int64_t i = 0; i < ne * 3; i++)
      if (elem_vtx[i] < 0 || elem_vtx[i] >= nv) { return cdm_fail(ctx, CDM_EINVAL, "cdm_mesh_from_arrays_simplex: vertex id out of range"); }
   cdm_mesh *m = new (std::nothrow) cdm_mesh;
   if (!m) { return cdm_fail(ctx, CDM_ENOMEM, "out of memory"); }
   m->geom = 1; m->dim = 2; m->nv = nv; m->ne = ne; m->nbe = nbe;
   m->vx.assign(vertices, vertices + nv * 2);
   m->ev.assign(elem_vtx, elem_vtx + ne * 3);
   if (nbe > 0)
   {
      if (!bdr_vtx || !bdr_attr) { delete m; return cdm_fail(ctx, CDM_EINVAL, "cdm_mesh_from_arrays_simplex: missing boundary arrays"); }
      m->bv.assign(bdr_vtx, bdr_vtx + nbe * 2);
      m->battr.assign(bdr_attr, bdr_attr + nbe);
   }
   *mesh = m;
   return CDM_OK;
}

// ------------------------------------------------------------------------------------------- uniform refinement (hexes)
// MFEM's UniformRefinement3D_base for hexahedra, restated: edges and quadrilateral faces numbered in first-encounter order
// over the elements (local edges / faces of hex_t); new vertices = old vertices, edge midpoints, face centres, element
// centres; child k of an element sits at its corner k, vertices listed in hex order; a boundary quadrilateral becomes four
// (children at its corners).  Averages as in Mesh::AverageVertices, recomputed by every element that visits an entity (the
// last visitor's summation order stays).
static int refine_hexes(cdm_ctx *ctx, const cdm_mesh *m, cdm_mesh **refined)
{
   static const int HE[12][2] = {{0,1},{1,2},{3,2},{0,3},{4,5},{5,6},{7,6},{4,7},{0,4},{1,5},{2,6},{3,7}};
   static const int HF[6][4] = {{3,2,1,0},{0,1,5,4},{1,2,6,5},{2,3,7,6},{3,0,4,7},{4,5,6,7}};
   struct ENode { int32_t hi, id, next; };
   struct FNode { int32_t b, c, id, next; };
   std::vector<int32_t> ehead((size_t)m->nv, -1), fhead((size_t)m->nv, -1);
   std::vector<ENode> epool; std::vector<FNode> fpool;
   int32_t nedges = 0, nfaces = 0;
   auto edge = [&](int32_t a, int32_t b, bool insert) -> int32_t
   {
      const int32_t lo = std::min(a, b), hi = std::max(a, b);
      for (int32_t n = ehead[lo]; n >= 0; n = epool[n].next) { if (epool[n].hi == hi) { return epool[n].id; } }
      if (!insert) { return -1; }
      epool.push_back(ENode{hi, nedges, ehead[lo]});
      ehead[lo] = (int32_t)epool.size() - 1;
      return nedges++;
   };
   auto face = [&](const int32_t *fv, bool insert) -> int32_t
   {
      int32_t s4[4] = {fv[0], fv[1], fv[2], fv[3]};
      std::sort(s4, s4 + 4);
      for (int32_t n = fhead[s4[0]]; n >= 0; n = fpool[n].next) { if (fpool[n].b == s4[1] && fpool[n].c == s4[2]) { return fpool[n].id; } }
      if (!insert) { return -1; }
      fpool.push_back(FNode{s4[1], s4[2], nfaces, fhead[s4[0]]});
      fhead[s4[0]] = (int32_t)fpool.size() - 1;
      return nfaces++;
   };
   std::vector<int32_t> el_edge((size_t)m->ne * 12), el_face((size_t)m->ne * 6);
   for (int64_t e = 0; e < m->ne; e++)
   {
      const int32_t *v = &m->ev[(size_t)e * 8];
      for (int k = 0; k < 12; k++) { el_edge[(size_t)e * 12 + k] = edge(v[HE[k][0]], v[HE[k][1]], true); }
   }
   for (int64_t e = 0; e < m->ne; e++)
   {
      const int32_t *v = &m->ev[(size_t)e * 8];
      for (int k = 0; k < 6; k++)
      {
         const int32_t fv[4] = {v[HF[k][0]], v[HF[k][1]], v[HF[k][2]], v[HF[k][3]]};
         el_face[(size_t)e * 6 + k] = face(fv, true);
      }
   }
   const int64_t oedge = m->nv, oface = oedge + nedges, oelem = oface + nfaces, nv2 = oelem + m->ne;
   if (nv2 > 2147483000LL || m->ne * 8 > 2147483000LL) { return cdm_fail(ctx, CDM_EUNSUP, "cdm_mesh_uniform_refine: more than 2^31 vertices or elements"); }
   cdm_mesh *r = new (std::nothrow) cdm_mesh;
   if (!r) { return cdm_fail(ctx, CDM_ENOMEM, "out of memory"); }
   r->geom = 0; r->dim = 3; r->nv = nv2; r->ne = m->ne * 8; r->nbe = m->nbe * 4;
   r->vx.assign((size_t)nv2 * 3, 0.0);
   std::copy(m->vx.begin(), m->vx.end(), r->vx.begin());
   auto average = [&](const int32_t *idx, int n, int64_t dst)
   {
      for (int c = 0; c < 3; c++)
      {
         double s = 0.0;
         for (int k = 0; k < n; k++) { s += r->vx[(size_t)idx[k] * 3 + c]; }
         r->vx[(size_t)dst * 3 + c] = s * (1.0 / n);
      }
   };
   r->ev.resize((size_t)r->ne * 8);
   for (int64_t i = 0; i < m->ne; i++)
   {
      const int32_t *v = &m->ev[(size_t)i * 8];
      int32_t e[12], f[6];
      const int32_t c = (int32_t)(oelem + i);
      average(v, 8, c);
      for (int k = 0; k < 6; k++)
      {
         f[k] = (int32_t)(oface + el_face[(size_t)i * 6 + k]);
         const int32_t fv[4] = {v[HF[k][0]], v[HF[k][1]], v[HF[k][2]], v[HF[k][3]]};
         average(fv, 4, f[k]);
      }
      for (int k = 0; k < 12; k++)
      {
         e[k] = (int32_t)(oedge + el_edge[(size_t)i * 12 + k]);
         const int32_t pair[2] = {v[HE[k][0]], v[HE[k][1]]};
         average(pair, 2, e[k]);
      }
      const int32_t ch[64] = {
         v[0], e[0], f[0], e[3], e[8], f[1], c, f[4],
         e[0], v[1], e[1], f[0], f[1], e[9], f[2], c,
         f[0], e[1], v[2], e[2], c, f[2], e[10], f[3],
         e[3], f[0], e[2], v[3], f[4], c, f[3], e[11],
         e[8], f[1], c, f[4], v[4], e[4], f[5], e[7],
         f[1], e[9], f[2], c, e[4], v[5], e[5], f[5],
         c, f[2], e[10], f[3], f[5], e[5], v[6], e[6],
         f[4], c, f[3], e[11], e[7], f[5], e[6], v[7]};
      std::copy(ch, ch + 64, &r->ev[(size_t)i * 64]);
   }
   r->bv.resize((size_t)r->nbe * 4);
   r->battr.resize((size_t)r->nbe);
   for (int64_t b = 0; b < m->nbe; b++)
   {
      const int32_t *v = &m->bv[(size_t)b * 4];
      int32_t e[4];
      for (int k = 0; k < 4; k++)
      {
         const int32_t id = edge(v[k], v[(k + 1) % 4], false);
         if (id < 0) { delete r; return cdm_fail(ctx, CDM_EINVAL, "cdm_mesh_uniform_refine: boundary element is not a face of the mesh"); }
         e[k] = (int32_t)(oedge + id);
      }
      const int32_t fid = face(v, false);
      if (fid < 0) { delete r; return cdm_fail(ctx, CDM_EINVAL, "cdm_mesh_uniform_refine: boundary element is not a face of the mesh"); }
      const int32_t q = (int32_t)(oface + fid);
      const int32_t ch[16] = {v[0], e[0], q, e[3],  e[0], v[1], e[1], q,  q, e[1], v[2], e[2],  e[3], q, e[2], v[3]};
      std::copy(ch, ch + 16, &r->bv[(size_t)b * 16]);
      for (int k = 0; k < 4; k++) { r->battr[(size_t)b * 4 + k] = m->battr[(size_t)b]; }
   }
   *refined = r;
   return CDM_OK;
}

// ------------------------------------------------------------------------------------------- uniform refinement (2D)
// Mesh::UniformRefinement() of a conforming triangle or quadrilateral mesh (linear_convection_diffusion_2D.cpp:295-298,
// serial_ref_levels / par_ref_levels; Input/input_diffusion_mms.yaml refines Mesh/unit_square.msh once).  MFEM's
// UniformRefinement2D_base, restated: edges are numbered in first-encounter order over the elements (local edges (0,1),
// (1,2), (2,0) resp. (0,1), (1,2), (2,3), (3,0)); new vertices = old vertices, then one midpoint per edge (id nv + edge),
// then one centre per quadrilateral; element i becomes its four children in the order
//   triangle (v0, e0, e2), (e1, e2, e0), (e0, v1, e1), (e2, e1, v2)
//   quad     (v0, e0, c, e3), (e0, v1, e1, c), (c, e1, v2, e2), (e3, c, e2, v3)
// and a boundary segment (a, b) becomes (a, m), (m, b) with its attribute.  Averages as in Mesh::AverageVertices: the
// coordinates are summed in the listed order and multiplied by 1/n.
int cdm_mesh_uniform_refine(cdm_ctx *ctx, const cdm_mesh *m, cdm_mesh **refined)
{
   if (!m || !refined) { return cdm_fail(ctx, CDM_EINVAL, "cdm_mesh_uniform_refine: bad arguments"); }
   if (m->is_part) { return cdm_fail(ctx, CDM_EUNSUP, "cdm_mesh_uniform_refine: refine before partitioning"); }
   if (m->dim == 3 && m->geom == 0) { return refine_hexes(ctx, m, refined); }
   if (m->dim != 2) { return cdm_fail(ctx, CDM_EUNSUP, "cdm_mesh_uniform_refine: triangles, quadrilaterals and hexahedra only"); }
   const int nvpe = (m->geom == 1) ? 3 : 4;
   // edge ids in first-encounter order: rows keyed by the smaller vertex, chained entries (as in host_space.cpp)
   struct Node { int32_t hi, id, next; };
   std::vector<int32_t> head((size_t)m->nv, -1);
   std::vector<Node> pool;
   pool.reserve((size_t)m->ne * 2 + (size_t)m->nv);
   int32_t nedges = 0;
   auto edge = [&](int32_t a, int32_t b, bool insert) -> int32_t
   {
      const int32_t lo = std::min(a, b), hi = std::max(a, b);
      for (int32_t n = head[lo]; n >= 0; n = pool[n].next) { if (pool[n].hi == hi) { return pool[n].id; } }
      if (!insert) { return -1; }
      pool.push_back(Node{hi, nedges, head[lo]});
      head[lo] = (int32_t)pool.size() - 1;
      return nedges++;
   };
   std::vector<int32_t> el_edge((size_t)m->ne * nvpe);
   for (int64_t e = 0; e < m->ne; e++)
   {
      const int32_t *v = &m->ev[(size_t)e * nvpe];
      for (int k = 0; k < nvpe; k++) { el_edge[(size_t)e * nvpe + k] = edge(v[k], v[(k + 1) % nvpe], true); }
   }
   const int64_t nquad = (m->geom == 1) ? 0 : m->ne;
   const int64_t oedge = m->nv, oelem = oedge + nedges, nv2 = oelem + nquad;
   if (nv2 > 2147483000LL || m->ne * 4 > 2147483000LL) { return cdm_fail(ctx, CDM_EUNSUP, "cdm_mesh_uniform_refine: more than 2^31 vertices or elements"); }
   cdm_mesh *r = new (std::nothrow) cdm_mesh;
   if (!r) { return cdm_fail(ctx, CDM_ENOMEM, "out of memory"); }
   r->geom = m->geom; r->dim = 2; r->nv = nv2; r->ne = m->ne * 4; r->nbe = m->nbe * 2;
   r->vx.assign((size_t)nv2 * 2, 0.0);
   std::copy(m->vx.begin(), m->vx.end(), r->vx.begin());
   r->ev.resize((size_t)r->ne * nvpe);
   auto average = [&](const int32_t *idx, int n, int64_t dst)
   {
      for (int c = 0; c < 2; c++)
      {
         double s = 0.0;
         for (int k = 0; k < n; k++) { s += r->vx[(size_t)idx[k] * 2 + c]; }
         r->vx[(size_t)dst * 2 + c] = s * (1.0 / n);
      }
   };
   for (int64_t e = 0; e < m->ne; e++)
   {
      const int32_t *v = &m->ev[(size_t)e * nvpe];
      int32_t em[4];
      for (int k = 0; k < nvpe; k++)
      {
         em[k] = (int32_t)(oedge + el_edge[(size_t)e * nvpe + k]);
         const int32_t pair[2] = {v[k], v[(k + 1) % nvpe]};
         average(pair, 2, em[k]);
      }
      int32_t *o = &r->ev[(size_t)e * 4 * nvpe];
      if (m->geom == 1)
      {
         const int32_t ch[12] = {v[0], em[0], em[2],  em[1], em[2], em[0],  em[0], v[1], em[1],  em[2], em[1], v[2]};
         std::copy(ch, ch + 12, o);
      }
      else
      {
         const int32_t c = (int32_t)(oelem + e);
         average(v, 4, c);
         const int32_t ch[16] = {v[0], em[0], c, em[3],  em[0], v[1], em[1], c,  c, em[1], v[2], em[2],  em[3], c, em[2], v[3]};
         std::copy(ch, ch + 16, o);
      }
   }
   r->bv.resize((size_t)r->nbe * 2);
   r->battr.resize((size_t)r->nbe);
   for (int64_t b = 0; b < m->nbe; b++)
   {
      const int32_t a0 = m->bv[(size_t)b * 2], a1 = m->bv[(size_t)b * 2 + 1];
      const int32_t id = edge(a0, a1, false);
      if (id < 0) { delete r; return cdm_fail(ctx, CDM_EINVAL, "cdm_mesh_uniform_refine: boundary element is not an edge of the mesh"); }
      const int32_t mid = (int32_t)(oedge + id);
      r->bv[(size_t)b * 4 + 0] = a0; r->bv[(size_t)b * 4 + 1] = mid;
      r->bv[(size_t)b * 4 + 2] = mid; r->bv[(size_t)b * 4 + 3] = a1;
      r->battr[(size_t)b * 2] = r->battr[(size_t)b * 2 + 1] = m->battr[(size_t)b];
   }
   *refined = r;
   return CDM_OK;
}

int cdm_mesh_geometry(const cdm_mesh *m, int *geom, int *verts_per_elem, int *verts_per_bdr)
{
   if (!m) { return CDM_EINVAL; }
   if (geom) { *geom = m->geom; }
   if (verts_per_elem) { *verts_per_elem = m->geom == 1 ? m->dim + 1 : (1 << m->dim); }
   if (verts_per_bdr) { *verts_per_bdr = m->geom == 1 ? m->dim : (1 << (m->dim - 1)); }
   return CDM_OK;
}

// ------------------------------------------------------------------------------------------- Gmsh 2.2 ASCII
// Elements of the highest dimension present become the mesh elements (physical tag -> attribute), elements one
// dimension lower the boundary elements; vertices keep their file order, unused ones are dropped.
// mark_for_refinement != 0 (the `refine` argument of mfem::Mesh(file, generate_edges, refine)): every triangle is rotated
// so that its longest edge comes first, ties broken by the global (length, edge index) order.
int cdm_mesh_read_gmsh(cdm_ctx *ctx, const char *path, int mark_for_refinement, cdm_mesh **mesh)
{
   if (!path || !mesh) { return cdm_fail(ctx, CDM_EINVAL, "cdm_mesh_read_gmsh: bad arguments"); }
   std::ifstream in(path);
   if (!in) { return cdm_fail(ctx, CDM_EINVAL, std::string("cdm_mesh_read_gmsh: cannot open ") + path); }
   std::string line;
   std::map<long long, int32_t> node_index;
   std::vector<double> xyz;
   struct El { int type, attr; int32_t v[8]; };
   std::vector<El> els;
   static const int NV[16] = {0, 2, 3, 4, 4, 8, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1};
   bool ok_format = false;
   while (std::getline(in, line))
   {
      if (line.rfind("$MeshFormat", 0) == 0)
      {
         double ver; int ftype, dsize;
         in >> ver >> ftype >> dsize;
         if (ver < 2.0 || ver >= 3.0 || ftype != 0) { return cdm_fail(ctx, CDM_EUNSUP, "cdm_mesh_read_gmsh: only Gmsh 2.x ASCII files are supported"); }
         ok_format = true;
      }
      else if (line.rfind("$Nodes", 0) == 0)
      {
         long long nn; in >> nn;
         xyz.reserve((size_t)nn * 3);
         for (long long i = 0; i < nn; i++)
         {
            long long id; double x, y, z;
            in >> id >> x >> y >> z;
            node_index[id] = (int32_t)i;
            xyz.push_back(x); xyz.push_back(y); xyz.push_back(z);
         }
      }
      else if (line.rfind("$Elements", 0) == 0)
      {
         long long ne; in >> ne;
         for (long long i = 0; i < ne; i++)
         {
            long long id; int type, ntags;
            in >> id >> type >> ntags;
            int attr = 1;
            for (int t = 0; t < ntags; t++) { int tag; in >> tag; if (t == 0) { attr = tag; } }
            if (type < 1 || type > 15 || NV[type] == 0) { return cdm_fail(ctx, CDM_EUNSUP, "cdm_mesh_read_gmsh: unsupported element type " + std::to_string(type)); }
            El e; e.type = type; e.attr = attr > 0 ? attr : 1;
            for (int k = 0; k < NV[type]; k++)
            {
               long long nid; in >> nid;
               auto it = node_index.find(nid);
               if (it == node_index.end()) { return cdm_fail(ctx, CDM_EINVAL, "cdm_mesh_read_gmsh: element refers to an unknown node"); }
               e.v[k] = it->second;
            }
            els.push_back(e);
         }
      }
   }
   if (!ok_format || xyz.empty() || els.empty()) { return cdm_fail(ctx, CDM_EINVAL, "cdm_mesh_read_gmsh: not a Gmsh 2.x mesh file"); }
   bool has_hex = false, has_tet = false, has_tri = false, has_quad = false;
   for (const El &e : els) { has_hex |= e.type == 5; has_tet |= e.type == 4; has_tri |= e.type == 2; has_quad |= e.type == 3; }
   if (has_tet) { return cdm_fail(ctx, CDM_EUNSUP, "cdm_mesh_read_gmsh: tetrahedra are not supported"); }
   const int dim = has_hex ? 3 : 2;
   if (dim == 2 && has_tri && has_quad) { return cdm_fail(ctx, CDM_EUNSUP, "cdm_mesh_read_gmsh: mixed triangle / quadrilateral meshes are not supported"); }
   if (dim == 2 && !has_tri && !has_quad) { return cdm_fail(ctx, CDM_EINVAL, "cdm_mesh_read_gmsh: no 2-D or 3-D elements"); }
   const int et = dim == 3 ? 5 : (has_tri ? 2 : 3), bt = dim == 3 ? 3 : 1;
   cdm_mesh *m = new (std::nothrow) cdm_mesh;
   if (!m) { return cdm_fail(ctx, CDM_ENOMEM, "out of memory"); }
   m->dim = dim; m->geom = (et == 2) ? 1 : 0;
   // used vertices, file order
   std::vector<int32_t> newid(xyz.size() / 3, -1);
   for (const El &e : els) if (e.type == et || e.type == bt) { for (int k = 0; k < NV[e.type]; k++) { newid[e.v[k]] = 0; } }
   int32_t nv = 0;
   for (auto &id : newid) if (id == 0) { id = nv++; }
   m->nv = nv;
   m->vx.resize((size_t)nv * dim);
   for (size_t i = 0; i < newid.size(); i++)
      if (newid[i] >= 0) { for (int c = 0; c < dim; c++) { m->vx[(size_t)newid[i] * dim + c] = xyz[i * 3 + c]; } }
   if (dim == 2)
   {
      for (size_t i = 0; i < newid.size(); i++)
         if (newid[i] >= 0 && std::fabs(xyz[i * 3 + 2]) > 1e-12) { delete m; return cdm_fail(ctx, CDM_EUNSUP, "cdm_mesh_read_gmsh: surface meshes in 3-D space are not supported"); }
   }
   for (const El &e : els)
   {
      if (e.type == et) { for (int k = 0; k < NV[et]; k++) { m->ev.push_back(newid[e.v[k]]); } }
      else if (e.type == bt) { for (int k = 0; k < NV[bt]; k++) { m->bv.push_back(newid[e.v[k]]); } m->battr.push_back(e.attr); }
   }
   m->ne = (int64_t)m->ev.size() / NV[et];
   m->nbe = (int64_t)m->battr.size();
   // counter-clockwise triangles / quads (Mesh::CheckElementOrientation with fix_it = true)
   if (dim == 2)
   {
      const int k = NV[et];
      for (int64_t e = 0; e < m->ne; e++)
      {
         int32_t *v = &m->ev[(size_t)e * k];
         const double *a = &m->vx[(size_t)v[0] * 2], *b = &m->vx[(size_t)v[1] * 2], *c = &m->vx[(size_t)v[2] * 2];
         if ((b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0]) < 0.0)
         {
            if (k == 3) { std::swap(v[0], v[1]); } else { std::swap(v[1], v[3]); }
         }
      }
   }
   if (m->geom == 1 && mark_for_refinement)
   {
      // global order of the edges by (length, first-encounter edge index); the marked edge of a triangle is the one
      // with the largest rank, and the triangle is rotated so that it becomes (v0, v1)
      static const int ED[3][2] = {{0, 1}, {1, 2}, {2, 0}};
      std::map<std::pair<int32_t, int32_t>, int32_t> eidx;
      std::vector<std::pair<double, int32_t>> len;
      std::vector<int32_t> tri_edge((size_t)m->ne * 3);
      for (int64_t e = 0; e < m->ne; e++)
         for (int k = 0; k < 3; k++)
         {
            int32_t a = m->ev[e * 3 + ED[k][0]], b = m->ev[e * 3 + ED[k][1]];
            if (a > b) { std::swap(a, b); }
            auto it = eidx.find({a, b});
            if (it == eidx.end())
            {
               it = eidx.emplace(std::make_pair(a, b), (int32_t)len.size()).first;
               const double dx = m->vx[(size_t)a * 2] - m->vx[(size_t)b * 2], dy = m->vx[(size_t)a * 2 + 1] - m->vx[(size_t)b * 2 + 1];
               len.push_back({std::sqrt(dx * dx + dy * dy), it->second});
            }
            tri_edge[e * 3 + k] = it->second;
         }
      std::sort(len.begin(), len.end());
      std::vector<int32_t> rank(len.size());
      for (size_t i = 0; i < len.size(); i++) { rank[len[i].second] = (int32_t)i; }
      for (int64_t e = 0; e < m->ne; e++)
      {
         int32_t *v = &m->ev[(size_t)e * 3];
         const int32_t r0 = rank[tri_edge[e * 3]], r1 = rank[tri_edge[e * 3 + 1]], r2 = rank[tri_edge[e * 3 + 2]];
         if (r0 >= r1 && r0 >= r2) { continue; }
         if (r1 >= r2) { const int32_t t = v[0]; v[0] = v[1]; v[1] = v[2]; v[2] = t; }          // (1,2) first
         else { const int32_t t = v[2]; v[2] = v[1]; v[1] = v[0]; v[0] = t; }                  // (2,0) first
      }
   }
   *mesh = m;
   return CDM_OK;
}

// ------------------------------------------------------------------------------------------- YAML / PETSc options
struct cdm_config { std::map<std::string, std::string> kv; };

static std::string trim(const std::string &s)
{
   const size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
   return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}

// The flat `key: value` subset of YAML the drivers use (Input/*.yaml): scalars, quoted strings, `[a, b, c]` flow
// sequences, `#` comments.  Nested maps / block sequences are rejected (CDM_EUNSUP).
int cdm_config_load(const char *path, cdm_config **out)
{
   if (!path || !out) { return CDM_EINVAL; }
   *out = nullptr;
   std::ifstream in(path);
   if (!in) { return CDM_EINVAL; }
   cdm_config *c = new (std::nothrow) cdm_config;
   if (!c) { return CDM_ENOMEM; }
   std::string line;
   while (std::getline(in, line))
   {
      bool inq = false; char qc = 0;
      size_t cut = std::string::npos;
      for (size_t i = 0; i < line.size(); i++)
      {
         const char ch = line[i];
         if (inq) { if (ch == qc) { inq = false; } }
         else if (ch == '"' || ch == '\'') { inq = true; qc = ch; }
         else if (ch == '#' && (i == 0 || line[i - 1] == ' ' || line[i - 1] == '\t')) { cut = i; break; }
      }
      if (cut != std::string::npos) { line = line.substr(0, cut); }
      const std::string t = trim(line);
      if (t.empty() || t == "---") { continue; }
      if (line[0] == ' ' || line[0] == '\t' || t[0] == '-') { delete c; return CDM_EUNSUP; }     // nested structure
      const size_t colon = line.find(':');
      if (colon == std::string::npos) { delete c; return CDM_EINVAL; }
      std::string key = trim(line.substr(0, colon)), val = trim(line.substr(colon + 1));
      if (val.size() >= 2 && (val.front() == '"' || val.front() == '\'') && val.back() == val.front()) { val = val.substr(1, val.size() - 2); }
      c->kv[key] = val;
   }
   *out = c;
   return CDM_OK;
}
int cdm_config_destroy(cdm_config *c) { delete c; return CDM_OK; }
int cdm_config_has(const cdm_config *c, const char *key) { return (c && key && c->kv.count(key)) ? 1 : 0; }
int cdm_config_get_string(const cdm_config *c, const char *key, char *buf, int buflen)
{
   if (!c || !key || !buf || buflen < 1) { return CDM_EINVAL; }
   auto it = c->kv.find(key);
   if (it == c->kv.end()) { return CDM_EINVAL; }
   std::snprintf(buf, (size_t)buflen, "%s", it->second.c_str());
   return CDM_OK;
}
int cdm_config_get_double(const cdm_config *c, const char *key, double *v)
{
   if (!c || !key || !v) { return CDM_EINVAL; }
   auto it = c->kv.find(key);
   if (it == c->kv.end()) { return CDM_EINVAL; }
   char *end = nullptr;
   const double d = std::strtod(it->second.c_str(), &end);
   if (end == it->second.c_str() || !trim(end).empty()) { return CDM_EINVAL; }
   *v = d;
   return CDM_OK;
}
int cdm_config_get_int(const cdm_config *c, const char *key, int *v)
{
   double d;
   const int rc = cdm_config_get_double(c, key, &d);
   if (rc) { return rc; }
   if (d != std::floor(d)) { return CDM_EINVAL; }
   *v = (int)d;
   return CDM_OK;
}
int cdm_config_get_bool(const cdm_config *c, const char *key, int *v)
{
   if (!c || !key || !v) { return CDM_EINVAL; }
   auto it = c->kv.find(key);
   if (it == c->kv.end()) { return CDM_EINVAL; }
   std::string s = it->second;
   std::transform(s.begin(), s.end(), s.begin(), ::tolower);
   if (s == "true" || s == "yes" || s == "on" || s == "1") { *v = 1; return CDM_OK; }
   if (s == "false" || s == "no" || s == "off" || s == "0") { *v = 0; return CDM_OK; }
   return CDM_EINVAL;
}
// `[a, b, c]` -> up to maxn doubles; *n = number of entries in the file
int cdm_config_get_doubles(const cdm_config *c, const char *key, double *v, int maxn, int *n)
{
   if (!c || !key || !n) { return CDM_EINVAL; }
   auto it = c->kv.find(key);
   if (it == c->kv.end()) { return CDM_EINVAL; }
   std::string s = trim(it->second);
   if (s.size() < 2 || s.front() != '[' || s.back() != ']') { return CDM_EINVAL; }
   s = s.substr(1, s.size() - 2);
   std::stringstream ss(s);
   std::string tok;
   int k = 0;
   while (std::getline(ss, tok, ','))
   {
      tok = trim(tok);
      if (tok.empty()) { continue; }
      char *end = nullptr;
      const double d = std::strtod(tok.c_str(), &end);
      if (end == tok.c_str()) { return CDM_EINVAL; }
      if (v && k < maxn) { v[k] = d; }
      k++;
   }
   *n = k;
   return CDM_OK;
}

// PETSc options file (Input/petsc.opts, Input/petsc_circle.opts): -ksp_type, -ksp_rtol, -ksp_atol, -ksp_max_it,
// -ksp_gmres_restart, -pc_type {none, jacobi, bjacobi, ilu}, -sub_pc_type ilu.  Unknown options are ignored (as PETSc
// ignores options nobody queries).  ksp_type: 0 gmres, 1 cg.  The preconditioner lands in opts->jacobi
// (0 none, 1 jacobi, 2 block-Jacobi + ILU(0)).
int cdm_petsc_options_load(const char *path, cdm_krylov_opts *opts, int *ksp_type)
{
   if (!path || !opts) { return CDM_EINVAL; }
   std::ifstream in(path);
   if (!in) { return CDM_EINVAL; }
   // KSP defaults: gmres(30), rtol 1e-5, atol 1e-50, max_it 10000, left preconditioning, ILU(0) on one rank
   opts->variant = CDM_GMRES_PETSC; opts->restart = 30; opts->max_it = 10000; opts->rtol = 1e-5; opts->atol = 1e-50;
   opts->zero_guess = 1; opts->jacobi = 2;
   int kt = 0;
   std::string pc, sub_pc = "ilu", line;
   while (std::getline(in, line))
   {
      const size_t h = line.find('#');
      if (h != std::string::npos) { line = line.substr(0, h); }
      std::stringstream ss(line);
      std::vector<std::string> tok;
      std::string w;
      while (ss >> w) { tok.push_back(w); }
      for (size_t i = 0; i < tok.size(); i++)
      {
         const std::string &name = tok[i];
         if (name.size() < 2 || name[0] != '-' || std::isdigit((unsigned char)name[1])) { continue; }
         std::string val;
         if (i + 1 < tok.size() && !(tok[i + 1].size() > 1 && tok[i + 1][0] == '-' && !std::isdigit((unsigned char)tok[i + 1][1]) && tok[i + 1][1] != '.'))
         { val = tok[++i]; }
         if (name == "-ksp_type") { if (val == "gmres") { kt = 0; } else if (val == "cg") { kt = 1; } else { return CDM_EUNSUP; } }
         else if (name == "-ksp_rtol") { opts->rtol = std::atof(val.c_str()); }
         else if (name == "-ksp_atol") { opts->atol = std::atof(val.c_str()); }
         else if (name == "-ksp_max_it") { opts->max_it = std::atoi(val.c_str()); }
         else if (name == "-ksp_gmres_restart") { opts->restart = std::atoi(val.c_str()); }
         else if (name == "-pc_type") { pc = val; }
         else if (name == "-sub_pc_type") { sub_pc = val; }
      }
   }
   if (!pc.empty())
   {
      if (pc == "none") { opts->jacobi = 0; }
      else if (pc == "jacobi") { opts->jacobi = 1; }
      else if (pc == "ilu" || (pc == "bjacobi" && sub_pc == "ilu")) { opts->jacobi = 2; }
      else { return CDM_EUNSUP; }
   }
   if (ksp_type) { *ksp_type = kt; }
   return CDM_OK;
}

}  // extern "C"
