// host_space.cpp -- host-side (setup-only) builders: 1-D rules and basis tables,
// H1 global numbering, ElementRestriction index arrays.
//
// Stands behind  H1_FECollection(order, dim) / ParFiniteElementSpace
// (linear_convection_diffusion_2D.cpp:311-312) and the ElementRestriction that a
// partially assembled BilinearForm builds from it (MFEM fem/restriction.cpp,
// upstream, not vendored).  Numbering conventions: SURVEY.md Appendix C.1-C.3.
#include "cdm_internal.hpp"
#include <chrono>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cmath>
#include <cstring>

// ------------------------------------------------------------------- rules
namespace
{
// P_n(z) and P_n'(z) via the Bonnet recurrence carried on (P_k, P_k')
inline void leg(int n, double z, double &P, double &dP)
{
   double pm = 1.0, pc = z, dm = 0.0, dc = 1.0;
   if (n == 0) { P = 1.0; dP = 0.0; return; }
   for (int k = 1; k < n; k++)
   {
      const double pn = ((2 * k + 1) * z * pc - k * pm) / (k + 1);
      const double dn = dm + (2 * k + 1) * pc;      // P'_{k+1} = P'_{k-1} + (2k+1) P_k
      pm = pc; pc = pn; dm = dc; dc = dn;
   }
   P = pc; dP = dc;
}
}

void cdm_host_gauss_legendre(int n, double *x, double *w)
{
   for (int i = 0; i < n; i++)
   {
      // Tricomi initial guess for the i-th root (descending), polished by Newton
      const double th = M_PI * (4.0 * (i + 1) - 1.0) / (4.0 * n + 2.0);
      double z = (1.0 - (n - 1.0) / (8.0 * n * n * n)) * std::cos(th);
      double P, dP;
      for (int it = 0; it < 50; it++)
      {
         leg(n, z, P, dP);
         const double dz = P / dP;
         z -= dz;
         if (std::fabs(dz) < 2e-16) { break; }
      }
      leg(n, z, P, dP);
      x[n - 1 - i] = 0.5 * (1.0 + z);
      w[n - 1 - i] = 1.0 / ((1.0 - z * z) * dP * dP);
   }
   // symmetrise (the rule is symmetric about 1/2)
   for (int i = 0; i < n / 2; i++)
   {
      const double xm = 0.5 * (x[i] + (1.0 - x[n - 1 - i]));
      x[i] = xm; x[n - 1 - i] = 1.0 - xm;
      const double wm = 0.5 * (w[i] + w[n - 1 - i]);
      w[i] = w[n - 1 - i] = wm;
   }
   if (n & 1) { x[n / 2] = 0.5; }
}

void cdm_host_gauss_lobatto(int n, double *x)
{
   if (n == 1) { x[0] = 0.5; return; }
   const int m = n - 1;
   x[0] = 0.0; x[m] = 1.0;
   for (int i = 1; i < m; i++)
   {
      double z = -std::cos(M_PI * i / m);           // ascending
      for (int it = 0; it < 50; it++)
      {
         double P, dP;
         leg(m, z, P, dP);
         const double d2P = (2.0 * z * dP - m * (m + 1.0) * P) / (1.0 - z * z);
         const double dz = dP / d2P;
         z -= dz;
         if (std::fabs(dz) < 2e-16) { break; }
      }
      x[i] = 0.5 * (1.0 + z);
   }
   for (int i = 1; i <= (m - 1) / 2; i++)
   {
      const double xm = 0.5 * (x[i] + (1.0 - x[m - i]));
      x[i] = xm; x[m - i] = 1.0 - xm;
   }
   if (n & 1) { x[m / 2] = 0.5; }
}

int cdm_host_q1d(int dim, int p)
{
   // DiffusionIntegrator / ConvectionIntegrator / MassIntegrator ::GetRule on a
   // Q1 (bi/tri-linear) mesh: order 2p+dim-1; IntRules.Get -> order/2+1 points
   return (2 * p + dim - 1) / 2 + 1;
}

// Lagrange basis on GLL nodes at the Gauss points, barycentric form
void cdm_host_basis(int p, int q1d, double *B, double *G, double *qw, double *nodes, double *qx)
{
   const int d1d = p + 1;
   std::vector<double> xn(d1d), xq(q1d), wq(q1d), bw(d1d);
   cdm_host_gauss_lobatto(d1d, xn.data());
   cdm_host_gauss_legendre(q1d, xq.data(), wq.data());
   for (int j = 0; j < d1d; j++)
   {
      double c = 1.0;
      for (int k = 0; k < d1d; k++) if (k != j) { c *= (xn[j] - xn[k]); }
      bw[j] = 1.0 / c;
   }
   for (int q = 0; q < q1d; q++)
   {
      const double x = xq[q];
      int hit = -1;
      for (int j = 0; j < d1d; j++) if (x == xn[j]) { hit = j; }
      if (hit < 0)
      {
         double L = 1.0, S = 0.0;                      // L = prod (x-x_k), S = sum 1/(x-x_k)
         for (int k = 0; k < d1d; k++) { L *= (x - xn[k]); S += 1.0 / (x - xn[k]); }
         for (int j = 0; j < d1d; j++)
         {
            const double lj = L * bw[j] / (x - xn[j]);
            B[q * d1d + j] = lj;
            G[q * d1d + j] = lj * (S - 1.0 / (x - xn[j]));
         }
      }
      else
      {
         // evaluation point coincides with node `hit`
         for (int j = 0; j < d1d; j++)
         {
            B[q * d1d + j] = (j == hit) ? 1.0 : 0.0;
            if (j != hit) { G[q * d1d + j] = (bw[j] / bw[hit]) / (xn[hit] - xn[j]); }
         }
         double s = 0.0;
         for (int k = 0; k < d1d; k++) if (k != hit) { s += 1.0 / (xn[hit] - xn[k]); }
         G[q * d1d + hit] = s;
      }
      if (qw) { qw[q] = wq[q]; }
      if (qx) { qx[q] = xq[q]; }
   }
   if (nodes) { for (int j = 0; j < d1d; j++) { nodes[j] = xn[j]; } }
}

// --------------------------------------------------------------- numbering
namespace
{
// phase times of the host-side setup on stderr when CDM_CFG_DEBUG is set
struct PhaseTimer
{
   bool on; std::chrono::steady_clock::time_point t0;
   PhaseTimer() : on(getenv("CDM_CFG_DEBUG") != nullptr), t0(std::chrono::steady_clock::now()) {}
   void lap(const char *what)
   {
      if (!on) { return; }
      const auto t1 = std::chrono::steady_clock::now();
      fprintf(stderr, "[cdm] host setup: %s %.3f s\n", what, std::chrono::duration<double>(t1 - t0).count());
      t0 = t1;
   }
};

const int HEX_E[12][2] = {{0,1},{1,2},{3,2},{0,3},{4,5},{5,6},{7,6},{4,7},{0,4},{1,5},{2,6},{3,7}};
const int HEX_F[6][4] = {{3,2,1,0},{0,1,5,4},{1,2,6,5},{2,3,7,6},{3,0,4,7},{4,5,6,7}};
const int QUAD_E[4][2] = {{0,1},{1,2},{2,3},{3,0}};

// flat open-addressing map: 64-bit key -> id in first-insertion order
// Edge and face tables: first-insertion ids, rows keyed by the smallest vertex id of the entity, the entries of a row chained
// through a node pool in insertion order (the shape of MFEM's DSTable / STable3D).  A mesh whose vertices are numbered with
// spatial locality touches rows and nodes it created a few elements ago, so the build runs out of cache: open-addressing hash
// tables (rounds 1-2) scattered 66 M probes of a 30 M-dof mesh over 1 GB and took 8 of the 11.5 s of the space setup.
struct EdgeMap
{
   struct Node { int32_t hi, id, next; };
   std::vector<int32_t> head; std::vector<Node> pool; int32_t count = 0;
   EdgeMap(size_t nv, size_t expect) { head.assign(nv, -1); pool.reserve(expect); }
   int32_t get(int32_t a, int32_t b, bool insert)
   {
      const int32_t lo = std::min(a, b), hi = std::max(a, b);
      for (int32_t n = head[lo]; n >= 0; n = pool[n].next) { if (pool[n].hi == hi) { return pool[n].id; } }
      if (!insert) { return -1; }
      pool.push_back(Node{hi, count, head[lo]});
      head[lo] = (int32_t)pool.size() - 1;
      return count++;
   }
};

// quad faces keyed by their three smallest vertex ids (unique in a conforming mesh)
struct FaceMap
{
   struct Node { int32_t b, c, id, next; };
   std::vector<int32_t> head; std::vector<Node> pool; int32_t count = 0;
   std::vector<int32_t> base;    // 4 vertices per face, as seen by the creating element
   FaceMap(size_t nv, size_t expect) { head.assign(nv, -1); pool.reserve(expect); base.reserve(4 * expect); }
   // the three smallest of four vertex ids, ascending (five compare-exchanges)
   static void smallest3(const int32_t *v, int32_t *s4)
   {
      int32_t a = v[0], b = v[1], c = v[2], d = v[3], t;
#define CDM_CSWAP(x, y) if (x > y) { t = x; x = y; y = t; }
      CDM_CSWAP(a, b) CDM_CSWAP(c, d) CDM_CSWAP(a, c) CDM_CSWAP(b, d) CDM_CSWAP(b, c)
#undef CDM_CSWAP
      s4[0] = a; s4[1] = b; s4[2] = c; s4[3] = d;
   }
   int32_t get(const int32_t *v, bool insert)
   {
      int32_t s4[4];
      smallest3(v, s4);
      for (int32_t n = head[s4[0]]; n >= 0; n = pool[n].next)
      {
         if (pool[n].b == s4[1] && pool[n].c == s4[2]) { return pool[n].id; }
      }
      if (!insert) { return -1; }
      pool.push_back(Node{s4[1], s4[2], count, head[s4[0]]});
      head[s4[0]] = (int32_t)pool.size() - 1;
      base.insert(base.end(), v, v + 4);
      return count++;
   }
};

// orientation of `test` relative to `base` (MFEM Mesh::GetQuadOrientation) and the
// induced map of interior face dofs (H1_FECollection QuadDofOrd)
inline int quad_ori(const int32_t *base, const int32_t *test)
{
   int i = 0;
   while (i < 4 && test[i] != base[0]) { i++; }
   return (test[(i + 1) & 3] == base[1]) ? 2 * i : 2 * i + 1;
}
inline int quad_perm(int ori, int m, int i, int j)   // m = p-1
{
   const int r = m - 1;
   switch (ori)
   {
      case 0: return i + j * m;
      case 1: return j + i * m;
      case 2: return j + (r - i) * m;
      case 3: return (r - i) + j * m;
      case 4: return (r - i) + (r - j) * m;
      case 5: return (r - j) + (r - i) * m;
      case 6: return (r - j) + i * m;
      default: return i + (r - j) * m;
   }
}

// native (vertices, edges, faces, interior) index of every lexicographic node
void lex_to_native(int dim, int p, std::vector<int> &map)
{
   const int p1 = p + 1;
   map.assign(dim == 2 ? p1 * p1 : p1 * p1 * p1, -1);
   int o = 0;
   if (dim == 2)
   {
      auto L = [&](int i, int j) -> int & { return map[i + p1 * j]; };
      L(0, 0) = o++; L(p, 0) = o++; L(p, p) = o++; L(0, p) = o++;
      for (int i = 1; i < p; i++) { L(i, 0) = o++; }
      for (int i = 1; i < p; i++) { L(p, i) = o++; }
      for (int i = 1; i < p; i++) { L(p - i, p) = o++; }
      for (int i = 1; i < p; i++) { L(0, p - i) = o++; }
      for (int j = 1; j < p; j++) for (int i = 1; i < p; i++) { L(i, j) = o++; }
      return;
   }
   auto L = [&](int i, int j, int k) -> int & { return map[i + p1 * (j + p1 * k)]; };
   const int c[8][3] = {{0,0,0},{p,0,0},{p,p,0},{0,p,0},{0,0,p},{p,0,p},{p,p,p},{0,p,p}};
   for (int v = 0; v < 8; v++) { L(c[v][0], c[v][1], c[v][2]) = o++; }
   // edges: running index from the first to the second local vertex
   for (int e = 0; e < 12; e++)
   {
      const int *a = c[HEX_E[e][0]], *b = c[HEX_E[e][1]];
      for (int i = 1; i < p; i++)
         L(a[0] + (b[0] - a[0]) / p * i, a[1] + (b[1] - a[1]) / p * i, a[2] + (b[2] - a[2]) / p * i) = o++;
   }
   // faces: (i,j) runs along v0->v1 and v0->v3 of the local face
   for (int f = 0; f < 6; f++)
   {
      const int *v0 = c[HEX_F[f][0]], *v1 = c[HEX_F[f][1]], *v3 = c[HEX_F[f][3]];
      for (int j = 1; j < p; j++)
         for (int i = 1; i < p; i++)
         {
            int xyz[3];
            for (int d = 0; d < 3; d++) { xyz[d] = v0[d] + (v1[d] - v0[d]) / p * i + (v3[d] - v0[d]) / p * j; }
            L(xyz[0], xyz[1], xyz[2]) = o++;
         }
   }
   for (int k = 1; k < p; k++) for (int j = 1; j < p; j++) for (int i = 1; i < p; i++) { L(i, j, k) = o++; }
}
}  // namespace

// Returns ndof, fills elem_dof (lexicographic within each element) and, per
// boundary element, the list of dofs lying on it.
int64_t cdm_host_h1_numbering(const cdm_mesh &m, int p, std::vector<int32_t> &elem_dof,
                              std::vector<int32_t> &bdr_off, std::vector<int32_t> &bdr_flat,
                              int64_t *class_off)
{
   const int dim = m.dim, p1 = p + 1, pm1 = p - 1;
   const int nd = (dim == 2) ? p1 * p1 : p1 * p1 * p1;
   const int nvpe = (dim == 2) ? 4 : 8, nepe = (dim == 2) ? 4 : 12;
   const bool need_tabs = pm1 > 0;
   EdgeMap em(need_tabs ? (size_t)m.nv : 0, need_tabs ? (size_t)(dim == 2 ? 2 : 3) * m.ne + m.nv : 0);
   FaceMap fm((need_tabs && dim == 3) ? (size_t)m.nv : 0, (need_tabs && dim == 3) ? (size_t)3 * m.ne + m.nbe : 0);
   std::vector<int32_t> e_edge, e_face;        // per element entity ids (first pass)
   PhaseTimer tm_;
   if (need_tabs)
   {
      e_edge.resize((size_t)m.ne * nepe);
      if (dim == 3) { e_face.resize((size_t)m.ne * 6); }
      for (int64_t e = 0; e < m.ne; e++)
      {
         const int32_t *v = &m.ev[(size_t)e * nvpe];
         for (int k = 0; k < nepe; k++)
         {
            const int *le = (dim == 2) ? QUAD_E[k] : HEX_E[k];
            e_edge[(size_t)e * nepe + k] = em.get(v[le[0]], v[le[1]], true);
         }
      }
      tm_.lap("edge table");
      if (dim == 3)
         for (int64_t e = 0; e < m.ne; e++)
         {
            const int32_t *v = &m.ev[(size_t)e * nvpe];
            for (int k = 0; k < 6; k++)
            {
               const int32_t fv[4] = {v[HEX_F[k][0]], v[HEX_F[k][1]], v[HEX_F[k][2]], v[HEX_F[k][3]]};
               e_face[(size_t)e * 6 + k] = fm.get(fv, true);
            }
         }
   }
   tm_.lap("face table");
   const int64_t nedges = em.count, nfaces = fm.count;
   const int64_t edge0 = m.nv, face0 = edge0 + nedges * pm1;
   const int64_t int0 = face0 + nfaces * (int64_t)pm1 * pm1;
   int nint = 1; for (int d = 0; d < dim; d++) { nint *= pm1; }
   if (pm1 <= 0) { nint = 0; }
   const int64_t ndof = int0 + m.ne * nint;
   if (class_off) { class_off[0] = 0; class_off[1] = edge0; class_off[2] = face0; class_off[3] = int0; class_off[4] = ndof; }

   std::vector<int> l2n;
   lex_to_native(dim, p, l2n);
   elem_dof.resize((size_t)m.ne * nd);
   // the elements are independent here: slices of the element range on the host threads
   auto number_range = [&](int64_t eb, int64_t ee)
   {
      std::vector<int32_t> native(nd);
      for (int64_t e = eb; e < ee; e++)
      {
         const int32_t *v = &m.ev[(size_t)e * nvpe];
         int o = 0;
         for (int k = 0; k < nvpe; k++) { native[o++] = v[k]; }
         if (need_tabs)
         {
            for (int k = 0; k < nepe; k++)
            {
               const int *le = (dim == 2) ? QUAD_E[k] : HEX_E[k];
               const bool fwd = v[le[0]] < v[le[1]];
               const int64_t b = edge0 + (int64_t)e_edge[(size_t)e * nepe + k] * pm1;
               for (int i = 0; i < pm1; i++) { native[o++] = (int32_t)(b + (fwd ? i : pm1 - 1 - i)); }
            }
            if (dim == 3)
               for (int k = 0; k < 6; k++)
               {
                  const int32_t fv[4] = {v[HEX_F[k][0]], v[HEX_F[k][1]], v[HEX_F[k][2]], v[HEX_F[k][3]]};
                  const int32_t id = e_face[(size_t)e * 6 + k];
                  const int ori = quad_ori(&fm.base[(size_t)4 * id], fv);
                  const int64_t b = face0 + (int64_t)id * pm1 * pm1;
                  for (int j = 0; j < pm1; j++)
                     for (int i = 0; i < pm1; i++) { native[o++] = (int32_t)(b + quad_perm(ori, pm1, i, j)); }
               }
            for (int i = 0; i < nint; i++) { native[o++] = (int32_t)(int0 + e * nint + i); }
         }
         int32_t *out = &elem_dof[(size_t)e * nd];
         for (int l = 0; l < nd; l++) { out[l] = native[l2n[l]]; }
      }
   };
   {
      unsigned nt = std::thread::hardware_concurrency();
      if (nt > 16) { nt = 16; }
      if (nt < 2 || m.ne < 65536) { number_range(0, m.ne); }
      else
      {
         std::vector<std::thread> th;
         for (unsigned t = 0; t < nt; t++) { th.emplace_back(number_range, m.ne * t / nt, m.ne * (t + 1) / nt); }
         for (auto &x : th) { x.join(); }
      }
   }
   tm_.lap("element dofs");
   // dofs of each boundary element (closure: vertices, edges, face interior)
   const int nvpf = (dim == 2) ? 2 : 4;
   bdr_off.assign(m.nbe + 1, 0);
   bdr_flat.clear();
   for (int64_t b = 0; b < m.nbe; b++)
   {
      const int32_t *v = &m.bv[(size_t)b * nvpf];
      for (int k = 0; k < nvpf; k++) { bdr_flat.push_back(v[k]); }
      if (need_tabs)
      {
         const int nbedge = (dim == 2) ? 1 : 4;
         for (int k = 0; k < nbedge; k++)
         {
            const int32_t id = em.get(v[k], v[(k + 1) % nvpf], false);
            if (id < 0) { return -1; }
            for (int i = 0; i < pm1; i++) { bdr_flat.push_back((int32_t)(edge0 + (int64_t)id * pm1 + i)); }
         }
         if (dim == 3)
         {
            const int32_t id = fm.get(v, false);
            if (id < 0) { return -1; }
            for (int i = 0; i < pm1 * pm1; i++) { bdr_flat.push_back((int32_t)(face0 + (int64_t)id * pm1 * pm1 + i)); }
         }
      }
      bdr_off[b + 1] = (int32_t)bdr_flat.size();
   }
   return ndof;
}

// counting sort of the (element, local dof) pairs by global dof (stable: the entries of a dof ascend).  Large inputs: the
// dof range is cut into one slice per host thread; every thread scans the whole gather map (sequential reads) and counts /
// places only the entries of its slice, whose counters (a few MB) stay in its cache -- the serial version scattered 10^8
// increments and stores over the whole dof range.
void cdm_host_restriction(int64_t ne, int nd, int64_t ndof, const std::vector<int32_t> &gather,
                          std::vector<int32_t> &offsets, std::vector<int32_t> &indices)
{
   PhaseTimer tm_;
   const int64_t n = ne * nd;
   offsets.assign(ndof + 1, 0);
   indices.resize(n);
   unsigned nt = std::thread::hardware_concurrency();
   if (nt > 16) { nt = 16; }
   if (nt < 2 || n < (int64_t)1 << 22)
   {
      for (int64_t i = 0; i < n; i++) { offsets[gather[i] + 1]++; }
      for (int64_t g = 0; g < ndof; g++) { offsets[g + 1] += offsets[g]; }
      std::vector<int32_t> cur(offsets.begin(), offsets.end() - 1);
      for (int64_t i = 0; i < n; i++) { indices[cur[gather[i]]++] = (int32_t)i; }
      tm_.lap("restriction transpose (offsets, indices)");
      return;
   }
   const int32_t *gm = gather.data();
   auto slice = [&](unsigned t) { return (int32_t)(ndof * (int64_t)t / nt); };
   auto run = [&](auto fn)
   {
      std::vector<std::thread> th;
      for (unsigned t = 0; t < nt; t++) { th.emplace_back(fn, t); }
      for (auto &x : th) { x.join(); }
   };
   run([&](unsigned t)
   {
      const int32_t lo = slice(t), hi = slice(t + 1);
      int32_t *cnt = offsets.data() + 1;
      const uint32_t w = (uint32_t)(hi - lo);
      for (int64_t i = 0; i < n; i++) { const int32_t g = gm[i]; if ((uint32_t)(g - lo) < w) { cnt[g]++; } }
   });
   for (int64_t g = 0; g < ndof; g++) { offsets[g + 1] += offsets[g]; }
   run([&](unsigned t)
   {
      const int32_t lo = slice(t), hi = slice(t + 1);
      std::vector<int32_t> cur(offsets.begin() + lo, offsets.begin() + hi);
      const uint32_t w = (uint32_t)(hi - lo);
      for (int64_t i = 0; i < n; i++) { const uint32_t r = (uint32_t)(gm[i] - lo); if (r < w) { indices[cur[r]++] = (int32_t)i; } }
   });
   tm_.lap("restriction transpose (offsets, indices), host threads");
}
