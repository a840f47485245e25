// host_simplex.cpp -- order-p H1 Lagrange triangles: node set, nodal basis tables, quadrature, global numbering.
//
// Stands behind H1_FECollection(order, 2) / ParFiniteElementSpace on the triangle meshes the reference ships
// (Input/input_2d.yaml:1-2: Mesh/unit_square.msh, 938 triangles, order 3; linear_convection_diffusion_2D.cpp:290-312)
// and the element matrices DiffusionIntegrator / ConvectionIntegrator / MassIntegrator assemble on them (:335-339).
// [MFEM-upstream, restated from memory -- MFEM is not vendored]: H1_TriangleElement places its nodes at
// (cp[i], cp[j]) / (cp[i] + cp[j] + cp[p-i-j]) with cp the Gauss-Lobatto points on [0,1]; native dof order is
// 3 vertices, 3 edges x (p-1) in the local directions (0,1), (1,2), (2,0), then the interior row by row.
// Simplices have no lexicographic reordering: the E-vector order of ElementRestriction is the native order.
// The discrete function SPACE is P_p whatever the nodes, so solutions, residual histories and error norms do not
// depend on the node set; dof VALUES and index maps do.
#include "cdm_internal.hpp"
#include <algorithm>
#include <cmath>
#include <unordered_map>

void cdm_host_tri_nodes(int p, double *xy)
{
   std::vector<double> cp(p + 1);
   cdm_host_gauss_lobatto(p + 1, cp.data());
   int o = 0;
   auto put = [&](double x, double y) { xy[2 * o] = x; xy[2 * o + 1] = y; o++; };
   put(cp[0], cp[0]); put(cp[p], cp[0]); put(cp[0], cp[p]);
   for (int i = 1; i < p; i++) { put(cp[i], cp[0]); }            // edge (0,1)
   for (int i = 1; i < p; i++) { put(cp[p - i], cp[i]); }        // edge (1,2)
   for (int i = 1; i < p; i++) { put(cp[0], cp[p - i]); }        // edge (2,0)
   for (int j = 1; j < p; j++)
      for (int i = 1; i + j < p; i++)
      {
         const double w = cp[i] + cp[j] + cp[p - i - j];
         put(cp[i] / w, cp[j] / w);
      }
}

// int_T f = int_0^1 int_0^1 f(u (1 - v), v) (1 - v) du dv with Gauss-Legendre in u and v: exact for total degree 2n-2
void cdm_host_tri_rule(int n, double *xy, double *w)
{
   std::vector<double> x1(n), w1(n);
   cdm_host_gauss_legendre(n, x1.data(), w1.data());
   int q = 0;
   for (int b = 0; b < n; b++)
      for (int a = 0; a < n; a++, q++)
      {
         xy[2 * q] = x1[a] * (1.0 - x1[b]);
         xy[2 * q + 1] = x1[b];
         w[q] = w1[a] * w1[b] * (1.0 - x1[b]);
      }
}

// phi_j = sum_k c_kj m_k with the monomials m_k = x^a y^b, a + b <= p, and C = V^{-1}, V_ik = m_k(node_i).
// The inverse is computed in long double with partial pivoting (orders <= 6: cond(V) < 1e7).
void cdm_host_tri_basis(int p, int npts, const double *xy, double *B, double *G)
{
   const int nd = (p + 1) * (p + 2) / 2;
   std::vector<double> nodes(2 * nd);
   cdm_host_tri_nodes(p, nodes.data());
   std::vector<int> ea, eb;
   for (int s = 0; s <= p; s++) for (int b = 0; b <= s; b++) { ea.push_back(s - b); eb.push_back(b); }
   std::vector<long double> M((size_t)nd * 2 * nd, 0.0L);          // [V | I]
   for (int i = 0; i < nd; i++)
   {
      for (int k = 0; k < nd; k++) { M[(size_t)i * 2 * nd + k] = std::pow((long double)nodes[2 * i], ea[k]) * std::pow((long double)nodes[2 * i + 1], eb[k]); }
      M[(size_t)i * 2 * nd + nd + i] = 1.0L;
   }
   for (int c = 0; c < nd; c++)
   {
      int piv = c;
      for (int r = c + 1; r < nd; r++) if (std::fabs(M[(size_t)r * 2 * nd + c]) > std::fabs(M[(size_t)piv * 2 * nd + c])) { piv = r; }
      if (piv != c) { for (int k = 0; k < 2 * nd; k++) { std::swap(M[(size_t)c * 2 * nd + k], M[(size_t)piv * 2 * nd + k]); } }
      const long double d = M[(size_t)c * 2 * nd + c];
      for (int k = 0; k < 2 * nd; k++) { M[(size_t)c * 2 * nd + k] /= d; }
      for (int r = 0; r < nd; r++)
      {
         if (r == c) { continue; }
         const long double f = M[(size_t)r * 2 * nd + c];
         if (f != 0.0L) { for (int k = 0; k < 2 * nd; k++) { M[(size_t)r * 2 * nd + k] -= f * M[(size_t)c * 2 * nd + k]; } }
      }
   }
   auto C = [&](int k, int j) { return M[(size_t)k * 2 * nd + nd + j]; };       // V^{-1}[k][j]
   for (int q = 0; q < npts; q++)
   {
      const long double x = xy[2 * q], y = xy[2 * q + 1];
      for (int j = 0; j < nd; j++)
      {
         long double v = 0.0L, gx = 0.0L, gy = 0.0L;
         for (int k = 0; k < nd; k++)
         {
            const int a = ea[k], b = eb[k];
            const long double xa = std::pow(x, a), yb = std::pow(y, b);
            v += C(k, j) * xa * yb;
            if (a > 0) { gx += C(k, j) * a * std::pow(x, a - 1) * yb; }
            if (b > 0) { gy += C(k, j) * b * xa * std::pow(y, b - 1); }
         }
         B[(size_t)q * nd + j] = (double)v;
         G[((size_t)0 * npts + q) * nd + j] = (double)gx;
         G[((size_t)1 * npts + q) * nd + j] = (double)gy;
      }
   }
}

// global numbering: [vertices | edges x (p-1), first-encounter order over elements and local edges (0,1),(1,2),(2,0),
// orientation low -> high global vertex | interiors, element by element]
int64_t cdm_host_h1_numbering_tri(const cdm_mesh &m, int p, std::vector<int32_t> &elem_dof,
                                  std::vector<int32_t> &bdr_off, std::vector<int32_t> &bdr_flat)
{
   const int pm1 = p - 1, nint = (p - 1) * (p - 2) / 2, nd = (p + 1) * (p + 2) / 2;
   static const int ED[3][2] = {{0, 1}, {1, 2}, {2, 0}};
   std::unordered_map<uint64_t, int32_t> edges;
   edges.reserve((size_t)m.ne * 2);
   auto key = [](int32_t a, int32_t b) { return ((uint64_t)(uint32_t)std::min(a, b) << 32) | (uint32_t)std::max(a, b); };
   std::vector<int32_t> e_edge((size_t)m.ne * 3);
   for (int64_t e = 0; e < m.ne; e++)
      for (int k = 0; k < 3; k++)
      {
         const int32_t a = m.ev[e * 3 + ED[k][0]], b = m.ev[e * 3 + ED[k][1]];
         auto it = edges.find(key(a, b));
         if (it == edges.end()) { it = edges.emplace(key(a, b), (int32_t)edges.size()).first; }
         e_edge[e * 3 + k] = it->second;
      }
   const int64_t nedges = (int64_t)edges.size();
   const int64_t off_e = m.nv, off_i = off_e + nedges * pm1, ndof = off_i + m.ne * nint;
   elem_dof.resize((size_t)m.ne * nd);
   for (int64_t e = 0; e < m.ne; e++)
   {
      int32_t *g = &elem_dof[(size_t)e * nd];
      for (int v = 0; v < 3; v++) { g[v] = m.ev[e * 3 + v]; }
      int o = 3;
      for (int k = 0; k < 3; k++)
      {
         const int32_t a = m.ev[e * 3 + ED[k][0]], b = m.ev[e * 3 + ED[k][1]];
         const int64_t base = off_e + (int64_t)e_edge[e * 3 + k] * pm1;
         for (int i = 0; i < pm1; i++) { g[o++] = (int32_t)(base + (a < b ? i : pm1 - 1 - i)); }
      }
      for (int i = 0; i < nint; i++) { g[o++] = (int32_t)(off_i + e * nint + i); }
   }
   // boundary segments: 2 vertices + the (p-1) dofs of their edge
   bdr_off.assign(m.nbe + 1, 0);
   bdr_flat.clear();
   for (int64_t b = 0; b < m.nbe; b++)
   {
      const int32_t a = m.bv[b * 2], c = m.bv[b * 2 + 1];
      auto it = edges.find(key(a, c));
      if (it == edges.end()) { return -1; }
      bdr_flat.push_back(a); bdr_flat.push_back(c);
      for (int i = 0; i < pm1; i++) { bdr_flat.push_back((int32_t)(off_e + (int64_t)it->second * pm1 + i)); }
      bdr_off[b + 1] = (int32_t)bdr_flat.size();
   }
   return ndof;
}
