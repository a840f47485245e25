// host_vtu.cpp -- ParaView output of fields on an H1 space (host side, after the solve):
//   ParaViewDataCollection dc(name, pmesh); dc.SetPrefixPath(path); dc.SetLevelsOfDetail(order);
//   dc.RegisterField("u", &u); dc.SetCycle(0); dc.SetTime(0.0); dc.Save();     linear_convection_diffusion_2D.cpp:421-433
// Layout on disk follows MFEM's collection: <prefix>/<name>/<name>.pvd, <prefix>/<name>/Cycle000000/data.pvtu and
// proc000000.vtu.  Every element is written with its own copy of its nodes (as MFEM does) and refined into order^dim
// linear sub-cells on its node lattice (LevelsOfDetail = order), VTK ASCII.
#include "cdm_internal.hpp"
#include <cerrno>
#include <cstdio>
#include <cstring>
#include <string>
#include <sys/stat.h>

namespace
{
bool make_dirs(const std::string &path)
{
   std::string cur;
   for (size_t i = 0; i <= path.size(); i++)
   {
      if (i == path.size() || path[i] == '/')
      {
         if (!cur.empty() && mkdir(cur.c_str(), 0777) != 0 && errno != EEXIST) { return false; }
      }
      if (i < path.size()) { cur.push_back(path[i]); }
   }
   return true;
}

// native dof index of lattice point (i, j), i + j <= p, of the order-p triangle (host_simplex.cpp node order)
int tri_lattice(int p, int i, int j)
{
   const int k = p - i - j;
   if (j == 0 && k == p) { return 0; }
   if (i == p) { return 1; }
   if (j == p) { return 2; }
   if (j == 0) { return 3 + (i - 1); }                       // edge (0,1)
   if (k == 0) { return 3 + (p - 1) + (j - 1); }             // edge (1,2): points (p - t, t)
   if (i == 0) { return 3 + 2 * (p - 1) + (p - j - 1); }     // edge (2,0): points (0, p - t)
   int o = 3 + 3 * (p - 1);
   for (int jj = 1; jj < j; jj++) { o += p - 1 - jj; }
   return o + (i - 1);
}
}  // namespace

extern "C" int cdm_write_paraview(const cdm_space *sp, const char *prefix_path, const char *collection, int cycle, double time,
                                  int nfields, const char *const *names, const double *const *fields_host)
{
   if (!sp || !collection || nfields < 0 || (nfields > 0 && (!names || !fields_host))) { return CDM_EINVAL; }
   cdm_ctx *ctx = sp->ctx;
   const std::string base = std::string(prefix_path && prefix_path[0] ? prefix_path : ".") + "/" + collection;
   char cyc[64];
   std::snprintf(cyc, sizeof(cyc), "Cycle%06d", cycle);
   const std::string dir = base + "/" + cyc;
   if (!make_dirs(dir)) { return cdm_fail(ctx, CDM_EINVAL, "cdm_write_paraview: cannot create " + dir); }
   const int dim = sp->dim, p = sp->p, nd = sp->nd, d1d = sp->d1d;
   std::vector<double> X((size_t)sp->ndof * dim);
   cdm_space_dof_coords(sp, X.data());
   const int rank = ctx ? ctx->rank : 0;
   char procname[64];
   std::snprintf(procname, sizeof(procname), "proc%06d.vtu", rank);
   FILE *f = std::fopen((dir + "/" + procname).c_str(), "w");
   if (!f) { return cdm_fail(ctx, CDM_EINVAL, "cdm_write_paraview: cannot write " + dir); }
   const bool tri = sp->geom == 1;
   const int64_t sub = tri ? (int64_t)p * p : (dim == 2 ? (int64_t)p * p : (int64_t)p * p * p);
   const int64_t npts = sp->ne * nd, ncells = sp->ne * sub;
   const int vpc = tri ? 3 : (dim == 2 ? 4 : 8), vtk_type = tri ? 5 : (dim == 2 ? 9 : 12);
   std::fprintf(f, "<?xml version=\"1.0\"?>\n<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n<UnstructuredGrid>\n");
   std::fprintf(f, "<Piece NumberOfPoints=\"%lld\" NumberOfCells=\"%lld\">\n<Points>\n<DataArray type=\"Float64\" NumberOfComponents=\"3\" format=\"ascii\">\n",
                (long long)npts, (long long)ncells);
   for (int64_t e = 0; e < sp->ne; e++)
      for (int l = 0; l < nd; l++)
      {
         const double *x = &X[(size_t)sp->gather[(size_t)e * nd + l] * dim];
         std::fprintf(f, "%.16g %.16g %.16g\n", x[0], x[1], dim == 3 ? x[2] : 0.0);
      }
   std::fprintf(f, "</DataArray>\n</Points>\n<Cells>\n<DataArray type=\"Int64\" Name=\"connectivity\" format=\"ascii\">\n");
   for (int64_t e = 0; e < sp->ne; e++)
   {
      const long long o = (long long)e * nd;
      if (tri)
      {
         for (int j = 0; j < p; j++)
            for (int i = 0; i + j < p; i++)
            {
               std::fprintf(f, "%lld %lld %lld\n", o + tri_lattice(p, i, j), o + tri_lattice(p, i + 1, j), o + tri_lattice(p, i, j + 1));
               if (i + j + 1 < p)
                  std::fprintf(f, "%lld %lld %lld\n", o + tri_lattice(p, i + 1, j), o + tri_lattice(p, i + 1, j + 1), o + tri_lattice(p, i, j + 1));
            }
      }
      else if (dim == 2)
      {
         for (int j = 0; j < p; j++)
            for (int i = 0; i < p; i++)
               std::fprintf(f, "%lld %lld %lld %lld\n", o + i + d1d * j, o + i + 1 + d1d * j, o + i + 1 + d1d * (j + 1), o + i + d1d * (j + 1));
      }
      else
      {
         auto L = [&](int i, int j, int k) { return o + i + d1d * (j + d1d * k); };
         for (int k = 0; k < p; k++)
            for (int j = 0; j < p; j++)
               for (int i = 0; i < p; i++)
                  std::fprintf(f, "%lld %lld %lld %lld %lld %lld %lld %lld\n", L(i, j, k), L(i + 1, j, k), L(i + 1, j + 1, k), L(i, j + 1, k),
                               L(i, j, k + 1), L(i + 1, j, k + 1), L(i + 1, j + 1, k + 1), L(i, j + 1, k + 1));
      }
   }
   std::fprintf(f, "</DataArray>\n<DataArray type=\"Int64\" Name=\"offsets\" format=\"ascii\">\n");
   for (int64_t cidx = 1; cidx <= ncells; cidx++) { std::fprintf(f, "%lld\n", (long long)cidx * vpc); }
   std::fprintf(f, "</DataArray>\n<DataArray type=\"UInt8\" Name=\"types\" format=\"ascii\">\n");
   for (int64_t cidx = 0; cidx < ncells; cidx++) { std::fprintf(f, "%d\n", vtk_type); }
   std::fprintf(f, "</DataArray>\n</Cells>\n<PointData>\n");
   for (int k = 0; k < nfields; k++)
   {
      std::fprintf(f, "<DataArray type=\"Float64\" Name=\"%s\" NumberOfComponents=\"1\" format=\"ascii\">\n", names[k]);
      for (int64_t e = 0; e < sp->ne; e++)
         for (int l = 0; l < nd; l++) { std::fprintf(f, "%.16g\n", fields_host[k][sp->gather[(size_t)e * nd + l]]); }
      std::fprintf(f, "</DataArray>\n");
   }
   std::fprintf(f, "</PointData>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n");
   std::fclose(f);
   if (rank == 0)
   {
      const int nr = ctx ? ctx->nranks : 1;
      FILE *g = std::fopen((dir + "/data.pvtu").c_str(), "w");
      if (!g) { return cdm_fail(ctx, CDM_EINVAL, "cdm_write_paraview: cannot write data.pvtu"); }
      std::fprintf(g, "<?xml version=\"1.0\"?>\n<VTKFile type=\"PUnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n<PUnstructuredGrid GhostLevel=\"0\">\n");
      std::fprintf(g, "<PPoints>\n<PDataArray type=\"Float64\" Name=\"Points\" NumberOfComponents=\"3\"/>\n</PPoints>\n<PPointData>\n");
      for (int k = 0; k < nfields; k++) { std::fprintf(g, "<PDataArray type=\"Float64\" Name=\"%s\" NumberOfComponents=\"1\" format=\"ascii\"/>\n", names[k]); }
      std::fprintf(g, "</PPointData>\n");
      for (int r = 0; r < nr; r++) { std::fprintf(g, "<Piece Source=\"proc%06d.vtu\"/>\n", r); }
      std::fprintf(g, "</PUnstructuredGrid>\n</VTKFile>\n");
      std::fclose(g);
      FILE *h = std::fopen((base + "/" + collection + ".pvd").c_str(), "w");
      if (!h) { return cdm_fail(ctx, CDM_EINVAL, "cdm_write_paraview: cannot write the .pvd file"); }
      std::fprintf(h, "<?xml version=\"1.0\"?>\n<VTKFile type=\"Collection\" version=\"0.1\">\n<Collection>\n");
      std::fprintf(h, "<DataSet timestep=\"%.16g\" group=\"\" part=\"0\" file=\"%s/data.pvtu\"/>\n", time, cyc);
      std::fprintf(h, "</Collection>\n</VTKFile>\n");
      std::fclose(h);
   }
   return CDM_OK;
}
