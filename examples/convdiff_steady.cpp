// convdiff_steady.cpp -- the reference's steady driver (linear_convection_diffusion_2D.cpp:238-446)
// re-hosted on the B200 library through the MFEM-shaped shim: same sequence of calls
// (mesh -> H1 space -> essential dofs -> Diffusion+Convection+Mass form -> linear form ->
// boundary projection -> FormLinearSystem -> GMRES+Jacobi -> L2 error vs the manufactured solution),
// every step on the device,
// on an inline-hex (or inline-quad) Cartesian mesh instead of the Gmsh triangle mesh.
//
//   ./convdiff_steady [dim=3] [n=16] [order=3] [legacy]      (legacy: assembled CSR + SpMV, the app's own algorithm)
// exit codes as in the reference: 0 ok, 3 runtime failure (:435-442).
#include "cdm_mfem_shim.hpp"
#include <cmath>
#include <cstdio>
#include <cstdlib>

int main(int argc, char **argv)
{
   const int dim = argc > 1 ? std::atoi(argv[1]) : 3;
   const int n = argc > 2 ? std::atoi(argv[2]) : 16;
   const int order = argc > 3 ? std::atoi(argv[3]) : 3;
   const bool legacy = argc > 4 && std::string(argv[4]) == "legacy";
   const double kappa = 0.1, s = 1.0, c[3] = {1.0, -2.0, 0.5};     // Input/input_2d.yaml:7-10
   const int mn = 3;                                                 // mode_n = mode_m = 3 (:13-14)
   try
   {
      cdm::Device device(0);
      cdm::Mesh mesh = dim == 3 ? cdm::Mesh::MakeCartesian3D(device, n, n, n, 0.1) : cdm::Mesh::MakeCartesian2D(device, n, n, 0.1);
      cdm::H1Space fespace(mesh, order);
      std::printf("Global true dofs: %lld\n", (long long)fespace.GetTrueVSize());
      std::vector<int> ess_bdr(dim == 3 ? 6 : 4, 1);                 // all attributes essential (:319-320)
      std::vector<int32_t> ess_tdof_list;
      fespace.GetEssentialTrueDofs(ess_bdr, ess_tdof_list);

      auto exact = [&](const double *x)
      {
         double v = std::sin(mn * M_PI * x[0]) * std::sin(mn * M_PI * x[1]);
         return dim == 3 ? v * std::sin(mn * M_PI * x[2]) : v;
      };
      auto forcing = [&](const double *x)                            // -kappa lap u + c.grad u + s u
      {
         const double a = mn * M_PI;
         double S[3], C[3];
         for (int d = 0; d < dim; d++) { S[d] = std::sin(a * x[d]); C[d] = std::cos(a * x[d]); }
         if (dim == 2) { return kappa * 2 * a * a * S[0] * S[1] + c[0] * a * C[0] * S[1] + c[1] * a * S[0] * C[1] + s * S[0] * S[1]; }
         return kappa * 3 * a * a * S[0] * S[1] * S[2] + c[0] * a * C[0] * S[1] * S[2] + c[1] * a * S[0] * C[1] * S[2]
                + c[2] * a * S[0] * S[1] * C[2] + s * S[0] * S[1] * S[2];
      };

      // ParBilinearForm a: Diffusion + Convection + Mass (:335-339)
      cdm::ConvectionDiffusionForm a(fespace);
      a.AddDiffusionIntegrator(kappa);
      a.AddConvectionIntegrator(std::vector<double>(c, c + dim));
      a.AddMassIntegrator(s);
      a.SetEssentialTrueDofs(ess_tdof_list);
      if (legacy) { a.SetAssemblyLevel(cdm::ConvectionDiffusionForm::AssemblyLevel::LEGACY); }   // default: PARTIAL
      a.Assemble();

      // ParLinearForm b: DomainLFIntegrator(f) (:341-343)
      const int64_t N = fespace.GetTrueVSize();
      cdm::Vector b(device, N), u(device, N), X(device, N);
      fespace.AssembleDomainLF(forcing, b);
      // u = 0; u.ProjectBdrCoefficient(exact, ess_bdr) (:345-347)
      u = 0.0;
      fespace.ProjectBdrCoefficient(exact, ess_tdof_list, u);
      a.FormLinearSystem(u, b);                                       // (:351)

      // PetscLinearSolver with Input/petsc.opts: gmres, rtol 1e-10, atol 1e-12, max_it 500, jacobi (:368-374)
      cdm::GMRESSolver solver(CDM_GMRES_PETSC);
      solver.SetRelTol(1e-10); solver.SetAbsTol(1e-12); solver.SetMaxIter(500); solver.SetJacobi(true);
      solver.SetPrintLevel(0);
      solver.SetOperator(a);
      solver.Mult(b, X);
      if (!solver.GetConverged())
      {
         throw std::runtime_error("solver did not converge. Iterations=" + std::to_string(solver.GetNumIterations()) +
                                  ", residual=" + std::to_string(solver.GetFinalNorm()));
      }
      // u.ComputeL2Error(exact, irs) / ComputeGlobalLpNorm with rules of order max(2, 2p+3) (:383-392)
      const double abs_l2 = fespace.ComputeL2Error(X, exact);
      const double exact_l2 = fespace.ComputeGlobalL2Norm(exact);
      const double rel_l2 = (exact_l2 > 1.0e-14) ? abs_l2 / exact_l2 : 0.0;
      std::printf("GMRES iterations: %d, final residual %.3e, solve %.3f ms\n", solver.GetNumIterations(),
                  solver.GetFinalNorm(), solver.GetSolveSeconds() * 1e3);
      std::printf("L2 error: abs %.6e  rel %.6e\n", abs_l2, rel_l2);
   }
   catch (const std::exception &e)
   {
      std::fprintf(stderr, "Error: %s\n", e.what());
      return 3;
   }
   return 0;
}
