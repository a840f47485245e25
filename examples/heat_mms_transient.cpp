// heat_mms_transient.cpp -- the reference's transient heat-equation driver (diffusion_mms.cpp:286-470)
// re-hosted on the B200 library through the MFEM-shaped shim.  Backward Euler for
//     du/dt - alpha Lap(u) = f,   u = sin(t) cos(q),  q = 2 |x - 1/2|^2     (diffusion_mms.cpp:136-178)
// with every per-step operation on the device:
//     rhs = M u^n                       mass_form.Mult(u, rhs_local)               (:430)
//     rhs += dt (f^{n+1}, v)            ParLinearForm + DomainLFIntegrator         (:433-437)
//     u|_bdr = u_exact(t^{n+1})         u.ProjectBdrCoefficient(exact, ess_bdr)    (:441)
//     (M + alpha dt K) u^{n+1} = rhs    FormLinearSystem + GMRES/Jacobi            (:444-456)
//     L2 error                          u.ComputeL2Error(exact, irs)               (:383-392 of the steady driver)
// Defaults follow Input/input_diffusion_mms.yaml:11-13 (alpha 0.1, dt 0.05, t_final 2).
//
//   ./heat_mms_transient [dim=2] [n=16] [order=2] [dt=0.05] [t_final=2.0]
// exit codes as in the reference: 0 ok, 3 runtime failure.
#include "cdm_mfem_shim.hpp"
#include <cmath>
#include <cstdio>
#include <cstdlib>

int main(int argc, char **argv)
{
   const int dim = argc > 1 ? std::atoi(argv[1]) : 2;
   const int n = argc > 2 ? std::atoi(argv[2]) : 16;
   const int order = argc > 3 ? std::atoi(argv[3]) : 2;
   const double dt = argc > 4 ? std::atof(argv[4]) : 0.05;
   const double t_final = argc > 5 ? std::atof(argv[5]) : 2.0;
   const double alpha = 0.1;
   try
   {
      cdm::Device device(0);
      cdm::Mesh mesh = dim == 3 ? cdm::Mesh::MakeCartesian3D(device, n, n, n, 0.0) : cdm::Mesh::MakeCartesian2D(device, n, n, 0.0);
      cdm::H1Space fespace(mesh, order);
      std::vector<int> ess_bdr(dim == 3 ? 6 : 4, 1);
      std::vector<int32_t> ess_tdof_list;
      fespace.GetEssentialTrueDofs(ess_bdr, ess_tdof_list);
      const int64_t N = fespace.GetTrueVSize();
      std::printf("Global true dofs: %lld, essential: %zu\n", (long long)N, ess_tdof_list.size());

      double t = 0.0;                                                    // Coefficient::SetTime
      auto r2 = [&](const double *x) { double s = 0.0; for (int d = 0; d < dim; d++) { s += (x[d] - 0.5) * (x[d] - 0.5); } return s; };
      auto exact = [&](const double *x) { return std::sin(t) * std::cos(2.0 * r2(x)); };
      auto forcing = [&](const double *x)                                // du/dt - alpha Lap(u)
      {
         const double rr = r2(x), q = 2.0 * rr;
         const double lap = std::sin(t) * (-16.0 * rr * std::cos(q) - 4.0 * dim * std::sin(q));
         return std::cos(t) * std::cos(q) - alpha * lap;
      };

      cdm::ConvectionDiffusionForm mass_form(fespace);                   // M
      mass_form.AddMassIntegrator(1.0);
      mass_form.Assemble();
      cdm::ConvectionDiffusionForm lhs_form(fespace);                    // M + alpha dt K
      lhs_form.AddMassIntegrator(1.0);
      lhs_form.AddDiffusionIntegrator(alpha * dt);
      lhs_form.SetEssentialTrueDofs(ess_tdof_list);
      lhs_form.Assemble();

      cdm::Vector u(device, N), rhs(device, N), X(device, N);
      u = 0.0;                                                           // u(x, 0) = 0
      cdm::GMRESSolver solver(CDM_GMRES_PETSC);                          // Input/petsc.opts
      solver.SetRelTol(1e-10); solver.SetAbsTol(1e-12); solver.SetMaxIter(500); solver.SetJacobi(true);
      solver.SetOperator(lhs_form);

      const int nsteps = (int)std::ceil(t_final / dt - 1.0e-12);
      std::printf("Time steps: %d, dt=%g, t_final=%g\n", nsteps, dt, nsteps * dt);
      int total_its = 0;
      double solve_s = 0.0, max_err = 0.0;
      for (int step = 1; step <= nsteps; step++)
      {
         t = step * dt;
         mass_form.MultUnconstrained(u, rhs);                            // rhs = M u^n
         fespace.AssembleDomainLF(forcing, rhs, dt, /*accumulate=*/true);// rhs += dt (f^{n+1}, v)
         fespace.ProjectBdrCoefficient(exact, ess_tdof_list, u);         // Dirichlet data at t^{n+1}
         lhs_form.FormLinearSystem(u, rhs);
         solver.Mult(rhs, X);
         if (!solver.GetConverged())
         {
            throw std::runtime_error("solver did not converge at step " + std::to_string(step) + ". Iterations=" +
                                     std::to_string(solver.GetNumIterations()) + ", residual=" + std::to_string(solver.GetFinalNorm()));
         }
         u.Assign(X);                                                    // RecoverFEMSolution
         total_its += solver.GetNumIterations();
         solve_s += solver.GetSolveSeconds();
         const double err = fespace.ComputeL2Error(u, exact);
         max_err = std::max(max_err, err);
         if (step % 10 == 0 || step == nsteps) { std::printf("step %3d t=%.3f its=%3d L2_error=%.6e\n", step, t, solver.GetNumIterations(), err); }
      }
      t = nsteps * dt;
      std::printf("Final L2 error at t=%g: %.6e  (max over steps %.6e)\n", t, fespace.ComputeL2Error(u, exact), max_err);
      std::printf("GMRES iterations total %d, solve time %.3f ms\n", total_its, solve_s * 1e3);
   }
   catch (const std::exception &e)
   {
      std::fprintf(stderr, "Error: %s\n", e.what());
      return 3;
   }
   return 0;
}
