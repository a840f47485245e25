// shim_generic_operator.cpp -- the parts of the MFEM-shaped surface that go beyond the fused form + solver pair:
//   * IterativeSolver::SetOperator(const Operator &) with a user-defined cdm::Operator and a preconditioner Operator
//     (mfem::Solver::SetOperator / SetPreconditioner): the generic Krylov loop through virtual Mult must reproduce the
//     fused device-resident drivers (same iteration count, histories to 1e-10);
//   * two MassIntegrators on one form (diffusion_mms_ale.cpp:1018-1021: Mass(J) and Mass(-div phi));
//   * a.RecoverFEMSolution (linear_convection_diffusion_2D.cpp:377);
//   * BilinearFormIntegrator::AddMultPA / AssembleDiagonalPA on E-vectors.
//   ./shim_generic_operator [n=8] [order=3]        exit code 0 = all checks passed, 1 = mismatch, 3 = runtime failure
#include "cdm_mfem_shim.hpp"
#include <cmath>
#include <cstdio>
#include <cstdlib>

namespace
{
// y = A x through the wrapped form: what a user-written mfem::Operator subclass looks like
class Wrapped : public cdm::Operator
{
public:
   explicit Wrapped(const cdm::ConvectionDiffusionForm &a) : cdm::Operator(a.Height()), a_(a) {}
   void Mult(const cdm::Vector &x, cdm::Vector &y) const override { a_.Mult(x, y); }
private:
   const cdm::ConvectionDiffusionForm &a_;
};
// OperatorJacobiSmoother: y = x ./ diag(A)
class JacobiSmoother : public cdm::Operator
{
public:
   JacobiSmoother(const cdm::Device &dev, const cdm::ConvectionDiffusionForm &a) : cdm::Operator(a.Height()), d_(dev, a.Height()), dinv_(dev, a.Height())
   {
      a.AssembleDiagonal(d_);
      std::vector<double> h = d_.HostCopy();
      for (double &v : h) { v = 1.0 / v; }
      dinv_.SetFromHost(h.data());
   }
   void Mult(const cdm::Vector &x, cdm::Vector &y) const override
   { cdm::check(x.ctx(), cdm_pointwise_mult(x.ctx(), x.Size(), dinv_.Read(), x.Read(), y.ReadWrite()), "cdm_pointwise_mult"); }
private:
   cdm::Vector d_, dinv_;
};
double rel(const std::vector<double> &a, const std::vector<double> &b)
{
   double s = 0.0, t = 0.0;
   for (size_t i = 0; i < a.size(); i++) { s += (a[i] - b[i]) * (a[i] - b[i]); t += b[i] * b[i]; }
   return std::sqrt(s / (t > 0 ? t : 1.0));
}
}  // namespace

int main(int argc, char **argv)
{
   const int n = argc > 1 ? std::atoi(argv[1]) : 8;
   const int order = argc > 2 ? std::atoi(argv[2]) : 3;
   int bad = 0;
   try
   {
      cdm::Device device(0);
      cdm::Mesh mesh = cdm::Mesh::MakeCartesian3D(device, n, n + 1, n, 0.1);
      cdm::H1Space fes(mesh, order);
      std::vector<int> ess_bdr(6, 1);
      std::vector<int32_t> ess;
      fes.GetEssentialTrueDofs(ess_bdr, ess);
      const int64_t N = fes.GetTrueVSize();
      const std::vector<double> vel = {1.0, -2.0, 0.5};

      // ---- two mass integrators = one with the summed coefficient
      cdm::ConvectionDiffusionForm a(fes), a_ref(fes);
      a.AddDiffusionIntegrator(0.1); a.AddConvectionIntegrator(vel);
      a.AddMassIntegrator(0.4);
      a.AddMassIntegrator(std::vector<double>((size_t)fes.GetNE() * fes.GetNQ(), 0.6));
      a.SetEssentialTrueDofs(ess); a.Assemble();
      a_ref.AddDiffusionIntegrator(0.1); a_ref.AddConvectionIntegrator(vel); a_ref.AddMassIntegrator(1.0);
      a_ref.SetEssentialTrueDofs(ess); a_ref.Assemble();
      cdm::Vector x(device, N), y(device, N), y2(device, N), b(device, N);
      std::vector<double> xh(N);
      for (int64_t i = 0; i < N; i++) { xh[i] = std::sin(1.0 + 0.37 * i); }
      x.SetFromHost(xh.data());
      a.Mult(x, y); a_ref.Mult(x, y2);
      double e = rel(y.HostCopy(), y2.HostCopy());
      std::printf("two mass integrators vs their sum: %.2e\n", e);
      bad |= !(e < 1e-13);

      // ---- generic SetOperator(const Operator &) + SetPreconditioner vs the fused drivers
      for (int i = 0; i < (int)N; i++) { xh[i] = std::cos(0.5 + 0.11 * i); }
      b.SetFromHost(xh.data());
      cdm::Vector g(device, N);
      a_ref.FormLinearSystem(g, b);                               // homogeneous boundary data: b[ess] = 0
      Wrapped W(a_ref);
      JacobiSmoother J(device, a_ref);
      for (int variant : {CDM_GMRES_PETSC, CDM_GMRES_MFEM})
      {
         cdm::GMRESSolver fused(variant), generic(variant);
         fused.SetOperator(a_ref); fused.SetJacobi(true);
         generic.SetOperator(W); generic.SetPreconditioner(J);
         cdm::Vector X1(device, N), X2(device, N);
         fused.Mult(b, X1); generic.Mult(b, X2);
         const std::vector<double> &h1 = fused.ResidualHistory(), &h2 = generic.ResidualHistory();
         double he = 0.0;
         for (size_t i = 0; i < std::min(h1.size(), h2.size()); i++) { he = std::max(he, std::fabs(h1[i] - h2[i]) / h1[0]); }
         e = rel(X2.HostCopy(), X1.HostCopy());
         std::printf("GMRES variant %d: fused %d its, generic Operator %d its, history diff %.2e, solution diff %.2e\n", variant,
                     fused.GetNumIterations(), generic.GetNumIterations(), he, e);
         bad |= !(fused.GetConverged() && generic.GetConverged() && std::abs(fused.GetNumIterations() - generic.GetNumIterations()) <= 1 &&
                  he < 1e-10 && e < 1e-9);
      }
      {
         cdm::ConvectionDiffusionForm k(fes);
         k.AddDiffusionIntegrator(1.0); k.SetEssentialTrueDofs(ess); k.Assemble();
         Wrapped Wk(k);
         cdm::CGSolver fused, generic;
         fused.SetOperator(k); generic.SetOperator(Wk);
         cdm::Vector X1(device, N), X2(device, N);
         fused.Mult(b, X1); generic.Mult(b, X2);
         e = rel(X2.HostCopy(), X1.HostCopy());
         std::printf("CG: fused %d its, generic Operator %d its, solution diff %.2e\n", fused.GetNumIterations(), generic.GetNumIterations(), e);
         bad |= !(fused.GetConverged() && generic.GetConverged() && std::abs(fused.GetNumIterations() - generic.GetNumIterations()) <= 1 && e < 1e-9);
      }

      // ---- RecoverFEMSolution: u_L = P X (single rank: the identity on the true dofs)
      cdm::Vector u(device, a_ref.LocalSize());
      a_ref.RecoverFEMSolution(x, b, u);
      e = rel(u.HostCopy(), x.HostCopy());
      std::printf("RecoverFEMSolution: %.2e\n", e);
      bad |= !(e == 0.0);

      // ---- integrator level: G^T AddMultPA(G x) == BilinearForm::Mult(x); diag likewise
      cdm::ConvectionDiffusionForm au(fes);
      au.AddDiffusionIntegrator(0.1); au.AddConvectionIntegrator(vel); au.AddMassIntegrator(1.0); au.Assemble();
      int nd = 1; for (int d = 0; d < 3; d++) { nd *= order + 1; }
      const int64_t NE = fes.GetNE() * (int64_t)nd;
      cdm::Vector xE(device, NE), yE(device, NE), yL(device, N);
      cdm::check(device.ctx(), cdm_restriction_mult(fes.handle(), x.Read(), xE.ReadWrite()), "cdm_restriction_mult");
      au.AddMultPA(xE, yE);
      cdm::check(device.ctx(), cdm_restriction_mult_transpose(fes.handle(), yE.Read(), yL.ReadWrite()), "cdm_restriction_mult_transpose");
      au.MultUnconstrained(x, y);
      e = rel(yL.HostCopy(), y.HostCopy());
      std::printf("AddMultPA through ElementRestriction vs BilinearForm::Mult: %.2e\n", e);
      bad |= !(e < 1e-12);
      cdm::Vector dE(device, NE), dL(device, N), dref(device, N);
      au.AssembleDiagonalPA(dE);
      cdm::check(device.ctx(), cdm_restriction_mult_transpose(fes.handle(), dE.Read(), dL.ReadWrite()), "cdm_restriction_mult_transpose");
      au.AssembleDiagonal(dref);
      e = rel(dL.HostCopy(), dref.HostCopy());
      std::printf("AssembleDiagonalPA through ElementRestriction vs AssembleDiagonal: %.2e\n", e);
      bad |= !(e < 1e-13);
   }
   catch (const std::exception &ex)
   {
      std::fprintf(stderr, "Error: %s\n", ex.what());
      return 3;
   }
   std::printf(bad ? "FAILED\n" : "all shim checks passed\n");
   return bad ? 1 : 0;
}
