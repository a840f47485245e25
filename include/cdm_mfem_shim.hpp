// cdm_mfem_shim.hpp -- header-only C++ layer over the C ABI (cdm_b200.h) with the MFEM
// surface that myapps/convection_diffusion calls on this path, so the app's main() keeps
// its shape (same names, argument meaning and error behaviour):
//
//   mfem::Vector                      -> cdm::Vector          (device-resident, host mirror on demand)
//   mfem::Operator::Mult              -> cdm::Operator::Mult  (linear_convection_diffusion_1D.cpp:544)
//   ParBilinearForm + integrators     -> cdm::ConvectionDiffusionForm
//        AddDomainIntegrator / Assemble / FormLinearSystem / RecoverFEMSolution
//        (linear_convection_diffusion_2D.cpp:335-351, :377)
//   PetscLinearSolver / GMRESSolver   -> cdm::GMRESSolver     (:368-374, Input/petsc.opts:2-6)
//   mfem::CGSolver                    -> cdm::CGSolver        (mesh_recession_handler.cpp:270-276)
//
// Errors: every non-zero C-ABI return becomes std::runtime_error, matching the app's
// try/catch -> exit code 3 (linear_convection_diffusion_2D.cpp:435-442).  A solve that does
// not converge is NOT an exception: query GetConverged() like the app does (:371).
#pragma once
#include "cdm_b200.h"
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace cdm
{
inline void check(cdm_ctx *ctx, int rc, const char *what)
{
   if (rc != CDM_OK)
   {
      throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) + "): " +
                               (ctx ? cdm_last_error(ctx) : "no context"));
   }
}

// Device("cuda") analogue (linear_convection_diffusion_2D.cpp:287 uses Device("cpu"))
class Device
{
public:
   explicit Device(int device = 0, void *stream = nullptr)
   {
      const int rc = cdm_init(device, stream, &ctx_);
      if (rc != CDM_OK) { throw std::runtime_error("cdm_init failed: no usable CUDA device (no CPU fallback)"); }
   }
   ~Device() { cdm_finalize(ctx_); }
   Device(const Device &) = delete;
   Device &operator=(const Device &) = delete;
   cdm_ctx *ctx() const { return ctx_; }
   void Sync() const { check(ctx_, cdm_sync(ctx_), "cdm_sync"); }
private:
   cdm_ctx *ctx_ = nullptr;
};

class Vector
{
public:
   Vector() = default;
   Vector(const Device &dev, int64_t n) { SetSize(dev, n); }
   ~Vector() { Destroy(); }
   Vector(const Vector &) = delete;
   Vector &operator=(const Vector &) = delete;
   void SetSize(const Device &dev, int64_t n) { SetSizeOn(dev.ctx(), n); }
   void SetSizeOn(cdm_ctx *ctx, int64_t n)
   {
      Destroy();
      ctx_ = ctx; n_ = n;
      check(ctx_, cdm_vec_alloc(ctx_, n, &d_), "cdm_vec_alloc");
      check(ctx_, cdm_vec_set(ctx_, n, 0.0, d_), "cdm_vec_set");
   }
   int64_t Size() const { return n_; }
   double *Read() const { return d_; }             // device pointer (mfem::Vector::Read with a device backend)
   double *ReadWrite() { return d_; }
   Vector &operator=(double v) { check(ctx_, cdm_vec_set(ctx_, n_, v, d_), "cdm_vec_set"); return *this; }
   void SetFromHost(const double *h) { check(ctx_, cdm_vec_upload(ctx_, n_, h, d_), "cdm_vec_upload"); }
   void GetToHost(double *h) const { check(ctx_, cdm_vec_download(ctx_, n_, d_, h), "cdm_vec_download"); }
   std::vector<double> HostCopy() const { std::vector<double> h(n_); GetToHost(h.data()); return h; }
   void Add(double a, const Vector &x) { check(ctx_, cdm_axpy(ctx_, n_, a, x.d_, d_), "cdm_axpy"); }   // *this += a x
   void Assign(const Vector &x) { *this = 0.0; Add(1.0, x); }                                          // *this = x (device copy)
   void Scale(double a) { check(ctx_, cdm_add(ctx_, n_, d_, a - 1.0, d_, d_), "cdm_add"); }              // *this *= a  (x + (a-1) x)
   void CopyFromDevice(const double *src) { *this = 0.0; check(ctx_, cdm_axpy(ctx_, n_, 1.0, src, d_), "cdm_axpy"); }
   double operator*(const Vector &y) const { double r; check(ctx_, cdm_dot(ctx_, n_, d_, y.d_, &r), "cdm_dot"); return r; }
   double Norml2() const { double r; check(ctx_, cdm_norm2(ctx_, n_, d_, &r), "cdm_norm2"); return r; }
   cdm_ctx *ctx() const { return ctx_; }
private:
   void Destroy() { if (d_) { cdm_vec_free(ctx_, d_); d_ = nullptr; } }
   cdm_ctx *ctx_ = nullptr;
   double *d_ = nullptr;
   int64_t n_ = 0;
};

// mfem::Operator
class Operator
{
public:
   explicit Operator(int64_t s = 0) : height(s), width(s) {}
   virtual ~Operator() = default;
   int64_t Height() const { return height; }
   int64_t Width() const { return width; }
   virtual void Mult(const Vector &x, Vector &y) const = 0;
protected:
   int64_t height, width;
};

class Mesh
{
public:
   static Mesh MakeCartesian3D(const Device &dev, int64_t nx, int64_t ny, int64_t nz, double perturb = 0.0)
   {
      Mesh m; m.ctx_ = dev.ctx();
      const int64_t n[3] = {nx, ny, nz};
      check(m.ctx_, cdm_mesh_cartesian(m.ctx_, 3, n, nullptr, perturb, &m.h_), "cdm_mesh_cartesian");
      return m;
   }
   static Mesh MakeCartesian2D(const Device &dev, int64_t nx, int64_t ny, double perturb = 0.0)
   {
      Mesh m; m.ctx_ = dev.ctx();
      const int64_t n[3] = {nx, ny, 0};
      check(m.ctx_, cdm_mesh_cartesian(m.ctx_, 2, n, nullptr, perturb, &m.h_), "cdm_mesh_cartesian");
      return m;
   }
   Mesh(Mesh &&o) noexcept : ctx_(o.ctx_), h_(o.h_) { o.h_ = nullptr; }
   ~Mesh() { if (h_) { cdm_mesh_destroy(h_); } }
   int Dimension() const { int d; cdm_mesh_sizes(h_, &d, nullptr, nullptr, nullptr); return d; }
   // mfem::Mesh::UniformRefinement() (linear_convection_diffusion_2D.cpp:295-298)
   void UniformRefinement()
   {
      cdm_mesh *r = nullptr;
      check(ctx_, cdm_mesh_uniform_refine(ctx_, h_, &r), "cdm_mesh_uniform_refine");
      cdm_mesh_destroy(h_);
      h_ = r;
   }
   cdm_mesh *handle() const { return h_; }
   cdm_ctx *ctx() const { return ctx_; }
private:
   Mesh() = default;
   cdm_ctx *ctx_ = nullptr;
   cdm_mesh *h_ = nullptr;
};

// H1_FECollection(order, dim) + FiniteElementSpace
class H1Space
{
public:
   H1Space(const Mesh &mesh, int order) : ctx_(mesh.ctx())
   {
      check(ctx_, cdm_space_create_h1(ctx_, mesh.handle(), order, &h_), "cdm_space_create_h1");
      cdm_space_sizes(h_, &dim_, &order_, &ne_, &ndof_, &d1d_, &q1d_, &ntrue_);
   }
   ~H1Space() { cdm_space_destroy(h_); }
   H1Space(const H1Space &) = delete;
   int64_t GetTrueVSize() const { return ntrue_; }
   int64_t GetNE() const { return ne_; }
   int GetNQ() const { int nq = 1; for (int d = 0; d < dim_; d++) { nq *= q1d_; } return nq; }
   int Dimension() const { return dim_; }
   // fespace.GetEssentialTrueDofs(ess_bdr, ess_tdof_list)
   void GetEssentialTrueDofs(const std::vector<int> &ess_bdr, std::vector<int32_t> &list) const
   {
      std::vector<int32_t> marker(ess_bdr.begin(), ess_bdr.end());
      int64_t n = 0;
      check(ctx_, cdm_space_essential_dofs(h_, marker.data(), (int)marker.size(), nullptr, &n), "cdm_space_essential_dofs");
      list.resize(n);
      check(ctx_, cdm_space_essential_dofs(h_, marker.data(), (int)marker.size(), list.data(), &n), "cdm_space_essential_dofs");
   }
   std::vector<double> DofCoordinates() const
   {
      std::vector<double> x((size_t)ndof_ * dim_);
      cdm_space_dof_coords(h_, x.data());
      return x;
   }
   std::vector<double> QuadraturePointCoordinates() const
   {
      std::vector<double> x((size_t)ne_ * GetNQ() * dim_);
      cdm_space_qpt_coords(h_, x.data());
      return x;
   }
   int GetOrder() const { return order_; }
   // physical points of the rule IntRules.Get(geom, order) in every element (what Coefficient::Eval sees)
   const std::vector<double> &RulePoints(int rule_order) const
   {
      const int q1d = cdm_rule_points(rule_order);
      auto it = rule_pts_.find(q1d);
      if (it != rule_pts_.end()) { return it->second; }
      size_t nq = 1;
      for (int d = 0; d < dim_; d++) { nq *= (size_t)q1d; }
      std::vector<double> x((size_t)ne_ * nq * dim_);
      check(ctx_, cdm_space_rule_coords(h_, q1d, x.data()), "cdm_space_rule_coords");
      return rule_pts_.emplace(q1d, std::move(x)).first->second;
   }
   // FunctionCoefficient evaluated at the points of a rule
   std::vector<double> EvalAtRule(int rule_order, const std::function<double(const double *)> &f) const
   {
      const std::vector<double> &x = RulePoints(rule_order);
      std::vector<double> v(x.size() / dim_);
      for (size_t i = 0; i < v.size(); i++) { v[i] = f(&x[i * dim_]); }
      return v;
   }
   // ParLinearForm b(&fes); b.AddDomainIntegrator(new DomainLFIntegrator(f)); b.Assemble();
   // b = [b +] scale * (f, v)  (linear_convection_diffusion_2D.cpp:341-343, diffusion_mms.cpp:433-437)
   void AssembleDomainLF(const std::function<double(const double *)> &f, Vector &b, double scale = 1.0,
                         bool accumulate = false) const
   {
      const int ord = 2 * order_;                              // DomainLFIntegrator(Q, a = 2, b = 0)
      const std::vector<double> fq = EvalAtRule(ord, f);
      check(ctx_, cdm_domain_lf(h_, cdm_rule_points(ord), fq.data(), scale, accumulate ? 1 : 0, b.ReadWrite()), "cdm_domain_lf");
      check(ctx_, cdm_sync(ctx_), "cdm_sync");                 // fq is a temporary
   }
   // u.ProjectBdrCoefficient(g, ess_bdr) on the dofs of `ess` (linear_convection_diffusion_2D.cpp:347)
   void ProjectBdrCoefficient(const std::function<double(const double *)> &g, const std::vector<int32_t> &ess, Vector &u) const
   {
      if (dof_x_.empty()) { dof_x_ = DofCoordinates(); }
      std::vector<double> vals(ess.size());
      for (size_t i = 0; i < ess.size(); i++) { vals[i] = g(&dof_x_[(size_t)ess[i] * dim_]); }
      check(ctx_, cdm_vec_set_indexed(ctx_, (int64_t)ess.size(), ess.data(), vals.data(), u.ReadWrite()), "cdm_vec_set_indexed");
      check(ctx_, cdm_sync(ctx_), "cdm_sync");
   }
   // u.ComputeL2Error(exact, irs) with irs of order max(2, 2p+3) (linear_convection_diffusion_2D.cpp:383-390)
   double ComputeL2Error(const Vector &u, const std::function<double(const double *)> &exact) const
   {
      const int ord = std::max(2, 2 * order_ + 3);
      const std::vector<double> ex = EvalAtRule(ord, exact);
      double r = 0.0;
      check(ctx_, cdm_l2_error(h_, cdm_rule_points(ord), u.Read(), ex.data(), &r), "cdm_l2_error");
      return r;
   }
   // ComputeGlobalLpNorm(2, exact, mesh, irs) (:391)
   double ComputeGlobalL2Norm(const std::function<double(const double *)> &exact) const
   {
      const int ord = std::max(2, 2 * order_ + 3);
      const std::vector<double> ex = EvalAtRule(ord, exact);
      double r = 0.0;
      check(ctx_, cdm_l2_error(h_, cdm_rule_points(ord), nullptr, ex.data(), &r), "cdm_l2_error");
      return r;
   }
   cdm_space *handle() const { return h_; }
   cdm_ctx *ctx() const { return ctx_; }
private:
   cdm_ctx *ctx_;
   cdm_space *h_ = nullptr;
   mutable std::map<int, std::vector<double>> rule_pts_;
   mutable std::vector<double> dof_x_;
   int dim_ = 0, order_ = 0, d1d_ = 0, q1d_ = 0;
   int64_t ne_ = 0, ndof_ = 0, ntrue_ = 0;
};

// ParBilinearForm with {Diffusion, Convection, Mass}Integrator under AssemblyLevel::PARTIAL
class ConvectionDiffusionForm : public Operator
{
public:
   explicit ConvectionDiffusionForm(H1Space &fes) : Operator(fes.GetTrueVSize()), fes_(fes) {}
   ~ConvectionDiffusionForm() override { cdm_operator_destroy(op_); }
   // a.AddDomainIntegrator(new DiffusionIntegrator(kappa))
   void AddDiffusionIntegrator(double kappa) { kappa_c_ = {kappa}; kap_ = {CDM_COEFF_CONST, 1, kappa_c_.data()}; }
   void AddDiffusionIntegrator(const std::vector<double> &per_qpt, int ncomp = 1)
   { kappa_c_ = per_qpt; kap_ = {CDM_COEFF_QPT, ncomp, kappa_c_.data()}; }
   // a.AddDomainIntegrator(new ConvectionIntegrator(velocity, alpha))
   void AddConvectionIntegrator(const std::vector<double> &velocity, double alpha = 1.0)
   { vel_c_ = velocity; vel_ = {CDM_COEFF_CONST, (int)velocity.size(), vel_c_.data()}; alpha_ = alpha; }
   void AddConvectionIntegratorQpt(const std::vector<double> &per_qpt, double alpha = 1.0)
   { vel_c_ = per_qpt; vel_ = {CDM_COEFF_QPT, fes_.Dimension(), vel_c_.data()}; alpha_ = alpha; }
   // a.AddDomainIntegrator(new MassIntegrator(s)).  May be called several times: the reference's ALE form carries two
   // mass integrators, Mass(J) and Mass(-div phi) (diffusion_mms_ale.cpp:1018-1021); mass integrators are additive in
   // their coefficient, so the coefficients are summed point by point and ONE D_mass stream is stored and read.
   void AddMassIntegrator(double s)
   {
      if (mass_.kind == CDM_COEFF_NONE) { mass_c_ = {s}; mass_ = {CDM_COEFF_CONST, 1, mass_c_.data()}; }
      else { for (double &v : mass_c_) { v += s; } }
   }
   void AddMassIntegrator(const std::vector<double> &per_qpt)
   {
      if (mass_.kind == CDM_COEFF_NONE) { mass_c_ = per_qpt; }
      else if (mass_.kind == CDM_COEFF_CONST) { const double c0 = mass_c_[0]; mass_c_ = per_qpt; for (double &v : mass_c_) { v += c0; } }
      else
      {
         if (per_qpt.size() != mass_c_.size()) { throw std::runtime_error("AddMassIntegrator: per-point arrays of different length"); }
         for (size_t i = 0; i < per_qpt.size(); i++) { mass_c_[i] += per_qpt[i]; }
      }
      mass_ = {CDM_COEFF_QPT, 1, mass_c_.data()};
   }
   void SetEssentialTrueDofs(const std::vector<int32_t> &ess) { ess_ = ess; }
   // a.SetAssemblyLevel(AssemblyLevel::PARTIAL / LEGACY): PARTIAL (default) is the matrix-free operator;
   // LEGACY assembles the sparse matrix the application builds today and applies it as a CSR SpMV
   enum class AssemblyLevel { LEGACY, PARTIAL };
   void SetAssemblyLevel(AssemblyLevel level)
   {
      level_ = level;
      if (op_) { check(fes_.ctx(), cdm_operator_set_option(op_, "assembly", level_ == AssemblyLevel::LEGACY ? 1 : 0), "cdm_operator_set_option"); }
   }
   // a.Assemble(): quadrature data on the device (and the CSR matrix for AssemblyLevel::LEGACY)
   void Assemble()
   {
      if (op_) { cdm_operator_destroy(op_); op_ = nullptr; }
      check(fes_.ctx(), cdm_operator_create(fes_.handle(), &kap_, &vel_, alpha_, &mass_, ess_.data(), (int64_t)ess_.size(), &op_),
            "cdm_operator_create");
      if (level_ == AssemblyLevel::LEGACY) { check(fes_.ctx(), cdm_operator_set_option(op_, "assembly", 1), "cdm_operator_set_option"); }
   }
   // Operator::Mult of the constrained system operator
   void Mult(const Vector &x, Vector &y) const override
   { check(fes_.ctx(), cdm_operator_apply(op_, x.Read(), y.ReadWrite()), "cdm_operator_apply"); }
   // BilinearForm::Mult on L-vectors, no constraints (mass_form.Mult(c, rhs))
   void MultUnconstrained(const Vector &x, Vector &y) const
   { check(fes_.ctx(), cdm_operator_apply_unconstrained(op_, x.Read(), y.ReadWrite()), "cdm_operator_apply_unconstrained"); }
   // a.FormLinearSystem(ess_tdof_list, u, b, A, X, B): X = u, B = b - A u_ess, B[ess] = u[ess]
   void FormLinearSystem(const Vector &u, Vector &b) const
   { check(fes_.ctx(), cdm_eliminate_rhs(op_, u.Read(), b.ReadWrite()), "cdm_eliminate_rhs"); }
   void AssembleDiagonal(Vector &d) const { check(fes_.ctx(), cdm_operator_diag(op_, d.ReadWrite()), "cdm_operator_diag"); }
   // a.RecoverFEMSolution(X, b, u): u_L = P X (linear_convection_diffusion_2D.cpp:377).  u has LocalSize() entries: the
   // true dofs followed by the ghost dofs of this rank, whose values are fetched from their owners.
   int64_t LocalSize() const { return cdm_space_local_size(fes_.handle()); }
   void RecoverFEMSolution(const Vector &X, const Vector & /*b*/, Vector &u) const
   {
      if (u.Size() < LocalSize()) { throw std::runtime_error("RecoverFEMSolution: u must have LocalSize() entries"); }
      check(fes_.ctx(), cdm_prolongate(fes_.handle(), X.Read(), u.ReadWrite()), "cdm_prolongate");
   }
   // BilinearFormIntegrator-level entry points (AssemblyLevel::PARTIAL): E-vectors [ne][nd] in, E-vectors out
   void AddMultPA(const Vector &xE, Vector &yE) const
   { check(fes_.ctx(), cdm_integrator_add_mult_pa(op_, xE.Read(), yE.ReadWrite()), "cdm_integrator_add_mult_pa"); }
   void AssembleDiagonalPA(Vector &diagE) const
   { check(fes_.ctx(), cdm_integrator_assemble_diagonal_pa(op_, diagE.ReadWrite()), "cdm_integrator_assemble_diagonal_pa"); }
   cdm_op *handle() const { return op_; }
   cdm_ctx *ctx() const { return fes_.ctx(); }
private:
   H1Space &fes_;
   cdm_op *op_ = nullptr;
   std::vector<double> kappa_c_, vel_c_, mass_c_;
   cdm_coeff kap_{CDM_COEFF_NONE, 0, nullptr}, vel_{CDM_COEFF_NONE, 0, nullptr}, mass_{CDM_COEFF_NONE, 0, nullptr};
   double alpha_ = 1.0;
   std::vector<int32_t> ess_;
   AssemblyLevel level_ = AssemblyLevel::PARTIAL;
};

// mfem::IterativeSolver surface.  SetOperator takes any cdm::Operator, like mfem::Solver::SetOperator(const Operator &):
// a ConvectionDiffusionForm runs the fused device-resident Krylov drivers (cdm_gmres / cdm_cg: lazily normalised basis,
// fused multi-dot / multi-axpy, ghost-consistent vectors); any other Operator runs the same algorithms through its
// virtual Mult and the C-ABI vector kernels, with an optional preconditioner Operator (SetPreconditioner).
class IterativeSolver
{
public:
   virtual ~IterativeSolver() = default;
   void SetRelTol(double v) { o_.rtol = v; }
   void SetAbsTol(double v) { o_.atol = v; }
   void SetMaxIter(int v) { o_.max_it = v; }
   void SetPrintLevel(int) {}
   void SetJacobi(bool on) { o_.jacobi = on ? 1 : 0; }          // -pc_type jacobi (ConvectionDiffusionForm only)
   void SetOperator(const Operator &op)
   {
      gen_ = &op;
      op_ = dynamic_cast<const ConvectionDiffusionForm *>(&op);
   }
   void SetPreconditioner(const Operator &M) { prec_ = &M; }    // generic operators: z = M r through M.Mult
   bool iterative_mode = false;                                  // PetscLinearSolver default
   int GetNumIterations() const { return r_.iters; }
   bool GetConverged() const { return r_.converged != 0; }
   double GetFinalNorm() const { return r_.final_norm; }
   double GetSolveSeconds() const { return r_.seconds; }
   const std::vector<double> &ResidualHistory() const { return hist_; }
   virtual void Mult(const Vector &b, Vector &x) = 0;
protected:
   template <typename F> void Run(F fn, const Vector &b, Vector &x, const char *what)
   {
      if (!op_) { throw std::runtime_error("SetOperator has not been called"); }
      o_.zero_guess = iterative_mode ? 0 : 1;
      hist_.assign((size_t)o_.max_it + 2, 0.0);
      check(op_->ctx(), fn(op_->handle(), b.Read(), x.ReadWrite(), &o_, &r_, hist_.data()), what);
      hist_.resize(r_.hist_len);
   }
   // z = M^{-1} r for a generic operator
   void Precondition(const Vector &r, Vector &z) const { if (prec_) { prec_->Mult(r, z); } else { z.Assign(r); } }
   cdm_krylov_opts o_{CDM_GMRES_PETSC, 0, 500, 1e-10, 1e-12, 1, 1};
   cdm_krylov_result r_{0, 0, 0.0, 0, 0.0};
   const ConvectionDiffusionForm *op_ = nullptr;
   const Operator *gen_ = nullptr, *prec_ = nullptr;
   std::vector<double> hist_;
};

// GMRES as configured by Input/petsc.opts (variant CDM_GMRES_PETSC) or mfem::GMRESSolver (CDM_GMRES_MFEM)
class GMRESSolver : public IterativeSolver
{
public:
   explicit GMRESSolver(int variant = CDM_GMRES_PETSC) { o_.variant = variant; }
   void SetKDim(int m) { o_.restart = m; }
   void Mult(const Vector &b, Vector &x) override
   {
      if (op_ && !prec_) { Run(cdm_gmres, b, x, "cdm_gmres"); return; }
      if (!gen_) { throw std::runtime_error("SetOperator has not been called"); }
      Generic(b, x);
   }
private:
   // left-preconditioned restarted GMRES through Operator::Mult: classical Gram-Schmidt with
   // one fused multi-dot + multi-axpy per step for the PETSc variant, modified Gram-Schmidt for the mfem variant
   void Generic(const Vector &b, Vector &x)
   {
      cdm_ctx *ctx = b.ctx();
      const int64_t n = b.Size();
      const int m = o_.restart > 0 ? o_.restart : (o_.variant == CDM_GMRES_PETSC ? 30 : 50);
      double *V = nullptr;
      check(ctx, cdm_vec_alloc(ctx, (int64_t)(m + 1) * n, &V), "cdm_vec_alloc");
      struct Free { cdm_ctx *c; double *p; ~Free() { cdm_vec_free(c, p); } } guard{ctx, V};
      Vector w, t, vj;
      w.SetSizeOn(ctx, n); t.SetSizeOn(ctx, n); vj.SetSizeOn(ctx, n);
      std::vector<double> H((size_t)(m + 1) * m, 0.0), cs(m), sn(m), s(m + 1), yv(m), hc(m + 2);
      hist_.clear();
      int it = 0, conv = 0;
      double rnorm = 0.0, ttol = 0.0;
      bool first = true;
      if (!iterative_mode) { x = 0.0; }
      auto col = [&](int j) { return V + (int64_t)j * n; };
      while (true)
      {
         if (first && !iterative_mode) { Precondition(b, w); }
         else { gen_->Mult(x, t); t.Scale(-1.0); t.Add(1.0, b); Precondition(t, w); }
         double b2; check(ctx, cdm_dot(ctx, n, w.Read(), w.Read(), &b2), "cdm_dot");
         const double beta = std::sqrt(b2);
         rnorm = beta;
         if (first)
         {
            double ref = beta;
            if (iterative_mode && o_.variant == CDM_GMRES_PETSC) { Precondition(b, t); ref = t.Norml2(); }
            ttol = std::max(o_.rtol * ref, o_.atol);
            hist_.push_back(beta);
            first = false;
         }
         if (rnorm <= ttol) { conv = 1; break; }
         if (it >= o_.max_it) { break; }
         check(ctx, cdm_vec_set(ctx, n, 0.0, col(0)), "cdm_vec_set");
         check(ctx, cdm_axpy(ctx, n, 1.0 / beta, w.Read(), col(0)), "cdm_axpy");
         std::fill(s.begin(), s.end(), 0.0);
         s[0] = beta;
         int j = 0;
         while (j < m && it < o_.max_it)
         {
            vj.CopyFromDevice(col(j));
            gen_->Mult(vj, t);
            Precondition(t, w);
            if (o_.variant == CDM_GMRES_PETSC)
            {
               check(ctx, cdm_mdot(ctx, n, j + 1, w.Read(), V, n, hc.data()), "cdm_mdot");
               check(ctx, cdm_maxpy(ctx, n, j + 1, hc.data(), V, n, w.ReadWrite()), "cdm_maxpy");
            }
            else
            {
               for (int i = 0; i <= j; i++)
               {
                  check(ctx, cdm_dot(ctx, n, w.Read(), col(i), &hc[i]), "cdm_dot");
                  check(ctx, cdm_axpy(ctx, n, -hc[i], col(i), w.ReadWrite()), "cdm_axpy");
               }
            }
            const double hn = w.Norml2();
            check(ctx, cdm_vec_set(ctx, n, 0.0, col(j + 1)), "cdm_vec_set");
            if (hn > 0.0) { check(ctx, cdm_axpy(ctx, n, 1.0 / hn, w.Read(), col(j + 1)), "cdm_axpy"); }
            hc[j + 1] = hn;
            for (int i = 0; i < j; i++)
            {
               const double a = cs[i] * hc[i] + sn[i] * hc[i + 1];
               hc[i + 1] = -sn[i] * hc[i] + cs[i] * hc[i + 1];
               hc[i] = a;
            }
            const double den = std::hypot(hc[j], hc[j + 1]);
            cs[j] = hc[j] / den; sn[j] = hc[j + 1] / den;
            hc[j] = den; hc[j + 1] = 0.0;
            s[j + 1] = -sn[j] * s[j];
            s[j] = cs[j] * s[j];
            for (int i = 0; i <= j; i++) { H[(size_t)j * (m + 1) + i] = hc[i]; }
            j++; it++;
            rnorm = std::fabs(s[j]);
            hist_.push_back(rnorm);
            if (rnorm <= ttol) { conv = 1; break; }
            if (hn == 0.0) { break; }
         }
         for (int i = j - 1; i >= 0; i--)
         {
            double a = s[i];
            for (int k = i + 1; k < j; k++) { a -= H[(size_t)k * (m + 1) + i] * yv[k]; }
            yv[i] = a / H[(size_t)i * (m + 1) + i];
         }
         for (int i = 0; i < j; i++) { yv[i] = -yv[i]; }              // cdm_maxpy subtracts
         check(ctx, cdm_maxpy(ctx, n, j, yv.data(), V, n, x.ReadWrite()), "cdm_maxpy");
         if (conv || it >= o_.max_it) { break; }
      }
      r_.iters = it; r_.converged = conv; r_.final_norm = rnorm; r_.hist_len = (int)hist_.size(); r_.seconds = 0.0;
   }
};

// mfem::CGSolver (rel 1e-12, abs 0, max 500 at mesh_recession_handler.cpp:271-273)
class CGSolver : public IterativeSolver
{
public:
   CGSolver() { o_.rtol = 1e-12; o_.atol = 0.0; o_.jacobi = 0; }
   void Mult(const Vector &b, Vector &x) override
   {
      if (op_ && !prec_) { Run(cdm_cg, b, x, "cdm_cg"); return; }
      if (!gen_) { throw std::runtime_error("SetOperator has not been called"); }
      Generic(b, x);
   }
private:
   // mfem::CGSolver::Mult through Operator::Mult
   void Generic(const Vector &b, Vector &x)
   {
      cdm_ctx *ctx = b.ctx();
      const int64_t n = b.Size();
      Vector r, d, z;
      r.SetSizeOn(ctx, n); d.SetSizeOn(ctx, n); z.SetSizeOn(ctx, n);
      hist_.clear();
      int it = 0, conv = 0;
      if (!iterative_mode) { x = 0.0; r.Assign(b); }
      else { gen_->Mult(x, r); r.Scale(-1.0); r.Add(1.0, b); }
      Precondition(r, d);
      double nom = d * r, betanom = nom, den = 0.0;
      const double r0 = std::max(nom * o_.rtol * o_.rtol, o_.atol * o_.atol);
      hist_.push_back(nom);
      if (nom <= r0) { conv = 1; }
      else
      {
         gen_->Mult(d, z);
         den = z * d;
         if (den > 0.0)
         {
            for (it = 1;; it++)
            {
               const double alpha = nom / den;
               x.Add(alpha, d);
               r.Add(-alpha, z);
               if (prec_) { prec_->Mult(r, z); betanom = r * z; }
               else { betanom = r * r; }
               hist_.push_back(betanom);
               if (betanom <= r0) { conv = 1; break; }
               if (it >= o_.max_it) { break; }
               const double beta = betanom / nom;
               d.Scale(beta);
               d.Add(1.0, prec_ ? z : r);
               gen_->Mult(d, z);
               den = d * z;
               if (den <= 0.0) { break; }
               nom = betanom;
            }
         }
      }
      r_.iters = it; r_.converged = conv; r_.final_norm = std::sqrt(std::fabs(betanom)); r_.hist_len = (int)hist_.size(); r_.seconds = 0.0;
   }
};
}  // namespace cdm
