// The reference-side binding of INTEGRATION.md section 2 as a header that compiles: what a DuneCopasi maintainer would
// add as dune/copasi/model/b200_operator.hh.  It wraps the C ABI (include/dune_copasi_b200.h) in the two operator
// classes the time stepper of dune/copasi/model/make_step_operator.hh:164-444 is assembled from:
//   B200StageOperator  <-  the instationary assembler operator (`instationary_op`, :290-406): apply(x, r) additive,
//                          matrix-free derivative, matrix-based derivative values
//   B200LinearSolver   <-  LinearSolver::apply (:102-146)
// Containers: anything contiguous with data() / size() (the reference's ISTLUniformBackend<double> coefficient
// vector exposes exactly that through native(), model_single_compartment.impl.hh:146).
// Compiled and run by tests/test_capi.py against examples/dune_shim/stub/ (stand-in PDELab names) -- with a real
// DUNE stack the stub directory is simply not on the include path.
#pragma once
#include <algorithm>
#include <string>

#include <dune/pdelab/operator/operator.hh>

#include "dune_copasi_b200.h"

namespace Dune::Copasi {

namespace B200Impl {
template <class V> const double* raw(const V& v) { return v.data(); }
template <class V> double* raw(V& v) { return v.data(); }
inline PDELab::ErrorCondition ok() { return {}; }
inline PDELab::ErrorCondition fail() { return PDELab::make_error_condition(PDELab::Convergence::Reason::DivergedNull); }
}  // namespace B200Impl

// Instationary operator of one Runge-Kutta stage: r += wM * M(x) + wA * A(t, x)
template <class Coefficients, class Residual>
class B200StageOperator : public PDELab::Operator<Coefficients, Residual> {
public:
  explicit B200StageOperator(dcb_operator* op) : _op{op} {}

  // PDELab sets these per stage (make_step_operator.hh:411-442: "time", "duration", "instationary_coefficients")
  void setStage(double time, double mass_weight, double stiffness_weight) {
    _t = time; _wM = mass_weight; _wA = stiffness_weight;
  }

  // Operator::apply(x, r), additive (make_step_operator.hh:223)
  PDELab::ErrorCondition apply(const Coefficients& x, Residual& r) override {
    using namespace B200Impl;
    return dcb_residual(_op, _t, _wM, _wA, raw(x), raw(r)) ? fail() : ok();   // message: dcb_last_error()
  }

  // matrix-free derivative (MatrixFreeAdapter::apply, make_step_operator.hh:70-75): y = J(x) z
  PDELab::ErrorCondition jacobianApply(const Coefficients& x, const Coefficients& z, Residual& y) {
    using namespace B200Impl;
    std::fill_n(raw(y), y.size(), 0.);
    return dcb_jacobian_apply(_op, _t, _wM, _wA, raw(x), raw(z), raw(y)) ? fail() : ok();
  }

  // matrix-based derivative: the values of the BCRSMatrix whose pattern came from dcb_grid_pattern (sorted columns
  // == pattern.sort(); patternToMatrix, make_step_operator.hh:380-384)
  PDELab::ErrorCondition jacobian(const Coefficients& x, double* bcrs_values) {
    using namespace B200Impl;
    return dcb_jacobian(_op, _t, _wM, _wA, raw(x), bcrs_values) ? fail() : ok();
  }

private:
  dcb_operator* _op;
  double _t = 0, _wM = 0, _wA = 0;
};

// LinearSolver (make_step_operator.hh:55-157): linearise at x, then apply(b, z) solves J(x) z = b
template <class Domain, class Range>
class B200LinearSolver : public PDELab::Operator<Range, Domain> {
public:
  explicit B200LinearSolver(dcb_solver* solver) : _solver{solver} {}

  PDELab::ErrorCondition linearize(double time, double mass_weight, double stiffness_weight, const Domain& x) {
    using namespace B200Impl;
    return dcb_solver_linearize(_solver, time, mass_weight, stiffness_weight, raw(x)) ? fail() : ok();
  }

  PDELab::ErrorCondition apply(const Range& b, Domain& z) override {
    using namespace B200Impl;
    const double rel_tol = this->template get<double>("convergence_condition.relative_tolerance");   // :132
    if (dcb_solver_solve(_solver, raw(b), raw(z), rel_tol, &last) != 0 || !last.converged)
      return PDELab::make_error_condition(PDELab::Convergence::Reason::DivergedByDivergenceTolarance);   // :144
    return ok();
  }

  dcb_solve_result last{};

private:
  dcb_solver* _solver;
};

}  // namespace Dune::Copasi
