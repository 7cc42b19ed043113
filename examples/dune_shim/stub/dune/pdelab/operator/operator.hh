// STAND-IN, not DUNE: the few names of dune-pdelab's operator interface that examples/dune_shim/b200_operator.hh
// touches (PDELab::Operator<Domain, Range> with apply(), PDELab::ErrorCondition, PDELab::Convergence::Reason), so that
// the shim of INTEGRATION.md section 2 can be compiled and run against the real C ABI on a box without the DUNE
// stack.  A DuneCopasi build puts the real <dune/pdelab/operator/operator.hh> first on the include path; nothing
// here is installed or shipped.  Shapes follow the way the reference uses them
// (dune/copasi/model/make_step_operator.hh:55-157: `ErrorCondition apply(Range&, Domain&)`, `get<T>(key)`,
// `make_error_condition(Convergence::Reason::...)`).
#pragma once
#include <any>
#include <map>
#include <string>
#include <system_error>

namespace Dune::PDELab {

using ErrorCondition = std::error_condition;

namespace Convergence {
enum class Reason { Converged = 0, DivergedNull = 1, DivergedByDivergenceTolarance = 2 };
}
inline ErrorCondition make_error_condition(Convergence::Reason r) {
  return r == Convergence::Reason::Converged ? ErrorCondition{} : ErrorCondition{static_cast<int>(r), std::generic_category()};
}

template <class Domain, class Range>
class Operator {
public:
  virtual ~Operator() = default;
  virtual ErrorCondition apply(const Domain&, Range&) { return make_error_condition(Convergence::Reason::DivergedNull); }
  // property tree of the operator ("time", "duration", "convergence_condition.relative_tolerance", ...)
  template <class T>
  T& get(const std::string& key) { return *std::any_cast<T>(&_props.at(key)); }
  template <class T>
  void set(const std::string& key, T value) { _props[key] = std::move(value); }

private:
  std::map<std::string, std::any> _props;
};

}  // namespace Dune::PDELab
