// Time stepping glue on the device: RungeKutta o (Newton | linear defect correction) o LinearSolver
// o instationary operator, as wired by dune/copasi/model/make_step_operator.hh:164-444, plus the
// adaptive step-size control of dune/copasi/common/stepper.hh:337-368.  State and all work
// vectors stay resident in HBM; the host drives the control flow (a handful of scalars per
// Newton / Krylov iteration cross PCIe).
#pragma once
#include <memory>
#include <vector>

#include "solver.hpp"

namespace dcb {

struct StepStats {
  long long steps = 0, failed_steps = 0, stages = 0;
  long long newton_iterations = 0, linear_solves = 0, linear_iterations = 0, linear_half_iterations = 0;
  long long residual_evaluations = 0, linearizations = 0;
};

class StepOperator {
 public:
  StepOperator(std::shared_ptr<DeviceOperator> op, const PTree& time_step_cfg, Communicator* comm = nullptr);

  // u (device, ndofs) is advanced from t by dt in place; returns false if the step failed
  // (Newton / linear solver did not converge) in which case u is unchanged.
  bool step(double* u, double t, double dt);
  // adaptive evolution to t_end (TimeStepper::evolve + snap_to_time over SimpleAdaptiveStepper::do_step,
  // common/stepper.hh:145-239, 337-386): returns the number of accepted steps
  int evolve(double* u, double* t, double t_end, double* dt, int max_steps);
  // one adaptive step: u, t advanced by the dt that succeeded, dt replaced by the next suggestion
  bool do_step(double* u, double* t, double* dt);
  bool check_dt(double dt) const;

  std::string rk_type;
  bool is_linear = false;
  double newton_rel = 1e-4, newton_abs = 0.0, lin_rel = 1e-4;
  int newton_max_it = 40;
  bool dx_fixed_tol = false;
  double dx_min_rel_tol = 0.1;
  double dt_min = 0.0, dt_max = 0.0, inc_factor = 1.1, dec_factor = 0.5;
  bool has_dt_min = false;
  StepStats stats;
  std::shared_ptr<DeviceOperator> op;
  std::unique_ptr<LinearSolver> linear;

 private:
  bool solve_stage(double* x, double ts, double wM, double wA, const double* constant);
  void stage_residual(const double* x, double ts, double wM, double wA, const double* constant, double* r);
  double norm2(const double* r);
  Communicator* comm_;
  bool snapping_ = false;       // evolve() has entered snap_to_time for snap_target_
  double snap_target_ = -1e300;
  int snap_count_ = 0;
  std::vector<std::vector<double>> a_, b_;
  std::vector<double> d_;
  std::vector<DeviceBuffer<double>> stage_;   // stage solutions u_1..u_s
  DeviceBuffer<double> const_, r_, z_, scal_;
  PinnedBuffer<double> hscal_;
  la::ReduceWorkspace ws_;
};

}  // namespace dcb
