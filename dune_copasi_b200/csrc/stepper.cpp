#include "stepper.hpp"

#include <cmath>

#include "comm.hpp"

namespace dcb {

StepOperator::StepOperator(std::shared_ptr<DeviceOperator> o, const PTree& cfg, Communicator* comm)
    : op(std::move(o)), comm_(comm) {
  rk_type = cfg.get("type", std::string("Alexander2"));   // make_step_operator.hh:413
  if (rk_type == "ImplicitEuler") {
    a_ = {{-1.0, 1.0}};
    b_ = {{0.0, 1.0}};
    d_ = {0.0, 1.0};
  } else if (rk_type == "Alexander2") {
    const double al = 1.0 - std::sqrt(2.0) / 2.0;
    a_ = {{-1.0, 1.0, 0.0}, {-1.0, 0.0, 1.0}};
    b_ = {{0.0, al, 0.0}, {0.0, 1.0 - al, al}};
    d_ = {0.0, al, 1.0};
  } else if (rk_type == "ExplicitEuler") {
    a_ = {{-1.0, 1.0}};
    b_ = {{1.0, 0.0}};
    d_ = {0.0, 1.0};
  } else if (rk_type == "Heun") {
    a_ = {{-1.0, 1.0, 0.0}, {-0.5, -0.5, 1.0}};
    b_ = {{1.0, 0.0, 0.0}, {0.0, 0.5, 0.0}};
    d_ = {0.0, 1.0, 1.0};
  } else if (rk_type == "Shu3") {
    a_ = {{-1.0, 1.0, 0.0, 0.0}, {-0.75, -0.25, 1.0, 0.0}, {-1.0 / 3.0, 0.0, -2.0 / 3.0, 1.0}};
    b_ = {{1.0, 0.0, 0.0, 0.0}, {0.0, 0.25, 0.0, 0.0}, {0.0, 0.0, 2.0 / 3.0, 0.0}};
    d_ = {0.0, 1.0, 0.5, 1.0};
  } else if (rk_type == "RungeKutta4") {
    a_ = {{-1.0, 1.0, 0.0, 0.0, 0.0}, {-1.0, 0.0, 1.0, 0.0, 0.0}, {-1.0, 0.0, 0.0, 1.0, 0.0}, {-1.0, 0.0, 0.0, 0.0, 1.0}};
    b_ = {{0.5, 0.0, 0.0, 0.0, 0.0}, {0.0, 0.5, 0.0, 0.0, 0.0}, {0.0, 0.0, 1.0, 0.0, 0.0},
          {1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0, 0.0}};
    d_ = {0.0, 0.5, 0.5, 1.0, 1.0};
  } else if (rk_type == "Alexander3") {
    const double al = 0.4358665215;
    const double b1 = -(6.0 * al * al - 16.0 * al + 1.0) / 4.0, b2 = (6.0 * al * al - 20.0 * al + 5.0) / 4.0;
    a_ = {{-1.0, 1.0, 0.0, 0.0}, {-1.0, 0.0, 1.0, 0.0}, {-1.0, 0.0, 0.0, 1.0}};
    b_ = {{0.0, al, 0.0, 0.0}, {0.0, (1.0 - al) / 2.0, al, 0.0}, {0.0, b1, b2, al}};
    d_ = {0.0, al, (1.0 + al) / 2.0, 1.0};
  } else if (rk_type == "FractionalStepTheta") {
    // make_step_operator.hh:434-435: sub-steps theta, 1-2theta, theta; implicit weight alpha*theta
    const double th = 1.0 - 0.5 * std::sqrt(2.0), alpha = 2.0 - std::sqrt(2.0), beta = 1.0 - alpha;
    a_ = {{-1.0, 1.0, 0.0, 0.0}, {0.0, -1.0, 1.0, 0.0}, {0.0, 0.0, -1.0, 1.0}};
    b_ = {{beta * th, alpha * th, 0.0, 0.0}, {0.0, alpha * (1.0 - 2.0 * th), alpha * th, 0.0},
          {0.0, 0.0, beta * th, alpha * th}};
    d_ = {0.0, th, 1.0 - th, 1.0};
  } else {
    fail("time_step_operator.type = '", rk_type,
         "' is not built (available: ExplicitEuler, ImplicitEuler, Heun, Shu3, RungeKutta4, Alexander2, "
         "FractionalStepTheta, Alexander3)");
  }
  is_linear = op->model->is_linear;
  const PTree& ls = cfg.sub("linear_solver");
  lin_rel = ls.get("convergence_condition.relative_tolerance", 1e-4);   // :203
  const PTree& nl = cfg.sub("nonlinear_solver");
  newton_rel = nl.get("convergence_condition.relative_tolerance", 1e-4);   // :253
  newton_abs = nl.get("convergence_condition.absolute_tolerance", 0.0);
  newton_max_it = (int)nl.get_vec("convergence_condition.iteration_range", {0, 40}).back();   // :254-255
  dx_fixed_tol = nl.get("dx_inverse_fixed_tolerance", false);
  dx_min_rel_tol = nl.get("dx_inverse_min_relative_tolerance", 0.1);
  std::string norm = nl.get("norm", std::string("l_2"));
  if (norm != "l_2") fail("nonlinear_solver.norm = '", norm, "' is not built (only l_2, i.e. the squared 2-norm)");
  has_dt_min = cfg.has_key("time_step_min");   // optional in the reference (src/dune_copasi.cc:389-396)
  dt_min = cfg.get("time_step_min", 0.0);
  dt_max = cfg.get("time_step_max", 0.0);
  inc_factor = cfg.get("time_step_increase_factor", 1.1);
  dec_factor = cfg.get("time_step_decrease_factor", 0.5);
  linear = std::make_unique<LinearSolver>(op, ls, comm);
  const int64_t n = op->ndofs;
  stage_.resize(a_.size());
  for (auto& s : stage_) s.alloc(n);
  const_.alloc(n); r_.alloc(n); z_.alloc(n);
  scal_.alloc(4); hscal_.alloc(4);
  la::reduce_workspace_create(&ws_);
}

double StepOperator::norm2(const double* r) {
  cudaStream_t s = op->stream;
  la::dot(op->owned, r, r, scal_.p, ws_, s);
  op->stats.launches++;
  if (comm_) comm_->allreduce_sum(scal_.p, 1, s);
  DCB_CUDA(cudaMemcpyAsync(hscal_.p, scal_.p, sizeof(double), cudaMemcpyDeviceToHost, s));
  { DeviceOperator::HostTimer ht(op.get(), "host_wait"); DCB_CUDA(cudaStreamSynchronize(s)); }
  return hscal_.p[0];
}

void StepOperator::stage_residual(const double* x, double ts, double wM, double wA, const double* constant, double* r) {
  cudaStream_t s = op->stream;
  la::copy(op->ndofs, constant, r, s);
  op->stats.launches++;
  op->residual(ts, wM, wA, x, r);
  if (op->ncons) { la::zero_values(op->ncons, op->cdofs.p, r, s); op->stats.launches++; }
  stats.residual_evaluations++;
}

bool StepOperator::solve_stage(double* x, double ts, double wM, double wA, const double* constant) {
  cudaStream_t s = op->stream;
  const int64_t n = op->ndofs;
  auto correct = [&](double tol) {
    linear->linearize(ts, wM, wA, x);
    stats.linearizations++;
    SolveResult res = linear->apply(r_.p, z_.p, tol);
    stats.linear_solves++;
    stats.linear_iterations += res.iterations;
    stats.linear_half_iterations += res.half_iterations;
    if (!res.converged) return false;
    la::axpy(n, -1.0, z_.p, x, s);   // x -= z
    op->stats.launches++;
    return true;
  };
  if (comm_) comm_->halo_update(x, s);
  stage_residual(x, ts, wM, wA, constant, r_.p);
  if (is_linear) {
    // one defect-correction solve per stage (make_step_operator.hh:215-243)
    return correct(lin_rel);
  }
  // Newton; "defect" is the squared 2-norm (make_step_operator.hh:274-276), classic PDELab
  // control of the linear tolerance unless dx_inverse_fixed_tolerance
  double cur = norm2(r_.p), first = cur, prev = cur;
  const double stop = std::max(first * newton_rel, newton_abs);
  int it = 0;
  while (cur > stop) {
    if (it >= newton_max_it || !std::isfinite(cur)) return false;
    double tol;
    if (dx_fixed_tol) tol = lin_rel;
    else if (stop / (10 * cur) > cur * cur / (prev * prev)) tol = stop / (10 * cur);
    else tol = std::min(dx_min_rel_tol, cur * cur / (prev * prev));
    if (!correct(tol)) return false;
    stage_residual(x, ts, wM, wA, constant, r_.p);
    prev = cur;
    cur = norm2(r_.p);
    ++it;
    stats.newton_iterations++;
  }
  return true;
}

bool StepOperator::step(double* u, double t, double dt) {
  DeviceOperator::HostTimer whole(op.get(), "host_step");
  cudaStream_t s = op->stream;
  const int64_t n = op->ndofs;
  const size_t nst = a_.size();
  if (comm_) comm_->halo_update(u, s);
  for (size_t st = 0; st < nst; ++st) {
    // constant part: contributions of the already known stages
    la::fill(n, 0.0, const_.p, s);
    op->stats.launches++;
    for (size_t j = 0; j <= st; ++j) {
      const double* uj = j == 0 ? u : stage_[j - 1].p;
      const double wM = a_[st][j], wA = dt * b_[st][j];
      if (wM != 0.0 || wA != 0.0) op->residual(t + d_[j] * dt, wM, wA, uj, const_.p);
    }
    double* x = stage_[st].p;
    la::copy(n, st == 0 ? u : stage_[st - 1].p, x, s);
    op->stats.launches++;
    if (op->ncons) { la::set_values(op->ncons, op->cdofs.p, op->cvals.p, x, s); op->stats.launches++; }
    stats.stages++;
    if (!solve_stage(x, t + d_[st + 1] * dt, a_[st][st + 1], dt * b_[st][st + 1], const_.p)) {
      stats.failed_steps++;
      return false;
    }
  }
  la::copy(n, stage_[nst - 1].p, u, s);
  op->stats.launches++;
  stats.steps++;
  return true;
}

// dune-common FloatCmp with its defaults (relativeWeak, epsilon = 8 ulp): what stepper.hh compares times with
static bool fc_eq(double a, double b) {
  return std::fabs(a - b) <= 8.0 * 2.220446049250313e-16 * std::max(std::fabs(a), std::fabs(b));
}
static bool fc_le(double a, double b) { return a < b || fc_eq(a, b); }
static bool fc_lt(double a, double b) { return a < b && !fc_eq(a, b); }

// SimpleAdaptiveStepper::check_dt (stepper.hh:375-386): a step outside [time_step_min, time_step_max] is an error
bool StepOperator::check_dt(double dt) const {
  if (has_dt_min && fc_lt(std::fabs(dt), std::fabs(dt_min))) return false;
  if (dt_max > 0 && !fc_le(std::fabs(dt), std::fabs(dt_max))) return false;
  return true;
}

// SimpleAdaptiveStepper::do_step (stepper.hh:337-368): retry with dt * decrease_factor until the step
// succeeds or the lower limit is reached; after a success dt grows by increase_factor (clamped to the maximum)
bool StepOperator::do_step(double* u, double* t, double* dt) {
  if (!check_dt(*dt)) return false;
  bool ok = step(u, *t, *dt);
  while (!ok) {
    *dt *= dec_factor;
    if (!check_dt(*dt)) return false;
    ok = step(u, *t, *dt);
  }
  *t += *dt;
  double next = *dt * inc_factor;
  if (dt_max > 0) next = std::min(std::max(next, -std::fabs(dt_max)), std::fabs(dt_max));
  *dt = next;
  return true;
}

// TimeStepper::evolve with snap_to_end_time (stepper.hh:145-176) and snap_to_time (:192-239): full steps
// while two of them still fit, then the remainder in ceil(remainder / dt) equal steps (recomputed after
// every step, since dt keeps growing), halving dt after a failure, at most 100 times.
// max_steps bounds the number of accepted steps of this call (an extension for callers that interleave output).
int StepOperator::evolve(double* u, double* t, double t_end, double* dt, int max_steps) {
  int accepted = 0;
  if (snap_target_ != t_end) { snapping_ = false; snap_target_ = t_end; snap_count_ = 0; }
  while (!snapping_ && accepted < max_steps) {
    if (!fc_le(*t + 2.0 * *dt, t_end)) { snapping_ = true; break; }
    if (!do_step(u, t, dt)) fail("Evolving system could not approach final time (t = ", *t, ", dt = ", *dt, ")");
    ++accepted;
  }
  while (snapping_ && accepted < max_steps && fc_lt(*t, t_end)) {
    const int n = (int)std::ceil((t_end - *t) / *dt);
    if (n <= 0) fail("Timestep doesn't make advances towards snap step");
    *dt = (t_end - *t) / n;
    const double t_before = *t, dt_try = *dt;
    if (do_step(u, t, dt)) {
      ++accepted;
    } else {
      *t = t_before;
      if (snap_count_++ == 100) fail("Snapping time exceeded maximum iteration count");
      *dt = dt_try * 0.5;
    }
  }
  if (!fc_lt(*t, t_end)) { snapping_ = false; snap_target_ = -1e300; snap_count_ = 0; }   // arrived: the next call starts afresh
  return accepted;
}

}  // namespace dcb
