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
  dt_min = cfg.get("time_step_min", 1e-12);
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

int StepOperator::evolve(double* u, double* t, double t_end, double* dt, int max_steps) {
  int accepted = 0;
  while (t_end - *t > 1e-12 * std::max(1.0, std::fabs(t_end)) && accepted < max_steps) {
    double dt_try = std::min(*dt, t_end - *t);   // snap to the end time (stepper.hh:192-239)
    for (;;) {
      if (step(u, *t, dt_try)) break;
      dt_try *= dec_factor;   // stepper.hh:350-357
      if (dt_try < dt_min) fail("time step underflow at t = ", *t);
    }
    *t += dt_try;
    ++accepted;
    double next = dt_try * inc_factor;   // stepper.hh:360-366
    if (dt_max > 0) next = std::min(next, dt_max);
    *dt = next;
  }
  return accepted;
}

}  // namespace dcb
