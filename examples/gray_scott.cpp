// A C++ host of the C ABI, end to end: the documented Gray-Scott model of the reference
// (doc/docusaurus/static/ini/next/grey_scott.ini: two species, cubic reaction, RestartedGMRes + Jacobi)
// on a 2-D lattice -- config -> model -> grid -> operator -> adaptive stepper -> VTK, no Python.
//
//   g++ -std=c++17 -Iinclude examples/gray_scott.cpp -o gray_scott -Ldune_copasi_b200 -ldune_copasi_b200
//       (plus -Wl,-rpath,$PWD/dune_copasi_b200 to run it from the build tree)
//   ./gray_scott [cells per axis = 128] [t_end = 20] [output dir]
//
// On a host without a CUDA device the operator cannot be created and the program says so (exit 2):
// the library has no CPU fallback.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

#include "dune_copasi_b200.h"

namespace {
const char* kIni = R"ini(
[parser_context.bump]
type = function
expression = x, y: 0.5*exp(-100*(x^2 + y^2))
[parser_context.F]
type = constant
value = 0.0420
[parser_context.k]
type = constant
value = 0.0610
[parser_context.D]
type = constant
value = 1e-5
[compartments]
compartment.expression = 1
[model.scalar_field.U]
compartment = compartment
initial.expression = 0.7
storage.expression = 1
reaction.expression = F*(1-U) - U*V^2
reaction.jacobian.U.expression = -F - V^2
reaction.jacobian.V.expression = -2*U*V
cross_diffusion.U.expression = D*2
[model.scalar_field.V]
compartment = compartment
initial.expression = bump(0.25-position_x, 0.25-position_y) + bump(0.75-position_x, 0.75-position_y)
storage.expression = 1
reaction.expression = -(F+k)*V + U*V^2
reaction.jacobian.U.expression = V^2
reaction.jacobian.V.expression = -(F+k) + 2*U*V
cross_diffusion.V.expression = D
[model.time_step_operator]
type = Alexander2
time_step_initial = 0.1
time_step_max = 50
[model.time_step_operator.linear_solver]
type = RestartedGMRes
preconditioner.type = Jacobi
matrix_free = true
[model.time_step_operator.nonlinear_solver]
convergence_condition.relative_tolerance = 1e-8
)ini";

[[noreturn]] void die(const char* what, int code = 1) {
  std::fprintf(stderr, "%s: %s\n", what, dcb_last_error());
  std::exit(code);
}
template <class T, void (*Destroy)(T*)>
struct Handle {
  T* p;
  explicit Handle(T* q) : p(q) {}
  ~Handle() { if (p) Destroy(p); }
  Handle(const Handle&) = delete;
  Handle& operator=(const Handle&) = delete;
  operator T*() const { return p; }
};
}  // namespace

int main(int argc, char** argv) {
  const int n = argc > 1 ? std::atoi(argv[1]) : 128;
  const double t_end = argc > 2 ? std::atof(argv[2]) : 20.0;
  const std::string out = argc > 3 ? argv[3] : "";

  Handle<dcb_config, dcb_config_destroy> cfg(dcb_config_create());
  if (!cfg || dcb_config_parse_ini(cfg, kIni) != 0) die("config");
  Handle<dcb_model, dcb_model_destroy> model(dcb_model_create(cfg, 2, 0, nullptr));
  if (!model) die("model");
  const int32_t cells[2] = {n, n};
  const double origin[2] = {0, 0}, extent[2] = {1, 1};
  Handle<dcb_grid, dcb_grid_destroy> grid(dcb_grid_create_structured(2, cells, origin, extent));
  if (!grid || dcb_grid_bind(grid, model) != 0) die("grid");
  const int64_t ndofs = dcb_grid_num_dofs(grid);
  std::vector<double> u(ndofs);
  if (dcb_grid_interpolate(grid, model, 0.0, u.data()) != 0) die("initial values");

  Handle<dcb_operator, dcb_operator_destroy> op(dcb_operator_create(model, grid));
  if (!op) die("operator (this program needs a CUDA device)", 2);
  Handle<dcb_stepper, dcb_stepper_destroy> stepper(dcb_stepper_create(op, cfg, nullptr));
  if (!stepper || dcb_stepper_set_state(stepper, u.data(), 0.0) != 0) die("stepper");

  double dt = 0.1, t = 0.0;
  int accepted = 0;
  if (dcb_stepper_evolve(stepper, t_end, &dt, 1 << 30, &accepted) != 0) die("evolve");
  if (dcb_stepper_get_state(stepper, u.data(), &t) != 0) die("state");
  dcb_step_stats st;
  dcb_stepper_stats(stepper, &st);
  double umin = 1e300, umax = -1e300, vmax = -1e300;
  for (int64_t i = 0; i < ndofs; i += 2) {
    umin = std::fmin(umin, u[i]); umax = std::fmax(umax, u[i]); vmax = std::fmax(vmax, u[i + 1]);
  }
  std::printf("t = %.6g after %d steps (%lld Newton iterations, %lld Krylov half iterations, %lld kernel launches): "
              "U in [%.6f, %.6f], max V = %.6f\n", t, accepted, (long long)st.newton_iterations,
              (long long)st.linear_half_iterations, (long long)st.kernel_launches, umin, umax, vmax);
  if (!out.empty() && dcb_grid_write_vtk(grid, model, u.data(), t, out.c_str(), 0) != 0) die("vtk");
  return 0;
}
