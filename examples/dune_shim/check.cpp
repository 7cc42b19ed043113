// Compiles the DUNE-side binding (b200_operator.hh) against the stand-in PDELab names and drives it through the real C
// ABI: residual and matrix-free derivative through B200StageOperator, a linear solve through B200LinearSolver.
//   g++ -std=c++17 -Iinclude -Iexamples/dune_shim/stub -Iexamples/dune_shim examples/dune_shim/check.cpp
//       -Ldune_copasi_b200 -ldune_copasi_b200 -Wl,-rpath,$PWD/dune_copasi_b200 -o shim_check
// Without a CUDA device the operator cannot be created (no CPU fallback): exit code 2 and the library's message.
#include <cmath>
#include <cstdio>
#include <vector>

#include "b200_operator.hh"

static const char* kIni = R"ini(
[compartments]
domain.expression = 1
[parser_context.F]
type = constant
value = 0.042
[parser_context.k]
type = constant
value = 0.061
[model.scalar_field.U]
compartment = domain
initial.expression = 0.7 + 0.1*position_x
storage.expression = 1
reaction.expression = F*(1-U) - U*V^2
reaction.jacobian.U.expression = -F - V^2
reaction.jacobian.V.expression = -2*U*V
cross_diffusion.U.expression = 2e-3
[model.scalar_field.V]
compartment = domain
initial.expression = 0.2 + 0.1*position_y*position_z
storage.expression = 1
reaction.expression = -(F+k)*V + U*V^2
reaction.jacobian.U.expression = V^2
reaction.jacobian.V.expression = -(F+k) + 2*U*V
cross_diffusion.V.expression = 1e-3
)ini";

int main() {
  using Vec = std::vector<double>;
  dcb_config* cfg = dcb_config_create();
  if (dcb_config_parse_ini(cfg, kIni)) { fprintf(stderr, "%s\n", dcb_last_error()); return 1; }
  const int32_t cells[3] = {12, 10, 8};
  const double origin[3] = {0, 0, 0}, extent[3] = {1.0, 0.9, 0.8};
  dcb_grid* grid = dcb_grid_create_structured(3, cells, origin, extent);
  dcb_model* model = dcb_model_create(cfg, 3, 0, nullptr);
  if (!grid || !model || dcb_grid_bind(grid, model)) { fprintf(stderr, "%s\n", dcb_last_error()); return 1; }
  dcb_operator* op = dcb_operator_create(model, grid);
  if (!op) { fprintf(stderr, "%s\n", dcb_last_error()); return 2; }
  const size_t n = (size_t)dcb_operator_num_dofs(op);
  Vec x(n), z(n), r(n, 0.0), r_direct(n, 0.0), y(n), b(n), sol(n, 0.0), check(n);
  dcb_grid_interpolate(grid, model, 0.0, x.data());
  for (size_t i = 0; i < n; ++i) z[i] = std::sin(0.37 * (double)i);

  Dune::Copasi::B200StageOperator<Vec, Vec> stage(op);
  stage.setStage(0.0, 1.0, 0.25);
  if (stage.apply(x, r)) { fprintf(stderr, "apply: %s\n", dcb_last_error()); return 1; }
  dcb_residual(op, 0.0, 1.0, 0.25, x.data(), r_direct.data());
  double dr = 0, nr = 0;
  for (size_t i = 0; i < n; ++i) { dr += (r[i] - r_direct[i]) * (r[i] - r_direct[i]); nr += r_direct[i] * r_direct[i]; }
  if (stage.jacobianApply(x, z, y)) { fprintf(stderr, "jacobianApply: %s\n", dcb_last_error()); return 1; }

  dcb_config* lcfg = dcb_config_create();
  dcb_config_set(lcfg, "type", "BiCGSTAB");
  dcb_config_set(lcfg, "preconditioner.type", "Jacobi");
  dcb_config_set(lcfg, "matrix_free", "true");
  dcb_solver* solver = dcb_solver_create(op, lcfg, nullptr);
  if (!solver) { fprintf(stderr, "%s\n", dcb_last_error()); return 1; }
  Dune::Copasi::B200LinearSolver<Vec, Vec> lin(solver);
  lin.set<double>("convergence_condition.relative_tolerance", 1e-10);
  b = y;   // right-hand side with the known solution z
  if (lin.linearize(0.0, 1.0, 0.25, x) || lin.apply(b, sol)) { fprintf(stderr, "solve: %s\n", dcb_last_error()); return 1; }
  double de = 0, ne = 0;
  for (size_t i = 0; i < n; ++i) { de += (sol[i] - z[i]) * (sol[i] - z[i]); ne += z[i] * z[i]; }
  printf("shim ok: dofs %zu, residual via shim vs ABI %.1e, solve error %.1e after %d iterations\n", n,
         std::sqrt(dr / nr), std::sqrt(de / ne), (int)lin.last.iterations);
  // (two runs of the same residual differ by the order of the fp64 atomics: rounding, not zero)
  const bool good = std::sqrt(dr / nr) < 1e-13 && std::sqrt(de / ne) < 1e-7;
  dcb_solver_destroy(solver); dcb_config_destroy(lcfg); dcb_operator_destroy(op); dcb_model_destroy(model);
  dcb_grid_destroy(grid); dcb_config_destroy(cfg);
  return good ? 0 : 3;
}
