/* A plain C99 host of the C ABI (include/dune_copasi_b200.h): what a non-Python caller -- e.g. the
 * C++ shim of INTEGRATION.md -- does.  Host-side entry points run everywhere; the compute entry
 * point must either work (GPU box) or fail with "no CUDA device" (CPU box): never fall back.
 * Prints "ok <ndofs simplex> <nnz simplex> <ndofs cubes> <nnz cubes> <residual norm^2 | nodevice>". */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "dune_copasi_b200.h"

static const char* INI =
    "[compartments.domain]\n"
    "expression = 1\n"
    "[model.scalar_field.u]\n"
    "compartment = domain\n"
    "storage.expression = 1\n"
    "cross_diffusion.u.expression = 0.01\n"
    "reaction.expression = -u*v\n"
    "reaction.jacobian.u.expression = -v\n"
    "reaction.jacobian.v.expression = -u\n"
    "initial.expression = 1 + position_x\n"
    "[model.scalar_field.v]\n"
    "compartment = domain\n"
    "storage.expression = 1\n"
    "cross_diffusion.v.expression = 0.02\n"
    "reaction.expression = u*v\n"
    "reaction.jacobian.u.expression = v\n"
    "reaction.jacobian.v.expression = u\n"
    "initial.expression = 0.5\n";

#define CHECK(x) do { if (!(x)) { fprintf(stderr, "FAILED %s: %s\n", #x, dcb_last_error()); return 1; } } while (0)

int main(void) {
  dcb_config* cfg = dcb_config_create();
  CHECK(cfg && dcb_config_parse_ini(cfg, INI) == 0);
  dcb_model* model = dcb_model_create(cfg, 3, 0, NULL);
  CHECK(model && dcb_model_num_species(model) == 2 && dcb_model_num_compartments(model) == 1);
  CHECK(strcmp(dcb_model_species_name(model, 1), "v") == 0);
  const int32_t cells[3] = {4, 3, 2};
  const double origin[3] = {0, 0, 0}, extent[3] = {1, 1, 1};
  long long out[4];
  dcb_grid* grids[2];
  grids[0] = dcb_grid_create_structured(3, cells, origin, extent);
  grids[1] = dcb_grid_create_structured_cubes(3, cells, origin, extent);
  for (int g = 0; g < 2; ++g) {
    CHECK(grids[g] && dcb_grid_bind(grids[g], model) == 0);
    CHECK(dcb_grid_nodes_per_element(grids[g]) == (g == 0 ? 4 : 8));
    CHECK(dcb_grid_num_vertices(grids[g]) == 5 * 4 * 3);
    CHECK(dcb_grid_num_elements(grids[g]) == (g == 0 ? 6 : 1) * 24);
    int64_t nrows = 0, nnz = 0;
    CHECK(dcb_grid_pattern(grids[g], model, &nrows, &nnz, NULL, NULL) == 0);
    CHECK(nrows == dcb_grid_num_dofs(grids[g]) && nrows == 2 * 60);
    int64_t* rowptr = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nrows + 1));
    int32_t* colidx = (int32_t*)malloc(sizeof(int32_t) * (size_t)nnz);
    CHECK(dcb_grid_pattern(grids[g], model, &nrows, &nnz, rowptr, colidx) == 0);
    CHECK(rowptr[0] == 0 && rowptr[nrows] == nnz);
    free(rowptr); free(colidx);
    out[2 * g] = (long long)nrows; out[2 * g + 1] = (long long)nnz;
  }
  /* compute: works on a GPU box, fails loudly elsewhere */
  int64_t n = dcb_grid_num_dofs(grids[0]);
  double* u = (double*)malloc(sizeof(double) * (size_t)n);
  double* r = (double*)calloc((size_t)n, sizeof(double));
  CHECK(dcb_grid_interpolate(grids[0], model, 0.0, u) == 0);
  char tail[64];
  dcb_operator* op = dcb_operator_create(model, grids[0]);
  if (dcb_device_count() > 0) {
    CHECK(op && dcb_residual(op, 0.0, 1.0, 0.5, u, r) == 0);
    double s = 0;
    for (int64_t i = 0; i < n; ++i) s += r[i] * r[i];
    CHECK(s > 0);
    snprintf(tail, sizeof tail, "%.17g", s);
    dcb_operator_destroy(op);
  } else {
    CHECK(op == NULL && strstr(dcb_last_error(), "no CUDA device") != NULL);
    snprintf(tail, sizeof tail, "nodevice");
  }
  printf("ok %lld %lld %lld %lld %s\n", out[0], out[1], out[2], out[3], tail);
  free(u); free(r);
  dcb_grid_destroy(grids[0]); dcb_grid_destroy(grids[1]);
  dcb_model_destroy(model);
  dcb_config_destroy(cfg);
  return 0;
}
