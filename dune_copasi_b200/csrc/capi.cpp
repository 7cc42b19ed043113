// extern "C" boundary: see include/dune_copasi_b200.h for the contract.
#include "../../include/dune_copasi_b200.h"

#include <cstring>
#include <map>
#include <memory>

#include "comm.hpp"
#include "grid.hpp"
#include "jit.hpp"
#include "model.hpp"
#include "operator.hpp"
#include "reduce.hpp"
#include "solver.hpp"
#include "stepper.hpp"

using namespace dcb;

struct dcb_config { PTree tree; };
struct dcb_model { std::shared_ptr<Model> m; std::string source; };
struct dcb_grid {
  std::shared_ptr<Grid> g;
  std::vector<int64_t> rowptr;
  std::vector<int32_t> colidx;
  std::map<std::string, std::vector<double>> vtk_timesteps;   // per output path, as Model::_writer_timesteps
};
struct dcb_comm { std::unique_ptr<Communicator> c; };
struct dcb_operator {
  std::shared_ptr<DeviceOperator> op;
  DeviceBuffer<double> x, z, r;   // staging for the host-pointer entry points
  void stage() {
    if (!x.n) { x.alloc(op->ndofs); z.alloc(op->ndofs); r.alloc(op->ndofs); }
  }
};
struct dcb_solver {
  std::unique_ptr<LinearSolver> s;
  dcb_operator* op;
  DeviceBuffer<double> x, b, z;
};
struct dcb_stepper {
  std::unique_ptr<StepOperator> s;
  dcb_operator* op;
  DeviceBuffer<double> u;
  double time = 0;
};
struct dcb_reducer {
  std::unique_ptr<Reducer> r;
  dcb_operator* op;
  DeviceBuffer<double> x;
};

namespace {
thread_local std::string g_error;

template <class F>
int guard(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_error = e.what();
    return 1;
  } catch (...) {
    g_error = "unknown error";
    return 1;
  }
}
template <class T, class F>
T* guard_new(F&& f) {
  T* out = nullptr;
  if (guard([&] { out = f(); })) return nullptr;
  return out;
}
std::vector<std::string> key_vec(int nkeys, const char* const* keys) {
  std::vector<std::string> v;
  for (int k = 0; k < nkeys; ++k) v.emplace_back(keys[k]);
  return v;
}
}  // namespace

extern "C" {

int dcb_version(void) { return 100; }
const char* dcb_last_error(void) { return g_error.c_str(); }
int dcb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

// ---- config
dcb_config* dcb_config_create(void) { return new dcb_config(); }
void dcb_config_destroy(dcb_config* c) { delete c; }
int dcb_config_parse_ini(dcb_config* c, const char* text) { return guard([&] { c->tree.parse_ini(text); }); }
int dcb_config_set(dcb_config* c, const char* key, const char* value) { return guard([&] { c->tree.set(key, value); }); }
size_t dcb_config_dump(const dcb_config* c, char* out, size_t cap) {
  std::string s = c->tree.dump();
  if (out && cap) { size_t n = std::min(cap - 1, s.size()); std::memcpy(out, s.data(), n); out[n] = 0; }
  return s.size() + 1;
}

// ---- grid
dcb_grid* dcb_grid_create_structured(int dim, const int32_t* cells, const double* origin, const double* extent) {
  return guard_new<dcb_grid>([&] {
    auto* g = new dcb_grid();
    g->g = std::make_shared<Grid>(Grid::structured(dim, cells, origin, extent));
    return g;
  });
}
dcb_grid* dcb_grid_create_structured_cubes(int dim, const int32_t* cells, const double* origin, const double* extent) {
  return guard_new<dcb_grid>([&] {
    auto* g = new dcb_grid();
    g->g = std::make_shared<Grid>(Grid::structured(dim, cells, origin, extent, 1));
    return g;
  });
}
dcb_grid* dcb_grid_create(int dim, int64_t nv, const double* coords, int64_t ne, const int32_t* elems,
                          int nkeys, const char* const* keys, const double* cell_data) {
  return guard_new<dcb_grid>([&] {
    auto* g = new dcb_grid();
    g->g = std::make_shared<Grid>(Grid::from_arrays(dim, nv, coords, ne, elems, key_vec(nkeys, keys), cell_data));
    return g;
  });
}
void dcb_grid_destroy(dcb_grid* g) { delete g; }
int dcb_grid_dim(const dcb_grid* g) { return g->g->dim; }
int64_t dcb_grid_num_vertices(const dcb_grid* g) { return g->g->nv; }
int64_t dcb_grid_num_elements(const dcb_grid* g) { return g->g->ne; }
int dcb_grid_nodes_per_element(const dcb_grid* g) { return g->g->nd(); }
int dcb_grid_get_coords(const dcb_grid* g, double* c) {
  std::memcpy(c, g->g->coords.data(), g->g->coords.size() * sizeof(double));
  return 0;
}
int dcb_grid_get_elements(const dcb_grid* g, int32_t* e) {
  std::memcpy(e, g->g->elems.data(), g->g->elems.size() * sizeof(int32_t));
  return 0;
}

// ---- model
dcb_model* dcb_model_create(const dcb_config* c, int dim, int nkeys, const char* const* keys) {
  return guard_new<dcb_model>([&] {
    auto* m = new dcb_model();
    m->m = std::make_shared<Model>(c->tree, dim, key_vec(nkeys, keys));
    return m;
  });
}
void dcb_model_destroy(dcb_model* m) { delete m; }
int dcb_model_num_compartments(const dcb_model* m) { return m->m->ncomp(); }
int dcb_model_num_species(const dcb_model* m) { return m->m->nspec(); }
const char* dcb_model_species_name(const dcb_model* m, int s) { return m->m->species.at(s).name.c_str(); }
int dcb_model_species_compartment(const dcb_model* m, int s) { return m->m->species.at(s).comp; }
const char* dcb_model_cuda_source(dcb_model* m) {
  if (guard([&] { m->source = jit_source(*m->m, jit_defines(*m->m)); })) return nullptr;
  return m->source.c_str();
}
const char* dcb_model_cuda_source_group(dcb_model* m, int group) {
  if (guard([&] {
        if (group < 0 || group > (int)JitGroup::TileQ1) fail("kernel group ", group, " does not exist");
        m->source = jit_source(*m->m, jit_defines(*m->m), (JitGroup)group);
      })) return nullptr;
  return m->source.c_str();
}
int64_t dcb_model_compile(dcb_model* m, int kind, char* out, size_t cap) {
  int64_t n = -1;
  guard([&] {
    std::string log;
    std::vector<char> bin = jit_compile(jit_source(*m->m, jit_defines(*m->m)), &log, kind == 1);
    n = (int64_t)bin.size();
    if (out && cap) std::memcpy(out, bin.data(), std::min(cap, bin.size()));
  });
  return n;
}

int dcb_model_precompile(dcb_model* m) {
  return guard([&] {
    std::string defs = jit_defines(*m->m);
    for (JitGroup g : {JitGroup::Patch, JitGroup::Element, JitGroup::Csr, JitGroup::Skeleton, JitGroup::Structured,
                       JitGroup::StructuredQ1, JitGroup::Tile, JitGroup::TileQ1})
      jit_compile_cached(jit_source(*m->m, defs, g));
  });
}

int dcb_model_precompile_group(dcb_model* m, int group) {
  return guard([&] {
    if (group < 0 || group > (int)JitGroup::TileQ1) fail("kernel group ", group, " does not exist");
    jit_compile_cached(jit_source(*m->m, jit_defines(*m->m), (JitGroup)group));
  });
}

// ---- binding
int dcb_grid_bind(dcb_grid* g, const dcb_model* m) {
  return guard([&] { g->g->bind(*m->m); g->rowptr.clear(); g->colidx.clear(); });
}
int64_t dcb_grid_num_dofs(const dcb_grid* g) { return g->g->ndofs; }
int dcb_grid_get_elem_compartment(const dcb_grid* g, int32_t* out) {
  std::memcpy(out, g->g->elem_comp.data(), g->g->elem_comp.size() * sizeof(int32_t));
  return 0;
}
int dcb_grid_get_elem_dof(const dcb_grid* g, int64_t* out) {
  const Grid& G = *g->g;
  for (int64_t e = 0; e < G.ne; ++e)
    for (int a = 0; a < G.nd(); ++a) out[e * G.nd() + a] = G.elem_dof(e, a);
  return 0;
}
int64_t dcb_grid_num_facets(const dcb_grid* g) { return (int64_t)g->g->f_in.size(); }
int dcb_grid_get_facets(const dcb_grid* g, int64_t* f_in, int64_t* f_out, int32_t* f_lin, int32_t* f_lout) {
  const Grid& G = *g->g;
  size_t n = G.f_in.size();
  std::memcpy(f_in, G.f_in.data(), n * 8); std::memcpy(f_out, G.f_out.data(), n * 8);
  std::memcpy(f_lin, G.f_lin.data(), n * 4); std::memcpy(f_lout, G.f_lout.data(), n * 4);
  return 0;
}
int dcb_grid_pattern(dcb_grid* g, const dcb_model* m, int64_t* nrows, int64_t* nnz, int64_t* rowptr, int32_t* colidx) {
  return guard([&] {
    if (g->rowptr.empty()) g->g->pattern(*m->m, g->rowptr, g->colidx);
    if (nrows) *nrows = g->g->ndofs;
    if (nnz) *nnz = (int64_t)g->colidx.size();
    if (rowptr) std::memcpy(rowptr, g->rowptr.data(), g->rowptr.size() * 8);
    if (colidx) std::memcpy(colidx, g->colidx.data(), g->colidx.size() * 4);
  });
}
int dcb_grid_interpolate(const dcb_grid* g, const dcb_model* m, double time, double* u) {
  return guard([&] {
    std::vector<double> v;
    g->g->interpolate(*m->m, time, v);
    std::memcpy(u, v.data(), v.size() * 8);
  });
}
int64_t dcb_grid_constraints(const dcb_grid* g, const dcb_model* m, int32_t* dofs, double* vals, int64_t cap) {
  int64_t n = -1;
  guard([&] {
    std::vector<int32_t> d;
    std::vector<double> v;
    g->g->constraints(*m->m, d, v);
    n = (int64_t)d.size();
    int64_t k = std::min(cap, n);
    if (dofs && k > 0) std::memcpy(dofs, d.data(), k * 4);
    if (vals && k > 0) std::memcpy(vals, v.data(), k * 8);
  });
  return n;
}

// ---- operator
dcb_operator* dcb_operator_create(dcb_model* m, dcb_grid* g) {
  return guard_new<dcb_operator>([&] {
    auto* o = new dcb_operator();
    o->op = std::make_shared<DeviceOperator>(m->m, g->g);
    return o;
  });
}
void dcb_operator_destroy(dcb_operator* o) { delete o; }
int64_t dcb_operator_num_dofs(const dcb_operator* o) { return o->op->ndofs; }
int64_t dcb_operator_nnz(dcb_operator* o) {
  int64_t n = -1;
  guard([&] { o->op->ensure_csr(); n = o->op->nnz(); });
  return n;
}
int64_t dcb_operator_launches(const dcb_operator* o) { return o->op->stats.launches; }
void* dcb_operator_stream(const dcb_operator* o) { return (void*)o->op->stream; }
int dcb_operator_sync(dcb_operator* o) { return guard([&] { DCB_CUDA(cudaStreamSynchronize(o->op->stream)); }); }

int dcb_operator_profile(dcb_operator* o, int enable) { return guard([&] { o->op->profile_enable(enable != 0); }); }
size_t dcb_operator_profile_report(dcb_operator* o, char* out, size_t cap) {
  std::string s;
  guard([&] {
    for (auto& kv : o->op->profile_collect())
      s += kv.first + " " + std::to_string(kv.second.first) + " " + std::to_string(kv.second.second) + "\n";
  });
  if (out && cap) { size_t n = std::min(cap - 1, s.size()); std::memcpy(out, s.data(), n); out[n] = 0; }
  return s.size() + 1;
}
int dcb_residual(dcb_operator* o, double t, double wM, double wA, const double* x, double* r) {
  return guard([&] {
    o->stage();
    cudaStream_t s = o->op->stream;
    o->x.upload(x, o->op->ndofs, s);
    o->r.upload(r, o->op->ndofs, s);
    o->op->residual(t, wM, wA, o->x.p, o->r.p);
    o->r.download(r, s);
  });
}
int dcb_jacobian(dcb_operator* o, double t, double wM, double wA, const double* x, double* vals) {
  return guard([&] {
    o->stage();
    cudaStream_t s = o->op->stream;
    o->op->ensure_csr();
    DeviceBuffer<double> v(o->op->nnz());
    v.zero(s);
    o->x.upload(x, o->op->ndofs, s);
    o->op->jacobian_csr(t, wM, wA, o->x.p, v.p);
    v.download(vals, s);
  });
}
int dcb_jacobian_apply(dcb_operator* o, double t, double wM, double wA, const double* x, const double* z, double* y) {
  return guard([&] {
    o->stage();
    cudaStream_t s = o->op->stream;
    o->x.upload(x, o->op->ndofs, s);
    o->z.upload(z, o->op->ndofs, s);
    o->r.upload(y, o->op->ndofs, s);
    o->op->jacobian_apply(t, wM, wA, o->x.p, o->z.p, o->r.p);
    o->r.download(y, s);
  });
}
int dcb_block_diagonal(dcb_operator* o, double t, double wM, double wA, const double* x, double* bdiag, int64_t cap) {
  return guard([&] {
    o->stage();
    cudaStream_t s = o->op->stream;
    if (cap < o->op->bdiag_size()) fail("dcb_block_diagonal: buffer too small, need ", o->op->bdiag_size());
    DeviceBuffer<double> b(o->op->bdiag_size());
    b.zero(s);
    o->x.upload(x, o->op->ndofs, s);
    o->op->block_diag(t, wM, wA, o->x.p, b.p);
    b.download(bdiag, s);
  });
}
int dcb_operator_uses_tiles(const dcb_operator* o) { return o->op->tile_ready() ? 1 : 0; }
int dcb_residual_dev(dcb_operator* o, double t, double wM, double wA, const double* x, double* r) {
  return guard([&] { o->op->residual(t, wM, wA, x, r); });
}
int dcb_jacobian_dev(dcb_operator* o, double t, double wM, double wA, const double* x, double* vals) {
  return guard([&] { o->op->jacobian_csr(t, wM, wA, x, vals); });
}
int dcb_jacobian_apply_dev(dcb_operator* o, double t, double wM, double wA, const double* x, const double* z, double* y) {
  return guard([&] { o->op->jacobian_apply(t, wM, wA, x, z, y); });
}

// ---- solver
dcb_solver* dcb_solver_create(dcb_operator* o, const dcb_config* cfg, dcb_comm* comm) {
  return guard_new<dcb_solver>([&] {
    auto* s = new dcb_solver();
    s->op = o;
    s->s = std::make_unique<LinearSolver>(o->op, cfg->tree, comm ? comm->c.get() : nullptr);
    s->x.alloc(o->op->ndofs); s->b.alloc(o->op->ndofs); s->z.alloc(o->op->ndofs);
    return s;
  });
}
void dcb_solver_destroy(dcb_solver* s) { delete s; }
int dcb_solver_linearize(dcb_solver* s, double t, double wM, double wA, const double* x) {
  return guard([&] {
    s->x.upload(x, s->op->op->ndofs, s->op->op->stream);
    s->s->linearize(t, wM, wA, s->x.p);
  });
}
int dcb_solver_solve(dcb_solver* s, const double* b, double* z, double rel_tol, dcb_solve_result* out) {
  return guard([&] {
    cudaStream_t st = s->op->op->stream;
    s->b.upload(b, s->op->op->ndofs, st);
    SolveResult r = s->s->apply(s->b.p, s->z.p, rel_tol);
    s->z.download(z, st);
    if (out) {
      out->iterations = r.iterations; out->half_iterations = r.half_iterations;
      out->converged = r.converged; out->reduction = r.reduction; out->defect0 = r.defect0;
    }
  });
}
int dcb_solver_is_fused(const dcb_solver* s) { return s->s->is_fused() ? 1 : 0; }
int dcb_solver_apply_operator(dcb_solver* s, const double* v, double* y) {
  return guard([&] {
    cudaStream_t st = s->op->op->stream;
    s->b.upload(v, s->op->op->ndofs, st);
    s->s->apply_operator(s->b.p, s->z.p);
    s->z.download(y, st);
  });
}

// ---- stepper
dcb_stepper* dcb_stepper_create(dcb_operator* o, const dcb_config* cfg, dcb_comm* comm) {
  return guard_new<dcb_stepper>([&] {
    auto* s = new dcb_stepper();
    s->op = o;
    s->s = std::make_unique<StepOperator>(o->op, cfg->tree.sub("model.time_step_operator"), comm ? comm->c.get() : nullptr);
    s->u.alloc(o->op->ndofs);
    s->u.zero(o->op->stream);
    return s;
  });
}
void dcb_stepper_destroy(dcb_stepper* s) { delete s; }
int dcb_stepper_set_state(dcb_stepper* s, const double* u, double time) {
  return guard([&] {
    s->u.upload(u, s->op->op->ndofs, s->op->op->stream);
    DCB_CUDA(cudaStreamSynchronize(s->op->op->stream));
    s->time = time;
  });
}
int dcb_stepper_get_state(dcb_stepper* s, double* u, double* time) {
  return guard([&] {
    if (u) s->u.download(u, s->op->op->stream);
    if (time) *time = s->time;
  });
}
double* dcb_stepper_state_dev(dcb_stepper* s) { return s->u.p; }
int dcb_stepper_set_time(dcb_stepper* s, double time) { s->time = time; return 0; }
int dcb_stepper_step(dcb_stepper* s, double dt, int* ok) {
  return guard([&] {
    bool good = s->s->step(s->u.p, s->time, dt);
    DCB_CUDA(cudaStreamSynchronize(s->op->op->stream));
    if (good) s->time += dt;
    if (ok) *ok = good;
  });
}
int dcb_stepper_evolve(dcb_stepper* s, double t_end, double* dt, int max_steps, int* accepted) {
  return guard([&] {
    int n = s->s->evolve(s->u.p, &s->time, t_end, dt, max_steps);
    DCB_CUDA(cudaStreamSynchronize(s->op->op->stream));
    if (accepted) *accepted = n;
  });
}
int dcb_stepper_stats(const dcb_stepper* s, dcb_step_stats* o) {
  const StepStats& st = s->s->stats;
  o->steps = st.steps; o->failed_steps = st.failed_steps; o->stages = st.stages;
  o->newton_iterations = st.newton_iterations; o->linear_solves = st.linear_solves;
  o->linear_iterations = st.linear_iterations; o->linear_half_iterations = st.linear_half_iterations;
  o->residual_evaluations = st.residual_evaluations; o->linearizations = st.linearizations;
  o->kernel_launches = s->op->op->stats.launches;
  return 0;
}

// ---- output
int dcb_grid_write_vtk(dcb_grid* g, const dcb_model* m, const double* u_host, double time, const char* path, int append) {
  return guard([&] { write_vtk(*g->g, *m->m, u_host, time, path, append != 0, g->vtk_timesteps[path]); });
}

// ---- reduce
dcb_reducer* dcb_reducer_create(dcb_operator* o, const dcb_config* cfg, dcb_comm* comm) {
  return guard_new<dcb_reducer>([&] {
    auto* r = new dcb_reducer();
    r->op = o;
    r->r = std::make_unique<Reducer>(o->op, cfg->tree, comm ? comm->c.get() : nullptr);
    return r;
  });
}
void dcb_reducer_destroy(dcb_reducer* r) { delete r; }
int dcb_reducer_num_keys(const dcb_reducer* r) { return r->r->size(); }
const char* dcb_reducer_key(const dcb_reducer* r, int k) {
  return k >= 0 && k < r->r->size() ? r->r->key(k).c_str() : nullptr;
}
static int reducer_apply(dcb_reducer* r, double time, const double* x_dev, double* values, int32_t* status) {
  int rc = 0;
  int g = guard([&] {
    auto out = r->r->apply(time, x_dev, false);
    std::string msg;
    for (size_t k = 0; k < out.size(); ++k) {
      if (values) values[k] = out[k].value;
      if (status) status[k] = out[k].status;
      if (out[k].status == 2) rc = 2;
    }
    if (rc == 2) g_error = r->r->last_error;   // the reference's ReductionError text (reduce.hh:241-247)
  });
  return g ? 1 : rc;
}
int dcb_reducer_apply_dev(dcb_reducer* r, double time, const double* x_dev, double* values, int32_t* status) {
  return reducer_apply(r, time, x_dev, values, status);
}
int dcb_reducer_apply(dcb_reducer* r, double time, const double* x_host, double* values, int32_t* status) {
  int g = guard([&] {
    r->x.upload(x_host, r->op->op->ndofs, r->op->op->stream);
  });
  return g ? 1 : reducer_apply(r, time, r->x.p, values, status);
}
int dcb_model_precompile_reduce(dcb_model* m, const dcb_config* cfg) {
  return guard([&] { Reducer::precompile(*m->m, cfg->tree); });
}

// ---- multi GPU
int dcb_nccl_unique_id(char id[128]) { return guard([&] { nccl_unique_id(id); }); }
dcb_grid* dcb_grid_partition(const dcb_grid* global, int rank, int size) {
  return dcb_grid_partition_method(global, rank, size, "auto");
}
dcb_grid* dcb_grid_partition_method(const dcb_grid* global, int rank, int size, const char* method) {
  return guard_new<dcb_grid>([&] {
    auto* g = new dcb_grid();
    g->g = std::make_shared<Grid>(global->g->partition(rank, size, method ? method : "auto"));
    return g;
  });
}
int64_t dcb_grid_num_owned_vertices(const dcb_grid* g) { return g->g->n_owned; }
int dcb_grid_owned_vertex_range(const dcb_grid* g, int64_t* begin, int64_t* end) {
  *begin = g->g->n_owned < 0 ? 0 : g->g->owned_begin;
  *end = g->g->n_owned < 0 ? g->g->nv : g->g->owned_begin + g->g->n_owned;
  return 0;
}
int dcb_grid_get_global_vertex_ids(const dcb_grid* g, int64_t* gids) {
  std::memcpy(gids, g->g->global_vid.data(), g->g->global_vid.size() * 8);
  return 0;
}
int dcb_grid_get_vertex_owner(const dcb_grid* g, int32_t* owner) {
  std::memcpy(owner, g->g->vowner.data(), g->g->vowner.size() * 4);
  return 0;
}
int dcb_grid_get_global_element_ids(const dcb_grid* g, int64_t* eids) {
  std::memcpy(eids, g->g->global_eid.data(), g->g->global_eid.size() * 8);
  return 0;
}
int dcb_grid_halo_num_peers(const dcb_grid* g, int rank) {
  std::vector<int> peers; std::vector<std::vector<int32_t>> s, r;
  g->g->halo_plan(rank, peers, s, r);
  return (int)peers.size();
}
int dcb_grid_halo_peer(const dcb_grid* g, int rank, int k, int32_t* peer, int64_t* nsend, int64_t* nrecv) {
  std::vector<int> peers; std::vector<std::vector<int32_t>> s, r;
  g->g->halo_plan(rank, peers, s, r);
  if (k < 0 || k >= (int)peers.size()) return 1;
  *peer = peers[k]; *nsend = (int64_t)s[k].size(); *nrecv = (int64_t)r[k].size();
  return 0;
}
int dcb_grid_halo_lists(const dcb_grid* g, int rank, int k, int32_t* send, int32_t* recv) {
  std::vector<int> peers; std::vector<std::vector<int32_t>> s, r;
  g->g->halo_plan(rank, peers, s, r);
  if (k < 0 || k >= (int)peers.size()) return 1;
  std::memcpy(send, s[k].data(), s[k].size() * 4);
  std::memcpy(recv, r[k].data(), r[k].size() * 4);
  return 0;
}
dcb_comm* dcb_comm_create(const char id[128], int rank, int size, dcb_operator* o) {
  return guard_new<dcb_comm>([&] {
    HaloPlan plan;
    o->op->grid->halo_plan(rank, plan.peers, plan.send_idx, plan.recv_idx);
    auto* c = new dcb_comm();
    c->c.reset(nccl_communicator_create(id, rank, size, plan));
    // reductions run over the owned dofs only
    std::vector<int64_t> b, e;
    o->op->grid->owned_ranges(b, e);
    if (b.size() > 8) fail("more than 8 compartments per rank are not supported in multi-GPU runs");
    o->op->owned.n = (int)b.size();
    for (size_t k = 0; k < b.size(); ++k) { o->op->owned.b[k] = b[k]; o->op->owned.e[k] = e[k]; }
    return c;
  });
}
void dcb_comm_destroy(dcb_comm* c) { delete c; }
int dcb_comm_uses_peer_memory(const dcb_comm* c) { return c && c->c && c->c->peer_active ? 1 : 0; }
int dcb_operator_owned_ranges(const dcb_operator* o, int64_t* begin, int64_t* end, int cap) {
  std::vector<int64_t> b, e;
  o->op->grid->owned_ranges(b, e);
  for (int k = 0; k < (int)b.size() && k < cap; ++k) { begin[k] = b[k]; end[k] = e[k]; }
  return (int)b.size();
}

}  // extern "C"
