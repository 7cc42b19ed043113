/* dune_copasi_b200 -- C ABI of the B200-native CG-P1 diffusion-reaction hot path.
 *
 * Drop-in boundary.  The reference exposes this path to dune-pdelab as a
 * *local* operator (dune/copasi/model/diffusion_reaction/local_operator.hh) that PDELab's grid
 * operator calls once per element.  The per-element call is exactly the overhead removed here, so
 * the replacement hooks one level up, at PDELab::Operator granularity, on raw arrays:
 *
 *   dcb_residual           <- Operator::apply(x, r), additive        make_step_operator.hh:223
 *                             = sum over elements of localAssembleVolume       local_operator.hh:417-491
 *                               + localAssembleSkeleton / Boundary             :777-971, :1398-1415
 *   dcb_jacobian           <- Operator::derivative(x) "container"    make_step_operator.hh:233, 109-111
 *                             = localAssembleJacobianVolume / Skeleton / Boundary  :541-707, :973-1199, :1417-1436
 *   dcb_jacobian_apply     <- matrix-free derivative apply           make_step_operator.hh:62-95
 *                             = localAssembleJacobianVolumeApply / SkeletonApply   :510-524, :1354-1396, :1438-1452
 *   dcb_grid_pattern       <- basisToPattern + patternToMatrix       make_step_operator.hh:378-384
 *                             = localAssemblePattern{Volume,Skeleton,Boundary}     :276-399
 *   dcb_solver_*           <- LinearSolver::apply                    make_step_operator.hh:102-146
 *                             (dune-istl BiCGSTAB/CG/RestartedGMRes + Jacobi/BlockJacobi, solver/istl/)
 *   dcb_stepper_*          <- PDELab::OneStep (RungeKutta o Newton)  make_step_operator.hh:408-443,
 *                             SimpleAdaptiveStepper                  common/stepper.hh:337-368
 *   dcb_reducer_*          <- DiffusionReaction::reduce              diffusion_reaction/reduce.hh:38-285
 *
 * wM / wA are the Runge-Kutta weights of the mass form (Form::Mass, storage terms) and of the
 * stiffness form (Form::Stiffness, reaction + diffusion + outflow), local_operator.hh:143-147:
 * every operator call evaluates wM*M(u) + wA*A(t,u).
 *
 * Conventions: plain pointers and sizes, no exceptions across the boundary.  Functions returning
 * int return 0 on success; constructors return NULL on failure; dcb_last_error() holds the message
 * (thread local).  Host-pointer entry points copy to/from the device inside the call; *_dev entry
 * points take device pointers and are ordered on the operator's stream.  The caller owns host
 * buffers, the library owns device buffers.  There is no CPU fallback: anything that computes
 * fails with an error when no CUDA device is present.
 */
#ifndef DUNE_COPASI_B200_H
#define DUNE_COPASI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dcb_config dcb_config;
typedef struct dcb_grid dcb_grid;
typedef struct dcb_model dcb_model;
typedef struct dcb_operator dcb_operator;
typedef struct dcb_solver dcb_solver;
typedef struct dcb_stepper dcb_stepper;
typedef struct dcb_reducer dcb_reducer;
typedef struct dcb_comm dcb_comm;

typedef struct {
  int32_t iterations;       /* ceil(it), as dune-istl's InverseOperatorResult */
  int32_t half_iterations;
  int32_t converged;
  double reduction;
  double defect0;
} dcb_solve_result;

typedef struct {
  int64_t steps, failed_steps, stages, newton_iterations, linear_solves, linear_iterations,
      linear_half_iterations, residual_evaluations, linearizations, kernel_launches;
} dcb_step_stats;

/* ---- library ---- */
int dcb_version(void);
const char* dcb_last_error(void);
int dcb_device_count(void);

/* ---- configuration: Dune::ParameterTree semantics (INI text + key=value overrides,
 *      src/dune_copasi.cc:270-282) ---- */
dcb_config* dcb_config_create(void);
void dcb_config_destroy(dcb_config*);
int dcb_config_parse_ini(dcb_config*, const char* ini_text);
int dcb_config_set(dcb_config*, const char* key, const char* value);
/* returns the number of bytes needed (incl. NUL); copies at most cap bytes */
size_t dcb_config_dump(const dcb_config*, char* out, size_t cap);

/* ---- mesh (structure-of-arrays; simplices only, as the reference: grid/move_geometry.hh:48-49) */
dcb_grid* dcb_grid_create_structured(int dim, const int32_t* cells, const double* origin,
                                     const double* extent);
/* The lattice cells themselves as Q1 (multilinear) elements, 2-point Gauss rule per axis: BASELINE
 * configs[3]'s element.  NOT a reference capability (PkLocalFiniteElementMap is simplex-only,
 * model_single_compartment_traits.hh:23-24; grid/make_multi_domain_grid.hh:84-90 creates simplex
 * grids) -- same weak form (local_operator.hh:417-707), checked against the oracle's own Q1 element.
 * One compartment over the whole lattice, scalar point-independent diffusion, no facet terms. */
dcb_grid* dcb_grid_create_structured_cubes(int dim, const int32_t* cells, const double* origin,
                                           const double* extent);
dcb_grid* dcb_grid_create(int dim, int64_t nv, const double* coords, int64_t ne,
                          const int32_t* elems, int nkeys, const char* const* keys,
                          const double* cell_data);
void dcb_grid_destroy(dcb_grid*);
int dcb_grid_dim(const dcb_grid*);
int64_t dcb_grid_num_vertices(const dcb_grid*);
int64_t dcb_grid_num_elements(const dcb_grid*);
int dcb_grid_nodes_per_element(const dcb_grid*);   /* dim+1 (simplices) or 2^dim (Q1 cubes) */
int dcb_grid_get_coords(const dcb_grid*, double* coords);
int dcb_grid_get_elements(const dcb_grid*, int32_t* elems);

/* ---- model: [compartments], [parser_context], [model.*] of the ini ---- */
dcb_model* dcb_model_create(const dcb_config*, int dim, int nkeys, const char* const* keys);
void dcb_model_destroy(dcb_model*);
int dcb_model_num_compartments(const dcb_model*);
int dcb_model_num_species(const dcb_model*);
const char* dcb_model_species_name(const dcb_model*, int species);
int dcb_model_species_compartment(const dcb_model*, int species);
/* generated CUDA translation unit (model functions + kernels); owned by the model */
const char* dcb_model_cuda_source(dcb_model*);
/* the translation unit of one kernel group as the run-time compiler sees it (csrc/jit.hpp JitGroup:
 * 0 all, 1 patch, 2 element, 3 CSR fill, 4 facets, 5 structured, 6 structured Q1, 7 tile, 8 tile Q1) */
const char* dcb_model_cuda_source_group(dcb_model*, int group);
/* NVRTC-compile the model kernels for sm_100a (no GPU needed). kind 0: cubin, 1: PTX.
 * Returns the byte count and copies at most cap bytes; <0 on error. */
int64_t dcb_model_compile(dcb_model*, int kind, char* out, size_t cap);

/* compile every kernel group of the model into the on-disk JIT cache (no GPU needed) */
int dcb_model_precompile(dcb_model*);
int dcb_model_precompile_group(dcb_model*, int group);   /* one kernel group, numbered as above */

/* ---- binding a model to a mesh: compartments, facets, DOF map, sparsity pattern (host) ---- */
int dcb_grid_bind(dcb_grid*, const dcb_model*);
int64_t dcb_grid_num_dofs(const dcb_grid*);
int dcb_grid_get_elem_compartment(const dcb_grid*, int32_t* elem_comp);
int dcb_grid_get_elem_dof(const dcb_grid*, int64_t* elem_dof /* [ne*nodes_per_element] */);
int64_t dcb_grid_num_facets(const dcb_grid*);
int dcb_grid_get_facets(const dcb_grid*, int64_t* f_in, int64_t* f_out, int32_t* f_lin, int32_t* f_lout);
/* pattern: first call with NULL arrays to get the sizes */
int dcb_grid_pattern(dcb_grid*, const dcb_model*, int64_t* nrows, int64_t* nnz, int64_t* rowptr,
                     int32_t* colidx);
int dcb_grid_interpolate(const dcb_grid*, const dcb_model*, double time, double* u);
/* Model::write_vtk (diffusion_reaction/model_multi_compartment.impl.hh:218-300): per compartment
 * "<path>/<stem>-<compartment>-<00000>.vtu" with one vertex-data array per species, and
 * "<path>/<stem>-<compartment>.pvd"; append = 0 restarts the time sequence of `path` */
int dcb_grid_write_vtk(dcb_grid*, const dcb_model*, const double* u_host, double time, const char* path, int append);
int64_t dcb_grid_constraints(const dcb_grid*, const dcb_model*, int32_t* dofs, double* vals, int64_t cap);

/* ---- device operator ---- */
dcb_operator* dcb_operator_create(dcb_model*, dcb_grid*);
void dcb_operator_destroy(dcb_operator*);
int64_t dcb_operator_num_dofs(const dcb_operator*);
int64_t dcb_operator_nnz(dcb_operator*);
int64_t dcb_operator_launches(const dcb_operator*);
void* dcb_operator_stream(const dcb_operator*); /* cudaStream_t */
int dcb_operator_sync(dcb_operator*);
/* per-kernel-kind device timing (CUDA events on the operator's stream); the report is text,
 * one line per kind: "<kind> <accumulated ms> <launches>"; reading it clears the record */
int dcb_operator_profile(dcb_operator*, int enable);
size_t dcb_operator_profile_report(dcb_operator*, char* out, size_t cap);
/* host-buffer entry points (copies inside the call) */
int dcb_residual(dcb_operator*, double time, double wM, double wA, const double* x, double* r);
int dcb_jacobian(dcb_operator*, double time, double wM, double wA, const double* x, double* vals);
int dcb_jacobian_apply(dcb_operator*, double time, double wM, double wA, const double* x,
                       const double* z, double* y);
int dcb_block_diagonal(dcb_operator*, double time, double wM, double wA, const double* x, double* bdiag,
                       int64_t cap);
/* device-pointer entry points (accumulate into r / y / vals) */
/* 1 when residual / Jacobian apply run on the tile-marching (owner computes, atomic free) kernels (diagnostic) */
int dcb_operator_uses_tiles(const dcb_operator*);
int dcb_residual_dev(dcb_operator*, double time, double wM, double wA, const double* x, double* r);
int dcb_jacobian_dev(dcb_operator*, double time, double wM, double wA, const double* x, double* vals);
int dcb_jacobian_apply_dev(dcb_operator*, double time, double wM, double wA, const double* x,
                           const double* z, double* y);

/* ---- linear solver (config = the `linear_solver` sub-tree) ---- */
dcb_solver* dcb_solver_create(dcb_operator*, const dcb_config* linear_solver_cfg, dcb_comm*);
void dcb_solver_destroy(dcb_solver*);
int dcb_solver_linearize(dcb_solver*, double time, double wM, double wA, const double* x_host);
int dcb_solver_solve(dcb_solver*, const double* b_host, double* z_host, double rel_tol, dcb_solve_result*);
int dcb_solver_apply_operator(dcb_solver*, const double* v_host, double* y_host);
/* 1 when the solve runs as BiCGSTAB fused into the tile-marching apply kernels (diagnostic) */
int dcb_solver_is_fused(const dcb_solver*);

/* ---- time stepping (config = the whole ini; uses model.time_step_operator.*) ---- */
dcb_stepper* dcb_stepper_create(dcb_operator*, const dcb_config*, dcb_comm*);
void dcb_stepper_destroy(dcb_stepper*);
int dcb_stepper_set_state(dcb_stepper*, const double* u_host, double time);
int dcb_stepper_get_state(dcb_stepper*, double* u_host, double* time);
double* dcb_stepper_state_dev(dcb_stepper*);
int dcb_stepper_set_time(dcb_stepper*, double time);
/* one step of size dt; *ok = 0 if the step failed (state unchanged) */
int dcb_stepper_step(dcb_stepper*, double dt, int* ok);
/* adaptive evolution to t_end with at most max_steps accepted steps */
int dcb_stepper_evolve(dcb_stepper*, double t_end, double* dt, int max_steps, int* accepted);
int dcb_stepper_stats(const dcb_stepper*, dcb_step_stats*);

/* ---- [model.reduce] functionals (config = the whole ini): per key evaluation / reduction /
 *      transformation / error / warn expressions over an order-4 quadrature of every cell,
 *      dune/copasi/model/diffusion_reaction/reduce.hh:38-285.  values[k] is the transformed result of
 *      key k, status[k] = 0 fine, 1 warn.expression fired, 2 error.expression fired.  The apply calls
 *      return 2 when an error expression fired (dcb_last_error() holds the reference's
 *      ReductionError text), 1 on any other failure. ---- */
dcb_reducer* dcb_reducer_create(dcb_operator*, const dcb_config*, dcb_comm*);
void dcb_reducer_destroy(dcb_reducer*);
int dcb_reducer_num_keys(const dcb_reducer*);
const char* dcb_reducer_key(const dcb_reducer*, int k);
int dcb_reducer_apply(dcb_reducer*, double time, const double* x_host, double* values, int32_t* status);
int dcb_reducer_apply_dev(dcb_reducer*, double time, const double* x_dev, double* values, int32_t* status);
/* compile the reduce kernels of (model, config) into the on-disk JIT cache (no GPU needed) */
int dcb_model_precompile_reduce(dcb_model*, const dcb_config*);

/* ---- multi-GPU: one process per GPU, NCCL ---- */
int dcb_nccl_unique_id(char id[128]);
/* Vertex partition of a global mesh for `rank` of `size`: returns the local mesh (owned vertices first,
 * then one layer of ghosts, each group ascending in global id).  Maps are retrievable below.
 * method: "slab" (structured lattices: planes along the last axis), "rcb" (recursive coordinate bisection
 * of the vertices, any rank count), "range" (contiguous ranges of the vertex numbering), "auto" / NULL
 * (slab for structured lattices, rcb otherwise -- what dcb_grid_partition does). */
dcb_grid* dcb_grid_partition(const dcb_grid* global, int rank, int size);
dcb_grid* dcb_grid_partition_method(const dcb_grid* global, int rank, int size, const char* method);
int64_t dcb_grid_num_owned_vertices(const dcb_grid* local);
/* owned vertices are the local ids [begin, end): a prefix for the general partition, the middle
 * planes of the slab for structured grids */
int dcb_grid_owned_vertex_range(const dcb_grid* local, int64_t* begin, int64_t* end);
int dcb_grid_get_global_vertex_ids(const dcb_grid* local, int64_t* gids);
int dcb_grid_get_vertex_owner(const dcb_grid* local, int32_t* owner);
int dcb_grid_get_global_element_ids(const dcb_grid* local, int64_t* eids);
/* halo plan of a bound local grid (also what dcb_comm_create uses): peer k of `rank`, and the local
 * dof indices sent to / received from it, both ordered by (compartment, global vertex, species) */
int dcb_grid_halo_num_peers(const dcb_grid* local, int rank);
int dcb_grid_halo_peer(const dcb_grid* local, int rank, int k, int32_t* peer, int64_t* nsend, int64_t* nrecv);
int dcb_grid_halo_lists(const dcb_grid* local, int rank, int k, int32_t* send, int32_t* recv);
/* communicator for a bound local grid (halo plan derived from the partition) */
dcb_comm* dcb_comm_create(const char id[128], int rank, int size, dcb_operator* op);
void dcb_comm_destroy(dcb_comm*);
/* 1 when the small all-reduces (and slab halo updates) run over NVLink peer memory mapped with CUDA IPC
 * instead of NCCL (environment DCB_PEER_COLLECTIVES=0 forces NCCL) */
int dcb_comm_uses_peer_memory(const dcb_comm*);
/* owned local dof ranges [begin,end) per compartment */
int dcb_operator_owned_ranges(const dcb_operator*, int64_t* begin, int64_t* end, int cap);

#ifdef __cplusplus
}
#endif
#endif
