// Krylov solve on the device: hand-written BiCGSTAB, CG and restarted GMRES in dune-istl's operation
// order with Jacobi / block-Jacobi / Richardson preconditioning, on an assembled CSR Jacobian or
// matrix-free.
//
// Mirrors the reference's LinearSolver adapter (dune/copasi/model/make_step_operator.hh:55-157):
// configuration sub-tree `linear_solver.*`, solver + preconditioner (re)built on every apply
// (:120-127), relative tolerance handed in per call (:132-135), non-convergence reported as an
// error condition (:141-145).  Registry names follow solver/istl/factory/iterative.hh:87-106 and
// factory/preconditioner.hh:96-113; only the data-parallel subset is built (BiCGSTAB, CG,
// RestartedGMRes; Richardson, Jacobi, BlockJacobi), anything else fails loudly.
#pragma once
#include <vector>
#include <map>
#include <memory>
#include <string>

#include "operator.hpp"

namespace dcb {

struct SolveResult {
  int iterations = 0;        // ceil(it) as dune-istl reports it
  int half_iterations = 0;   // BiCGSTAB tests convergence every half iteration
  bool converged = false;
  double reduction = 1.0;
  double defect0 = 0.0;
};

// communicator hooks for multi-GPU runs (comm.cpp); null => single device
struct Communicator;

class LinearSolver {
 public:
  LinearSolver(std::shared_ptr<DeviceOperator> op, const PTree& cfg, Communicator* comm = nullptr);
  ~LinearSolver();

  // Linearisation J = wM dM/du + wA dA/du at (t, x).  Assembles the CSR values (matrix based)
  // or only the (block) diagonal (matrix free) and sets up the preconditioner.
  void linearize(double t, double wM, double wA, const double* x);
  // solve J z = b to the relative defect reduction `rel_tol`; b is consumed (holds the final defect)
  SolveResult apply(double* b, double* z, double rel_tol);
  // y = J v with the current linearisation
  void apply_operator(const double* v, double* y, bool pushed = false, bool zeroed = false, bool scaled = false);
  // BiCGSTAB with its vector updates and dot products fused into the tile-marching apply kernels
  bool is_fused() const { return fused_; }

  bool matrix_free = false;
  std::string type, prec_type;
  int max_iterations = 500;
  int restart = 40;              // RestartedGMRes: Krylov space dimension between restarts
  double relaxation = 1.0;
  int prec_iterations = 1;       // preconditioner.iterations (sweeps per application)
  int verbosity = 0;
  // linear_solver.b200.speculation: BiCGSTAB enqueues the next half step before the host has seen
  // the defect norm of the current one (results are identical either way)
  bool speculation = true;
  DeviceBuffer<double> vals;     // CSR values of the current linearisation (matrix based)

 private:
  void precondition(const double* d, double* v);
  void precondition_sweep(const double* d, double* v);   // one sweep from v = 0
  void fetch(int n);
  SolveResult apply_bicgstab_fused(double* b, double* z, double rel_tol);
  SolveResult apply_krylov(double* b, double* z, double rel_tol);
  bool fused_ = false;
  bool yfree_ = false;   // BiCGSTAB without the stored preconditioned vectors (solver.cpp)
  DeviceBuffer<double> wdinv_;   // relaxation * dinv_ (yfree_ with relaxation != 1)
  DeviceBuffer<double> valt_;
  void fetch_slots(int first, int count, int total);
  std::shared_ptr<DeviceOperator> op_;
  Communicator* comm_;
  la::ReduceWorkspace ws_;
  DeviceBuffer<double> scal_;               // device scalars of the reductions
  PinnedBuffer<double> hscal_;
  DeviceBuffer<double> dinv_, bdiag_, work_[6], basis_, sweep_[3], xalt_;
  cudaEvent_t ev_[2] = {nullptr, nullptr};   // first / second half step of BiCGSTAB reached the host buffers
  // linear_solver.b200.overlap_halo: halo exchange on its own stream under the interior cells
  bool overlap_halo_ = false;
  cudaStream_t halo_stream_ = nullptr;
  cudaEvent_t halo_ev_[2] = {nullptr, nullptr};
  // SSOR / SOR / GaussSeidel (dune-istl SeqSSOR / SeqSOR / SeqGS; SSOR is the reference's default
  // preconditioner, solver/istl/factory/preconditioner.hh:17,101-104): level-scheduled sweeps over the
  // assembled CSR.  level_ptr_[l] .. level_ptr_[l+1] index level_rows_, the rows whose lower
  // neighbours all sit in earlier levels.
  bool sor_family() const { return prec_type == "SSOR" || prec_type == "SOR" || prec_type == "GaussSeidel"; }
  void build_levels();
  void sor_apply(const double* d, double* v);
  std::vector<int64_t> level_ptr_;
  DeviceBuffer<int32_t> level_rows_;
  // self-scheduled sweeps (la::sor_sweep, linear_solver.b200.sor_sweep = true): one launch per sweep instead of one per level
  bool sor_sweep_ = false, sor_graph_ = true;
  struct SorGraph { cudaGraphExec_t exec = nullptr; long long launches = 0; };
  std::map<std::pair<const void*, void*>, SorGraph> sor_graphs_;   // one captured application per (d, v) pair
  DeviceBuffer<int32_t> sweep_slots_, dep_idx_;
  DeviceBuffer<int64_t> dep_ptr_;
  DeviceBuffer<int> sweep_done_;
  int64_t sweep_nslots_ = 0;
  int sweep_epoch_ = 0;
  bool dep_is_pattern_ = true;
  // linearisation point
  double t_ = 0, wM_ = 0, wA_ = 0;
  const double* x_ = nullptr;
};

}  // namespace dcb
