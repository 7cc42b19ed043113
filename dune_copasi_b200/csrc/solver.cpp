#include "solver.hpp"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "comm.hpp"

namespace dcb {

LinearSolver::LinearSolver(std::shared_ptr<DeviceOperator> op, const PTree& cfg, Communicator* comm)
    : op_(std::move(op)), comm_(comm) {
  // defaults: the reference picks UMFPack (direct) when SuiteSparse exists, else BiCGSTAB
  // (factory/inverse.hh:11-15); the direct solvers are not data parallel and not built here.
  type = cfg.get("type", std::string("BiCGSTAB"));
  if (type != "BiCGSTAB" && type != "CG" && type != "RestartedGMRes")
    fail("linear_solver.type = '", type, "' is not built for the B200 path (available: BiCGSTAB, CG, RestartedGMRes)");
  restart = cfg.get("restart", 40);   // iterative.hh:64
  if (type == "RestartedGMRes" && (restart < 1 || restart > 500)) fail("linear_solver.restart = ", restart, " is out of range [1, 500]");
  // DUNE_COPASI_DEFAULT_PRECONDITIONER is SSOR (solver/istl/factory/preconditioner.hh:17): an ini without
  // preconditioner.type gets the reference's choice; where SSOR cannot run here (matrix free, partitioned
  // grids: checked below) the call fails loudly instead of solving with something else
  prec_type = cfg.get("preconditioner.type", std::string("SSOR"));
  if (prec_type != "Jacobi" && prec_type != "BlockJacobi" && prec_type != "Richardson" && !sor_family())
    fail("linear_solver.preconditioner.type = '", prec_type,
         "' is not built for the B200 path (available: Richardson, Jacobi, BlockJacobi, SSOR, SOR, GaussSeidel)");
  prec_iterations = cfg.get("preconditioner.iterations", 1);
  if (prec_iterations < 1) fail("linear_solver.preconditioner.iterations must be >= 1");
  relaxation = cfg.get("preconditioner.relaxation", 1.0);
  matrix_free = cfg.get("matrix_free", false);
  verbosity = cfg.get("verbosity", 0);
  speculation = cfg.get("b200.speculation", true);
  // measured on B200 (tools/bench_ms.py, tools/bench_ssor3d.py): the self-scheduled sweep loses to the level loop
  // (mitchell_schaefer 128^2: 706 vs 380 ms per step, Gray-Scott 64^3: 2888 vs 502) -- a few hundred thousand
  // spinning threads polling the L2 starve the chain of rows that can run; off by default, the level loop runs as a graph
  sor_sweep_ = cfg.get("b200.sor_sweep", false);
  sor_graph_ = cfg.get("b200.sor_graph", true);
  auto range = cfg.get_vec("convergence_condition.iteration_range", {1, 500});   // iterative.hh:53-54
  max_iterations = (int)range.back();
  la::reduce_workspace_create(&ws_);
  const bool gmres = type == "RestartedGMRes";
  scal_.alloc(gmres ? std::max(32, restart + 2) : 32);
  hscal_.alloc(gmres ? std::max(32, restart + 2) : 32);
  const int nwork = type == "BiCGSTAB" ? 6 : 2;
  for (int k = 0; k < nwork; ++k) work_[k].alloc(op_->ndofs);
  if (type == "BiCGSTAB") {
    xalt_.alloc(op_->ndofs);   // alternate iterate (speculative half steps, see apply)
    for (auto& e : ev_) DCB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  if (prec_iterations > 1 && prec_type != "Richardson" && !sor_family())
    for (auto& w : sweep_) w.alloc(op_->ndofs);
  if (gmres) basis_.alloc((int64_t)(restart + 1) * op_->ndofs);   // Krylov basis v_0 .. v_m
  if (!matrix_free) {
    op_->ensure_csr();
    vals.alloc(op_->nnz());
  }
  if (sor_family()) {
    const char* how = cfg.has_key("preconditioner.type") ? "" : " (the reference's default when preconditioner.type is not set)";
    if (matrix_free) fail("linear_solver.preconditioner.type = ", prec_type, how, " sweeps over the assembled matrix: set linear_solver.matrix_free = false or choose Jacobi / BlockJacobi");
    if (comm_ && comm_->size > 1) fail("linear_solver.preconditioner.type = ", prec_type, how, " is a sequential sweep in dof order and is not built for partitioned (multi-GPU) runs: choose Jacobi / BlockJacobi");
    build_levels();
  }
  // off by default: measured on 2 x B200 (256^3) the two extra launches for the halo layers cost
  // what the hidden exchange saves (61.1 vs 60.9 ms per step)
  overlap_halo_ = comm_ && comm_->size > 1 && matrix_free && op_->can_split_apply() && cfg.get("b200.overlap_halo", false);
  if (overlap_halo_) {
    int lo = 0, hi = 0;
    DCB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    DCB_CUDA(cudaStreamCreateWithPriority(&halo_stream_, cudaStreamNonBlocking, hi));
    for (auto& e : halo_ev_) DCB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  if (prec_type == "Jacobi") dinv_.alloc(op_->ndofs);
  // BiCGSTAB with the vector updates and dot products fused into the tile-marching apply kernels
  // (kernels/assembly_tile.cuh): matrix free, Jacobi folded, no Dirichlet rows
  fused_ = type == "BiCGSTAB" && matrix_free && prec_type == "Jacobi" && prec_iterations == 1 && op_->tile_ready() &&
           op_->ncons == 0 && !overlap_halo_ && cfg.get("b200.fused", true);
  if (fused_) valt_.alloc(op_->ndofs);
  // BiCGSTAB without the preconditioned vectors y = w D^-1 p, y2 = w D^-1 r: the structured apply forms them while it
  // loads its corners (fp64-bound kernel, the extra loads hit L1) and the closing sweep recomputes them from p, r and
  // D^-1 -- 21 instead of 25 vector passes per iteration, the same products in the same order (identical iterates)
  yfree_ = type == "BiCGSTAB" && matrix_free && prec_type == "Jacobi" && prec_iterations == 1 && !fused_ && !overlap_halo_ &&
           op_->apply_scale_ready() && cfg.get("b200.yfree", op_->grid->elem_kind == 0);
  // (on by default for P1 only: same box, 256^3: P1 94.8 -> 93.7 ms per step; Q1 55.9 -> 58.2 -- the Q1 apply is bound by
  // its loads, not by the fp64 pipe, and pays more for the extra corner loads than the sweeps save)
  if (prec_type == "BlockJacobi" || (matrix_free && prec_type == "Jacobi")) bdiag_.alloc(op_->bdiag_size());
}

LinearSolver::~LinearSolver() {
  for (auto& g : sor_graphs_) cudaGraphExecDestroy(g.second.exec);
  la::reduce_workspace_destroy(&ws_);
  for (auto& e : ev_)
    if (e) cudaEventDestroy(e);
  for (auto& e : halo_ev_)
    if (e) cudaEventDestroy(e);
  if (halo_stream_) cudaStreamDestroy(halo_stream_);
}

void LinearSolver::linearize(double t, double wM, double wA, const double* x) {
  t_ = t; wM_ = wM; wA_ = wA; x_ = x;
  cudaStream_t s = op_->stream;
  const Grid& g = *op_->grid;
  const int ncomp = op_->model->ncomp();
  if (!matrix_free) {
    vals.zero(s);
    op_->jacobian_csr(t, wM, wA, x, vals.p);
    if (op_->ncons) { la::csr_constrain(op_->ndofs, op_->rowptr.p, op_->colidx.p, vals.p, op_->cmask.p, s); op_->stats.launches++; }
    if (prec_type == "Jacobi") { la::csr_extract_diag_inv(op_->ndofs, op_->rowptr.p, op_->colidx.p, vals.p, dinv_.p, s); op_->stats.launches++; }
    if (prec_type == "BlockJacobi")
      for (int c = 0; c < ncomp; ++c) {
        int bs = op_->model->comp_nspec[c];
        if (bs == 0) continue;
        int64_t nb = (g.comp_offset[c + 1] - g.comp_offset[c]) / bs;
        la::csr_extract_block_diag(g.comp_offset[c], nb, bs, op_->rowptr.p, op_->colidx.p, vals.p, bdiag_.p + op_->bdiag_shift(c), s);
        op_->stats.launches++;
      }
  } else if (prec_type == "Jacobi" && (dinv_.zero(s), op_->scalar_diag(t, wM, wA, x, dinv_.p))) {
    // scalar diagonal assembled directly (no ns x ns blocks), then inverted in place
    la::invert_diag(op_->ndofs, dinv_.p, op_->cmask.p, s);
    op_->stats.launches++;
  } else if (prec_type != "Richardson") {
    bdiag_.zero(s);
    op_->block_diag(t, wM, wA, x, bdiag_.p);
    for (int c = 0; c < ncomp; ++c) {
      int bs = op_->model->comp_nspec[c];
      if (bs == 0) continue;
      int64_t nb = (g.comp_offset[c + 1] - g.comp_offset[c]) / bs;
      if (op_->ncons) { la::bdiag_constrain(g.comp_offset[c], nb, bs, bdiag_.p + op_->bdiag_shift(c), op_->cmask.p, s); op_->stats.launches++; }
      if (prec_type == "Jacobi") { la::block_diag_to_dinv(g.comp_offset[c], nb, bs, bdiag_.p + op_->bdiag_shift(c), dinv_.p, s); op_->stats.launches++; }
    }
  }
  // the fused sweeps form D^-1 p at ghost vertices themselves: the owners' diagonal entries are needed there
  if ((fused_ || yfree_) && comm_) comm_->halo_update(dinv_.p, s);
  if (yfree_ && relaxation != 1.0) {   // relax * dinv, rounded once: what the sweeps multiply p and r with
    if (wdinv_.n < (size_t)op_->ndofs) wdinv_.alloc(op_->ndofs);
    la::copy(op_->ndofs, dinv_.p, wdinv_.p, s);
    la::scale(op_->ndofs, relaxation, wdinv_.p, s);
    op_->stats.launches += 2;
  }
  if (prec_type == "BlockJacobi")
    for (int c = 0; c < ncomp; ++c) {
      int bs = op_->model->comp_nspec[c];
      if (bs == 0) continue;
      int64_t nb = (g.comp_offset[c + 1] - g.comp_offset[c]) / bs;
      // blocks of compartment c start at bdiag[comp_offset[c] * bs]
      la::block_invert(nb, bs, bdiag_.p + op_->bdiag_shift(c) + g.comp_offset[c] * bs, s);
      op_->stats.launches++;
    }
}

// pushed: the ghost planes of v are already on their way (the sweep that wrote v pushed them, link_push):
// only the pull is left.  zeroed: y is zero on entry (cleared by the sweep that consumed it last).
void LinearSolver::apply_operator(const double* v, double* y, bool pushed, bool zeroed, bool scaled) {
  cudaStream_t s = op_->stream;
  struct ScaleGuard {   // scaled: the operator reads w D^-1 v instead of v (yfree_)
    DeviceOperator* op;
    ~ScaleGuard() { if (op) op->set_apply_scale(nullptr, 1.0); }
  } guard{scaled ? op_.get() : nullptr};
  if (scaled) op_->set_apply_scale(relaxation == 1.0 ? dinv_.p : wdinv_.p, 1.0);
  if (comm_ && overlap_halo_) {
    // structured slabs, matrix free: the ghost planes of v travel on a second (high priority)
    // stream while the cells that only read owned vertices are integrated; the two cell layers next
    // to the ghost planes follow once the planes have arrived
    DCB_CUDA(cudaEventRecord(halo_ev_[0], s));
    DCB_CUDA(cudaStreamWaitEvent(halo_stream_, halo_ev_[0], 0));
    comm_->halo_update(const_cast<double*>(v), halo_stream_);
    DCB_CUDA(cudaEventRecord(halo_ev_[1], halo_stream_));
    la::fill(op_->ndofs, 0.0, y, s);
    op_->stats.launches++;
    op_->jacobian_apply(t_, wM_, wA_, x_, v, y, 1);
    DCB_CUDA(cudaStreamWaitEvent(s, halo_ev_[1], 0));
    op_->jacobian_apply(t_, wM_, wA_, x_, v, y, 2);
    if (op_->ncons) { la::copy_values(op_->ncons, op_->cdofs.p, v, y, s); op_->stats.launches++; }
    return;
  }
  if (comm_ && pushed) comm_->halo_pull(const_cast<double*>(v), s);
  else if (comm_) comm_->halo_update(const_cast<double*>(v), s);
  if (!matrix_free) {
    int avg = (int)(op_->nnz() / std::max<int64_t>(1, op_->ndofs));
    DeviceOperator::ProfScope ps(op_.get(), "spmv");
    la::spmv_csr(op_->ndofs, op_->rowptr.p, op_->rowptr32.p, op_->colidx.p, vals.p, v, y, avg, s);
    op_->stats.launches++;
  } else if (op_->tile_aligned({x_, v, y})) {
    // owner-computes sweep: y is written, not accumulated; identity rows included
    op_->tile_apply(t_, wM_, wA_, x_, v, y);
  } else {
    if (!zeroed) {
      la::fill(op_->ndofs, 0.0, y, s);   // MatrixFreeAdapter::apply zeroes y first (make_step_operator.hh:70-75)
      op_->stats.launches++;
    }
    op_->jacobian_apply(t_, wM_, wA_, x_, v, y);
    if (op_->ncons) { la::copy_values(op_->ncons, op_->cdofs.p, v, y, s); op_->stats.launches++; }   // identity rows
  }
}

// v = 0, then `prec_iterations` sweeps (preconditioner.iterations):
//   Jacobi (dune-istl SeqJac):  v += w D^-1 (d - A v), old iterate in every row;
//   BlockJacobi (block_jacobi.hh:102-127 as written): the copy of the right-hand side is modified
//   cumulatively, b_k = b_{k-1} - A v_{k-1} -- the true defect for the first two sweeps only.
// Level sets of the sweeps: level(j) = 1 + max level(i) over the rows i < j coupled with j in either
// direction (a_ji != 0: j reads the new value of i; a_ij != 0: i must read the old value of j -- the
// pattern need not be structurally symmetric, e.g. a species that reacts to another one-way).  Two
// rows of one level are then never coupled and every coupled row with a larger index sits in a later
// level, so ascending levels reproduce dune-istl's ascending-row sweep and descending levels its
// descending-row sweep exactly, whatever the order inside a level.
void LinearSolver::build_levels() {
  const auto& rp = op_->h_rowptr;
  const auto& ci = op_->h_colidx;
  const int64_t n = op_->ndofs;
  std::vector<int32_t> level(n, 0);
  int32_t nlev = 0;
  for (int64_t i = 0; i < n; ++i) {
    int32_t l = level[i];   // constraints pushed by smaller rows that read this one
    for (int64_t k = rp[i]; k < rp[i + 1]; ++k)
      if (ci[k] < i) l = std::max(l, level[ci[k]] + 1);
    level[i] = l;
    for (int64_t k = rp[i]; k < rp[i + 1]; ++k)
      if (ci[k] > i) level[ci[k]] = std::max(level[ci[k]], l + 1);
    nlev = std::max(nlev, l + 1);
  }
  level_ptr_.assign(nlev + 1, 0);
  for (int64_t i = 0; i < n; ++i) level_ptr_[level[i] + 1]++;
  for (int32_t l = 0; l < nlev; ++l) level_ptr_[l + 1] += level_ptr_[l];
  std::vector<int32_t> rows(n);
  std::vector<int64_t> cur(level_ptr_.begin(), level_ptr_.end() - 1);
  for (int64_t i = 0; i < n; ++i) rows[cur[level[i]]++] = (int32_t)i;
  level_rows_.upload(rows, op_->stream);
  if (sor_sweep_) {
    // slots of the self-scheduled sweeps: levels padded to whole warps
    std::vector<int32_t> slots;
    slots.reserve((size_t)n + 32 * (size_t)nlev);
    for (int32_t l = 0; l < nlev; ++l) {
      for (int64_t p = level_ptr_[l]; p < level_ptr_[l + 1]; ++p) slots.push_back(rows[p]);
      while (slots.size() % 32) slots.push_back(-1);
    }
    sweep_nslots_ = (int64_t)slots.size();
    sweep_slots_.upload(slots, op_->stream);
    // rows to wait for: the pattern of A + A^T.  Structurally symmetric patterns (the usual case) reuse A's.
    bool symmetric = true;
#pragma omp parallel for schedule(static) reduction(&& : symmetric)
    for (int64_t i = 0; i < n; ++i)
      for (int64_t k = rp[i]; k < rp[i + 1] && symmetric; ++k) {
        const int32_t j = ci[k];
        symmetric = symmetric && std::binary_search(ci.begin() + rp[j], ci.begin() + rp[j + 1], (int32_t)i);
      }
    dep_is_pattern_ = symmetric;
    if (!symmetric) {
      std::vector<std::vector<int32_t>> extra(n);   // transposed entries missing from the row
      for (int64_t i = 0; i < n; ++i)
        for (int64_t k = rp[i]; k < rp[i + 1]; ++k) {
          const int32_t j = ci[k];
          if (!std::binary_search(ci.begin() + rp[j], ci.begin() + rp[j + 1], (int32_t)i)) extra[j].push_back((int32_t)i);
        }
      std::vector<int64_t> dp(n + 1, 0);
      std::vector<int32_t> di;
      for (int64_t i = 0; i < n; ++i) {
        di.insert(di.end(), ci.begin() + rp[i], ci.begin() + rp[i + 1]);
        di.insert(di.end(), extra[i].begin(), extra[i].end());
        dp[i + 1] = (int64_t)di.size();
      }
      dep_ptr_.upload(dp, op_->stream);
      dep_idx_.upload(di, op_->stream);
    }
    sweep_done_.alloc(n);
    sweep_done_.zero(op_->stream);
    sweep_epoch_ = 0;
  }
  DCB_CUDA(cudaStreamSynchronize(op_->stream));
}

// v = 0; iterations x (forward sweep [; backward sweep])   -- SeqSSOR / SeqSOR / SeqGS::apply
void LinearSolver::sor_apply(const double* d, double* v) {
  cudaStream_t s = op_->stream;
  DeviceOperator::ProfScope ps(op_.get(), "precond");
  const int64_t* rp = (const int64_t*)op_->rowptr.p;
  const int nlev = (int)level_ptr_.size() - 1;
  const bool skip_diag = prec_type == "GaussSeidel";
  if (skip_diag && sweep_[0].n < (size_t)op_->ndofs) sweep_[0].alloc(op_->ndofs);
  // The application is a chain of 2 nlev + 1 dependent launches with fixed arguments per (d, v) pair: captured once
  // into a CUDA graph and replayed (linear_solver.b200.sor_graph, default on) -- the launches of a level loop are
  // issue-bound (2-5 us each on the stream, ~1 us as graph nodes).
  const auto key = std::make_pair((const void*)d, (void*)v);
  long long& L = op_->stats.launches;
  if (sor_graph_ && !sor_sweep_) {
    auto it = sor_graphs_.find(key);
    if (it != sor_graphs_.end()) {
      DCB_CUDA(cudaGraphLaunch(it->second.exec, s));
      L += it->second.launches;
      return;
    }
    if (sor_graphs_.size() >= 128) {   // GMRES basis vectors come back after a restart; anything beyond that is a leak
      for (auto& g : sor_graphs_) cudaGraphExecDestroy(g.second.exec);
      sor_graphs_.clear();
    }
    DCB_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
  }
  const long long L0 = L;
  la::fill(op_->ndofs, 0.0, v, s);
  L++;
  auto sweep = [&](bool backward, bool skip) {
    if (sor_sweep_) {
      if (sweep_epoch_ == INT32_MAX) { sweep_done_.zero(s); sweep_epoch_ = 0; }
      la::sor_sweep(sweep_slots_.p, sweep_nslots_, backward, rp, op_->colidx.p, vals.p, dep_is_pattern_ ? rp : dep_ptr_.p,
                    dep_is_pattern_ ? op_->colidx.p : dep_idx_.p, d, v, relaxation, skip, sweep_done_.p, ++sweep_epoch_, s);
      op_->stats.launches++;
      return;
    }
    for (int k = 0; k < nlev; ++k) {
      const int l = backward ? nlev - 1 - k : k;
      la::sor_level(level_rows_.p + level_ptr_[l], level_ptr_[l + 1] - level_ptr_[l], rp, op_->colidx.p, vals.p, d, v,
                    relaxation, skip, s);
    }
    op_->stats.launches += nlev;
  };
  for (int it = 0; it < prec_iterations; ++it) {
    if (skip_diag) { la::copy(op_->ndofs, v, sweep_[0].p, s); op_->stats.launches++; }   // dbgs: xold
    sweep(false, skip_diag);
    if (skip_diag) { la::relax_blend(op_->ndofs, relaxation, sweep_[0].p, v, s); op_->stats.launches++; }
    if (prec_type == "SSOR") sweep(true, false);
  }
  if (sor_graph_ && !sor_sweep_) {
    cudaGraph_t graph = nullptr;
    DCB_CUDA(cudaStreamEndCapture(s, &graph));
    SorGraph g;
    g.launches = L - L0;
    DCB_CUDA(cudaGraphInstantiate(&g.exec, graph, 0));
    cudaGraphDestroy(graph);
    sor_graphs_[key] = g;
    DCB_CUDA(cudaGraphLaunch(g.exec, s));   // the capture recorded the work, this runs it
  }
}

void LinearSolver::precondition(const double* d, double* v) {
  if (sor_family()) { sor_apply(d, v); return; }
  precondition_sweep(d, v);
  if (prec_iterations <= 1 || prec_type == "Richardson") return;
  cudaStream_t s = op_->stream;
  const int64_t n = op_->ndofs;
  double *b = sweep_[0].p, *t = sweep_[1].p, *c = sweep_[2].p;
  auto& L = op_->stats.launches;
  if (prec_type == "BlockJacobi") { la::copy(n, d, b, s); L++; }
  for (int it = 1; it < prec_iterations; ++it) {
    apply_operator(v, t);
    { DeviceOperator::ProfScope ps(op_.get(), "blas1"); la::sub(n, prec_type == "Jacobi" ? d : b, t, b, s); L++; }
    precondition_sweep(b, c);
    { DeviceOperator::ProfScope ps(op_.get(), "blas1"); la::axpy(n, 1.0, c, v, s); L++; }
  }
}

void LinearSolver::precondition_sweep(const double* d, double* v) {
  cudaStream_t s = op_->stream;
  DeviceOperator::ProfScope ps(op_.get(), "precond");
  const Grid& g = *op_->grid;
  if (prec_type == "Richardson") {
    la::copy(op_->ndofs, d, v, s);
    if (relaxation != 1.0) { la::fill(op_->ndofs, 0.0, v, s); la::axpy(op_->ndofs, relaxation, d, v, s); }
  } else if (prec_type == "Jacobi") {
    // SeqJac, one sweep from v = 0: v = w D^-1 d
    la::jacobi_apply(op_->ndofs, dinv_.p, relaxation, d, v, s);
  } else {
    // BlockJacobi::apply with iterations = 1 from v = 0 (block_jacobi.hh:102-127): v = w Dblk^-1 d
    for (int c = 0; c < op_->model->ncomp(); ++c) {
      int bs = op_->model->comp_nspec[c];
      if (bs == 0) continue;
      int64_t nb = (g.comp_offset[c + 1] - g.comp_offset[c]) / bs;
      la::block_jacobi_apply(g.comp_offset[c], nb, bs, bdiag_.p + op_->bdiag_shift(c), relaxation, d, v, s);
    }
  }
  op_->stats.launches++;
}

// all-reduce scal_[first, first+count) and bring scal_[0, total) to the host
void LinearSolver::fetch_slots(int first, int count, int total) {
  cudaStream_t s = op_->stream;
  if (comm_) comm_->allreduce_sum(scal_.p + first, count, s);
  DCB_CUDA(cudaMemcpyAsync(hscal_.p, scal_.p, sizeof(double) * total, cudaMemcpyDeviceToHost, s));
  DCB_CUDA(cudaStreamSynchronize(s));
}

void LinearSolver::fetch(int n) {
  cudaStream_t s = op_->stream;
  if (comm_) comm_->allreduce_sum(scal_.p, n, s);
  DCB_CUDA(cudaMemcpyAsync(hscal_.p, scal_.p, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
  DCB_CUDA(cudaStreamSynchronize(s));
}

// dune-istl BiCGSTABSolver::apply with everything between two operator applications fused into the
// apply kernels (kernels/assembly_tile.cuh), same operation order as the unfused loop in apply():
//   A: K1  p' = r + beta (p - omega v); v' = A (w D^-1 p');          <rt, v'>
//      K2  r~ = r - alpha v';           t  = A (w D^-1 r~);          |r~|^2, <t, r~>, <t, t>
//   B: K3  x' = x + alpha w D^-1 p' + omega w D^-1 r~;  r = r~ - omega t;   |r|^2, <rt, r>
// p, v and r are double buffered (a neighbouring tile may still read the old value of a vertex whose
// owner has already written the new one).  The defect norm of the first half step reaches the host
// one apply later than in the unfused loop, i.e. K2 of the last iteration may run in vain.
// Device scalars: s[2] = <rt,v>, s[3] = |r~|^2, s[4..5] = (<t,r~>, <t,t>), pair k at s[8+2k] as in apply().
SolveResult LinearSolver::apply_bicgstab_fused(double* b, double* x, double rel_tol) {
  cudaStream_t s = op_->stream;
  const int64_t n = op_->ndofs;
  const la::Ranges& own = op_->owned;
  SolveResult res;
  auto& L = op_->stats.launches;
  double *r = b, *rt = work_[0].p, *t = work_[3].p, *ralt = work_[4].p;
  double* pbuf[2] = {work_[1].p, work_[5].p};
  double* vbuf[2] = {work_[2].p, valt_.p};
  int pb = 0, vb = 0;
  double* sc = scal_.p;
  auto pair = [&](int it_index) { return sc + 8 + 2 * (it_index & 1); };
  { DeviceOperator::ProfScope ps(op_.get(), "blas1"); la::fill(n, 0.0, x, s); L++; }
  if (comm_) comm_->halo_update(r, s);   // the sweeps update ghost entries locally from here on
  { DeviceOperator::ProfScope ps(op_.get(), "blas1"); la::copy(n, r, rt, s); L++; }
  { DeviceOperator::ProfScope ps(op_.get(), "blas1"); la::dot(own, r, r, pair(-1) + 1, ws_, s); L++; }
  if (comm_) comm_->allreduce_sum(pair(-1) + 1, 1, s);
  DCB_CUDA(cudaMemcpyAsync(hscal_.p, pair(-1) + 1, sizeof(double), cudaMemcpyDeviceToHost, s));
  { DeviceOperator::HostTimer ht(op_.get(), "host_wait"); DCB_CUDA(cudaStreamSynchronize(s)); }
  double norm0 = std::sqrt(hscal_.p[0]), norm = norm0;
  res.defect0 = norm0;
  if (!(norm0 == norm0)) { res.converged = false; return res; }
  if (norm0 < 1e-30) { res.converged = true; res.reduction = 0; return res; }
  double rho = 1, alpha = 1, omega = 1, rho_new = hscal_.p[0];
  double *x_cur = x, *x_alt = xalt_.p;
  double* hA = hscal_.p + 16;   // [4 * parity]: <rt,v>, |r~|^2, <t,r~>, <t,t>
  double* hB = hscal_.p + 24;   // [2 * parity]: |r|^2, <rt,r>
  auto stage_a = [&](int i) {
    DeviceOperator::TileFused f;
    f.pro = 1; f.epi = 1; f.first = i == 0; f.relax = relaxation;
    f.r_in = r; f.p_in = pbuf[pb]; f.v_in = vbuf[vb]; f.dinv = dinv_.p; f.w = rt; f.p_out = pbuf[pb ^ 1];
    f.rho_new = pair(i - 1) + 1; f.rho = pair(i - 2) + 1; f.hptr = sc + 2; f.trtt = sc + 4;
    f.out = sc + 1; f.out_mask = 2;
    op_->tile_apply(t_, wM_, wA_, x_, nullptr, vbuf[vb ^ 1], &f);
    pb ^= 1; vb ^= 1;
    if (comm_) { comm_->allreduce_sum(sc + 2, 1, s); comm_->halo_update(vbuf[vb], s); }
    DeviceOperator::TileFused g;
    g.pro = 2; g.epi = 2; g.relax = relaxation;
    g.r_in = r; g.v_in = vbuf[vb]; g.dinv = dinv_.p; g.r_out = ralt;
    g.rho = pair(i - 1) + 1; g.hptr = sc + 2;
    g.out = sc + 3; g.out_mask = 7;
    op_->tile_apply(t_, wM_, wA_, x_, nullptr, t, &g);
    if (comm_) { comm_->allreduce_sum(sc + 3, 3, s); comm_->halo_update(t, s); }
    DCB_CUDA(cudaMemcpyAsync(hA + 4 * (i & 1), sc + 2, sizeof(double) * 4, cudaMemcpyDeviceToHost, s));
    DCB_CUDA(cudaEventRecord(ev_[0], s));
  };
  auto stage_b = [&](int i, const double* xin, double* xout) {
    { DeviceOperator::ProfScope ps(op_.get(), "blas1");
      la::bicg_final_fold(n, own, pair(i - 1) + 1, sc + 2, sc + 4, dinv_.p, relaxation, pbuf[pb], ralt, xin, xout, t, r, rt,
                          pair(i), ws_, s); L++; }
    if (comm_) comm_->allreduce_sum(pair(i), 2, s);
    DCB_CUDA(cudaMemcpyAsync(hB + 2 * (i & 1), pair(i), sizeof(double) * 2, cudaMemcpyDeviceToHost, s));
    DCB_CUDA(cudaEventRecord(ev_[1], s));
  };
  auto speculate = [&](double known_norm) { return speculation && known_norm >= 100.0 * rel_tol * norm0; };
  double it = 0.5;
  bool pending = false;
  int i = 0;
  bool queued = false;
  for (; it < max_iterations; it += 0.5, ++i) {
    if (std::fabs(rho) <= 1e-80 || std::fabs(omega) <= 1e-80) break;   // breakdown (SolverAbort)
    if (!queued) stage_a(i);
    queued = false;
    const bool ahead = speculate(norm);
    if (ahead) stage_b(i, x_cur, x_alt);                  // speculative
    { DeviceOperator::HostTimer ht(op_.get(), "host_wait"); DCB_CUDA(cudaEventSynchronize(ev_[0])); }
    pending = true;
    const double* a4 = hA + 4 * (i & 1);
    const double h = a4[0];
    alpha = rho_new / h;
    norm = std::sqrt(a4[1]);
    res.half_iterations++;
    if (std::fabs(h) < 1e-80 || !(norm == norm)) break;
    if (norm < rel_tol * norm0 || norm < 1e-30) { res.converged = true; break; }
    it += 0.5;
    if (!ahead) stage_b(i, x_cur, x_alt);
    if (it + 0.5 < max_iterations && speculate(norm)) { stage_a(i + 1); queued = true; }   // speculative
    { DeviceOperator::HostTimer ht(op_.get(), "host_wait"); DCB_CUDA(cudaEventSynchronize(ev_[1])); }
    pending = false;
    std::swap(x_cur, x_alt);
    const double* b2 = hB + 2 * (i & 1);
    omega = a4[2] / a4[3];
    rho = rho_new;
    rho_new = b2[1];
    norm = std::sqrt(b2[0]);
    res.half_iterations++;
    if (!(norm == norm)) break;
    if (norm < rel_tol * norm0 || norm < 1e-30) { res.converged = true; break; }
  }
  if (pending) {
    // x += alpha w D^-1 p of the first half step (the speculative second half, if any, wrote x_alt only)
    DeviceOperator::ProfScope ps(op_.get(), "blas1");
    la::bicg_x_half(n, pair(i - 1) + 1, sc + 2, dinv_.p, relaxation, pbuf[pb], x_cur, s); L++;
  }
  if (x_cur != x) { DeviceOperator::ProfScope ps(op_.get(), "blas1"); la::copy(n, x_cur, x, s); L++; }
  res.iterations = (int)std::ceil(std::min<double>(it, max_iterations));
  res.reduction = norm / norm0;
  if (comm_) comm_->halo_update(x, s);
  return res;
}

SolveResult LinearSolver::apply(double* b, double* x, double rel_tol) {
  SolveResult res = fused_ ? apply_bicgstab_fused(b, x, rel_tol) : apply_krylov(b, x, rel_tol);
  // the flag spins of the peer-memory collectives are bounded: a solve during which a peer never answered is
  // reported as what it is, whatever its sums happened to look like
  if (comm_ && comm_->peer_error())
    fail("a peer-memory collective gave up waiting for another rank (rank ", comm_->rank, " of ", comm_->size, ")");
  return res;
}

SolveResult LinearSolver::apply_krylov(double* b, double* x, double rel_tol) {
  cudaStream_t s = op_->stream;
  const int64_t n = op_->ndofs;
  const la::Ranges& own = op_->owned;
  SolveResult res;
  double* r = b;
  auto& L = op_->stats.launches;
  // x is the zero vector on entry (make_step_operator.hh:230, Newton's correction): A*0 = 0 is
  // skipped, r = b.
  { DeviceOperator::ProfScope ps(op_.get(), "blas1"); la::fill(n, 0.0, x, s); L++; }
  if (type == "BiCGSTAB") {
    // dune-istl BiCGSTABSolver::apply, run one half step ahead of the host: every step length is
    // formed on the device (kernels/linalg.cu), so the host only needs the defect norm to decide on
    // convergence.  While it waits for the norm of half step k, half step k+1 is already queued;
    // if k turns out to be the last one, the speculative work is dropped: the first half step never
    // touches x and the second writes its iterate into the alternate buffer.
    // Device scalars: s[0] = |r|^2 after the first half, s[2] = <rt,v>, s[4..5] = (<t,r>, <t,t>),
    // pair k at s[8+2k] = (|r|^2 after the second half, <rt,r>) of the iterations with parity k.
    double *rt = work_[0].p, *p = work_[1].p, *v = work_[2].p, *t = work_[3].p, *y = work_[4].p, *y2 = work_[5].p;
    double* sc = scal_.p;
    auto pair = [&](int it_index) { return sc + 8 + 2 * (it_index & 1); };   // it_index may be -1
    // Jacobi is folded into the sweeps that produce its argument; other preconditioners run on their own
    const double* fold = prec_type == "Jacobi" && prec_iterations == 1 ? dinv_.p : nullptr;
    { DeviceOperator::ProfScope ps(op_.get(), "blas1"); la::copy(n, r, rt, s); L++; }
    // Collectives and host traffic fused into the sweeps (kernels/linalg.cu, peer_device.cuh):
    //   * the last block of every reducing kernel all-reduces its sums over the peer mailboxes and mirrors them
    //     into mapped host memory (no k_allreduce launch, no device-to-host copy between two sweeps);
    //   * the sweeps that write the operator's input (y, y2; Jacobi folded) push the halo planes to the
    //     neighbours while they stream, the operator application only pulls them into the ghost range;
    //   * the sweeps that consume the operator's result clear it for the next application (no fill launch).
    const bool rlink = !comm_ || comm_->reduce_links_ready();              // sums reach hscal_ straight from the kernels
    const bool plink = comm_ && fold && comm_->push_links_ready() && !overlap_halo_;
    const bool zfuse = matrix_free && fold && !overlap_halo_ && !op_->tile_aligned({x_, y, v});
    const bool yfree = yfree_ && zfuse;
    const double* fold_store = yfree ? nullptr : fold;   // sweeps store w D^-1 (.) only when somebody reads it
    auto link = [&](bool reduce, double* host_out, bool push, bool zero) {
      la::Link k;
      if (comm_ && reduce && rlink) comm_->link_reduce(&k.peer);
      if (push && plink) comm_->link_push(&k.peer);
      k.host_out = rlink ? host_out : nullptr;
      k.zero_input = zero && zfuse;
      return k;
    };
    auto reduce_after = [&](double* dev, int count) { if (comm_ && !rlink) comm_->allreduce_sum(dev, count, s); };
    // <rt,r> = <r,r> at the start: stored where iteration 0 looks for its rho
    { DeviceOperator::ProfScope ps(op_.get(), "blas1");
      la::dot(own, r, r, pair(-1) + 1, ws_, s, link(true, hscal_.p, false, false)); L++; }
    reduce_after(pair(-1) + 1, 1);
    if (!rlink) DCB_CUDA(cudaMemcpyAsync(hscal_.p, pair(-1) + 1, sizeof(double), cudaMemcpyDeviceToHost, s));
    if (zfuse) { DeviceOperator::ProfScope ps(op_.get(), "blas1"); la::fill(n, 0.0, v, s); la::fill(n, 0.0, t, s); L += 2; }
    { DeviceOperator::HostTimer ht(op_.get(), "host_wait"); DCB_CUDA(cudaStreamSynchronize(s)); }
    double norm0 = std::sqrt(hscal_.p[0]), norm = norm0;
    res.defect0 = norm0;
    if (!(norm0 == norm0)) { res.converged = false; return res; }
    if (norm0 < 1e-30) { res.converged = true; res.reduction = 0; return res; }
    double rho = 1, alpha = 1, omega = 1, rho_new = hscal_.p[0];
    double *x_cur = x, *x_alt = xalt_.p;
    auto first_half = [&](int i) {
      { DeviceOperator::ProfScope ps(op_.get(), "blas1");
        la::bicg_p_prec(n, p, r, v, pair(i - 1) + 1, pair(i - 2) + 1, sc + 2, sc + 4, i == 0, fold_store, relaxation, y, ws_, s,
                        link(false, nullptr, true, true)); L++; }
      if (!fold) precondition(p, y);
      apply_operator(yfree ? p : y, v, plink, zfuse, yfree);
      { DeviceOperator::ProfScope ps(op_.get(), "blas1");
        la::dot(own, rt, v, sc + 2, ws_, s, link(true, hscal_.p + 2, false, false)); L++; }
      reduce_after(sc + 2, 1);
      { DeviceOperator::ProfScope ps(op_.get(), "blas1");
        la::bicg_r_prec(n, own, pair(i - 1) + 1, sc + 2, v, r, fold_store, relaxation, y2, sc, ws_, s,
                        link(true, hscal_.p, true, false)); L++; }
      reduce_after(sc, 1);
      if (!rlink) DCB_CUDA(cudaMemcpyAsync(hscal_.p, sc, sizeof(double) * 3, cudaMemcpyDeviceToHost, s));
      DCB_CUDA(cudaEventRecord(ev_[0], s));
    };
    auto second_half = [&](int i, const double* xin, double* xout) {
      if (!fold) precondition(r, y2);
      apply_operator(yfree ? r : y2, t, plink, zfuse, yfree);
      { DeviceOperator::ProfScope ps(op_.get(), "blas1");
        la::dot2(own, t, r, t, t, sc + 4, ws_, s, link(true, hscal_.p + 4, false, false)); L++; }
      reduce_after(sc + 4, 2);
      { DeviceOperator::ProfScope ps(op_.get(), "blas1");
        if (yfree)
          la::bicg_final_fold(n, own, pair(i - 1) + 1, sc + 2, sc + 4, dinv_.p, relaxation, p, r, xin, xout, t, r, rt, pair(i),
                              ws_, s, link(true, hscal_.p + 8 + 2 * (i & 1), false, true));
        else
          la::bicg_final(n, own, pair(i - 1) + 1, sc + 2, sc + 4, y, y2, xin, xout, t, r, rt, pair(i), ws_, s,
                         link(true, hscal_.p + 8 + 2 * (i & 1), false, true));
        L++; }
      reduce_after(pair(i), 2);
      if (!rlink) DCB_CUDA(cudaMemcpyAsync(hscal_.p + 4, sc + 4, sizeof(double) * 8, cudaMemcpyDeviceToHost, s));
      DCB_CUDA(cudaEventRecord(ev_[1], s));
    };
    // work ahead only while the last known defect is two orders above the target: the half steps
    // right before convergence are the ones whose speculative successor would be thrown away
    auto speculate = [&](double known_norm) { return speculation && known_norm >= 100.0 * rel_tol * norm0; };
    double it = 0.5;
    bool pending = false;   // x += alpha*y of the first half step is applied together with the second
    int i = 0;
    bool queued = false;    // first half of iteration i already enqueued
    for (; it < max_iterations; it += 0.5, ++i) {
      if (std::fabs(rho) <= 1e-80 || std::fabs(omega) <= 1e-80) break;   // breakdown (SolverAbort)
      if (!queued) first_half(i);
      queued = false;
      const bool ahead = speculate(norm);
      if (ahead) second_half(i, x_cur, x_alt);            // speculative
      { DeviceOperator::HostTimer ht(op_.get(), "host_wait"); DCB_CUDA(cudaEventSynchronize(ev_[0])); }
      pending = true;
      double h = hscal_.p[2];
      alpha = rho_new / h;
      norm = std::sqrt(hscal_.p[0]);
      res.half_iterations++;
      if (std::fabs(h) < 1e-80 || !(norm == norm)) break;
      if (norm < rel_tol * norm0 || norm < 1e-30) { res.converged = true; break; }
      it += 0.5;
      if (!ahead) second_half(i, x_cur, x_alt);
      if (it + 0.5 < max_iterations && speculate(norm)) { first_half(i + 1); queued = true; }   // speculative
      { DeviceOperator::HostTimer ht(op_.get(), "host_wait"); DCB_CUDA(cudaEventSynchronize(ev_[1])); }
      pending = false;
      std::swap(x_cur, x_alt);                            // the iterate of this iteration
      const double* pr = hscal_.p + 8 + 2 * (i & 1);
      omega = hscal_.p[4] / hscal_.p[5];
      rho = rho_new;
      rho_new = pr[1];
      norm = std::sqrt(pr[0]);
      res.half_iterations++;
      if (!(norm == norm)) break;
      if (norm < rel_tol * norm0 || norm < 1e-30) { res.converged = true; break; }
    }
    if (pending) {
      DeviceOperator::ProfScope ps(op_.get(), "blas1");
      if (yfree) la::bicg_x_half(n, pair(i - 1) + 1, sc + 2, dinv_.p, relaxation, p, x_cur, s);   // alpha = rho / h as on the host
      else la::axpy(n, alpha, y, x_cur, s);
      L++;
    }
    if (x_cur != x) { DeviceOperator::ProfScope ps(op_.get(), "blas1"); la::copy(n, x_cur, x, s); L++; }
    res.iterations = (int)std::ceil(std::min<double>(it, max_iterations));
    res.reduction = norm / norm0;
  } else if (type == "RestartedGMRes") {
    // dune-istl RestartedGMResSolver: left preconditioned GMRES(m), modified Gram-Schmidt, Givens
    // rotations on the host; convergence on the preconditioned defect.  The Gram-Schmidt
    // coefficients stay on the device (slot k = <v_k,w>, slot i+1 = <w,w>) and are read back once
    // per iteration for the Hessenberg column.
    const int m = restart;
    auto V = [&](int k) { return basis_.p + (int64_t)k * n; };
    double *w = work_[0].p, *tmp = work_[1].p;
    std::vector<double> H((size_t)(m + 1) * m, 0.0), sv(m + 1, 0.0), cs(m, 0.0), sn(m, 0.0), yv(m, 0.0);
    auto Hh = [&](int rr, int cc) -> double& { return H[(size_t)rr * m + cc]; };
    auto givens_apply = [](double& dx, double& dy, double c, double sgn) {
      double tt = c * dx + sgn * dy;
      dy = -sgn * dx + c * dy;
      dx = tt;
    };
    precondition(b, V(0));
    { DeviceOperator::ProfScope ps(op_.get(), "blas1"); la::dot(own, V(0), V(0), scal_.p, ws_, s); L++; }
    fetch(1);
    double norm0 = std::sqrt(hscal_.p[0]), norm = norm0;
    res.defect0 = norm0;
    if (!(norm0 == norm0)) { res.converged = false; return res; }
    if (norm0 < 1e-30) { res.converged = true; res.reduction = 0; return res; }
    int j = 1;
    while (j <= max_iterations && !res.converged) {
      int i = 0;
      { DeviceOperator::ProfScope ps(op_.get(), "blas1"); la::scale(n, 1.0 / norm, V(0), s); L++; }
      sv[0] = norm;
      for (int k = 1; k < m + 1; ++k) sv[k] = 0.0;
      for (i = 0; i < m && j <= max_iterations && !res.converged; ++i, ++j) {
        apply_operator(V(i), V(i + 1));
        precondition(V(i + 1), w);
        // modified Gram-Schmidt, dune-istl's order (h_k = <v_k, w>; w -= h_k v_k), one pass per basis vector:
        // step k subtracts the projection found by step k - 1 and forms the next product on the updated w;
        // the last step closes with <w, w>.  Sums are all-reduced by the kernels' last block (peer links) and
        // mirrored into hscal_.
        const bool rlink = !comm_ || comm_->reduce_links_ready();
        for (int k = 0; k <= i + 1; ++k) {
          DeviceOperator::ProfScope ps(op_.get(), "blas1");
          la::Link lk;
          if (comm_ && rlink) comm_->link_reduce(&lk.peer);
          if (rlink) lk.host_out = hscal_.p + k;
          la::mgs_step(n, own, k ? scal_.p + k - 1 : nullptr, k ? V(k - 1) : nullptr, w, k <= i ? V(k) : nullptr, scal_.p + k, ws_, s, lk);
          if (comm_ && !rlink) comm_->allreduce_sum(scal_.p + k, 1, s);
          L++;
        }
        { DeviceOperator::ProfScope ps(op_.get(), "blas1"); la::normalize_dev(n, w, scal_.p + i + 1, V(i + 1), s); L++; }
        if (!rlink) DCB_CUDA(cudaMemcpyAsync(hscal_.p, scal_.p, sizeof(double) * (i + 2), cudaMemcpyDeviceToHost, s));
        DCB_CUDA(cudaStreamSynchronize(s));
        for (int k = 0; k < i + 1; ++k) Hh(k, i) = hscal_.p[k];
        double hn = std::sqrt(hscal_.p[i + 1]);
        Hh(i + 1, i) = hn;
        res.half_iterations += 2;
        if (!(hn == hn) || std::fabs(hn) < 1e-80) { j = max_iterations + 1; break; }   // breakdown (SolverAbort)
        for (int k = 0; k < i; ++k) givens_apply(Hh(k, i), Hh(k + 1, i), cs[k], sn[k]);
        {
          double dx = Hh(i, i), dy = Hh(i + 1, i);
          if (std::fabs(dy) < 1e-300) { cs[i] = 1.0; sn[i] = 0.0; }
          else if (std::fabs(dx) < 1e-300) { cs[i] = 0.0; sn[i] = 1.0; }
          else { double nrm = std::sqrt(dx * dx + dy * dy); cs[i] = dx / nrm; sn[i] = dy / nrm; }
        }
        givens_apply(Hh(i, i), Hh(i + 1, i), cs[i], sn[i]);
        givens_apply(sv[i], sv[i + 1], cs[i], sn[i]);
        norm = std::fabs(sv[i + 1]);
        if (norm < rel_tol * norm0 || norm < 1e-30) res.converged = true;
      }
      // x += sum_k y_k v_k with H y = s (upper triangular after the rotations)
      for (int a = i - 1; a >= 0; --a) {
        double acc = sv[a];
        for (int c2 = a + 1; c2 < i; ++c2) acc -= Hh(a, c2) * yv[c2];
        yv[a] = acc / Hh(a, a);
      }
      for (int a = 0; a < i; ++a) { DeviceOperator::ProfScope ps(op_.get(), "blas1"); la::axpy(n, yv[a], V(a), x, s); L++; }
      if (!res.converged && j <= max_iterations) {
        // restart from the true preconditioned defect (dune-istl: the same condition as the outer loop)
        apply_operator(x, tmp);
        { DeviceOperator::ProfScope ps(op_.get(), "blas1"); la::sub(n, b, tmp, w, s); L++; }
        precondition(w, V(0));
        { DeviceOperator::ProfScope ps(op_.get(), "blas1"); la::dot(own, V(0), V(0), scal_.p, ws_, s); L++; }
        fetch(1);
        norm = std::sqrt(hscal_.p[0]);
      }
    }
    res.iterations = j - 1;
    res.reduction = norm / norm0;
  } else {   // CG
    double *p = work_[0].p, *q = work_[1].p;
    { DeviceOperator::ProfScope ps(op_.get(), "blas1"); la::dot(own, r, r, scal_.p, ws_, s); L++; }
    fetch(1);
    double norm0 = std::sqrt(hscal_.p[0]), norm = norm0;
    res.defect0 = norm0;
    if (norm0 < 1e-30) { res.converged = true; res.reduction = 0; return res; }
    precondition(r, p);
    { DeviceOperator::ProfScope ps(op_.get(), "blas1"); la::dot(own, p, r, scal_.p, ws_, s); L++; }
    fetch(1);
    double rholast = hscal_.p[0];
    int i = 1;
    for (; i <= max_iterations; ++i) {
      apply_operator(p, q);
      { DeviceOperator::ProfScope ps(op_.get(), "blas1"); la::dot(own, p, q, scal_.p, ws_, s); L++; }
      fetch(1);
      double lambda = rholast / hscal_.p[0];
      { DeviceOperator::ProfScope ps(op_.get(), "blas1"); la::axpy_pair_norm(n, own, lambda, p, x, q, r, nullptr, scal_.p, ws_, s); L++; }
      fetch(1);
      norm = std::sqrt(hscal_.p[0]);
      res.half_iterations += 2;
      if (norm < rel_tol * norm0 || norm < 1e-30) { res.converged = true; break; }
      precondition(r, q);
      { DeviceOperator::ProfScope ps(op_.get(), "blas1"); la::dot(own, q, r, scal_.p, ws_, s); L++; }
      fetch(1);
      double rho = hscal_.p[0], beta = rho / rholast;
      { DeviceOperator::ProfScope ps(op_.get(), "blas1"); la::xpby(n, p, q, beta, s); L++; }
      rholast = rho;
    }
    res.iterations = std::min(i, max_iterations);
    res.reduction = norm / norm0;
  }
  if (comm_) comm_->halo_update(x, s);
  return res;
}

}  // namespace dcb
