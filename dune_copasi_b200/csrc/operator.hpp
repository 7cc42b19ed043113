// Device-resident spatial operator: the instationary diffusion-reaction residual
//   r += wM * M(u) + wA * A(t,u)
// its Jacobian (assembled CSR, matrix-free apply, block diagonal) and the Dirichlet constraints.
// This is the global face of the reference's local operator (local_operator.hh) as PDELab's
// makeInstationaryMatrix{Based,Free}Assembler would drive it (make_step_operator.hh:291-293,
// 403-405): apply(x, r) is additive (cf. :223), derivative(x) yields either the "container"
// (CSR) or a matrix-free apply.
#pragma once
#include <chrono>
#include <initializer_list>
#include <map>
#include <memory>
#include <vector>

#include "grid.hpp"
#include "jit.hpp"
#include "kernels/kernel_args.h"
#include "kernels/linalg.hpp"
#include "model.hpp"

namespace dcb {

// dynamic shared memory of the tile-marching kernels (kernels/assembly_tile.cuh)
size_t tile_smem_bytes(int ns, int tile_w, int tile_r, int dim);

struct PatchSet {
  int comp = 0;
  int npatch = 0;
  int64_t elem_begin = 0;   // offset of this compartment in patch element order
  int64_t nelem = 0;
  int max_nodes = 0, max_elems = 0;   // budgets the patches of this compartment were cut with
  DeviceBuffer<int> node_ptr, nodes, elem_ptr, adj_ptr;
  DeviceBuffer<unsigned short> lconn, adj;
  // statistics (host)
  int64_t total_nodes = 0;
};

struct FacetList {   // one directional outflow pair (cs -> ct)
  int cs = 0, ct = 0;
  int64_t n = 0;
  DeviceBuffer<long long> f_self, f_other;
  DeviceBuffer<int> f_lself, f_lother;
};

struct OperatorStats {
  long long launches = 0;   // kernels launched by this operator (assembly + linear algebra)
};

class DeviceOperator {
 public:
  DeviceOperator(std::shared_ptr<const Model> model, std::shared_ptr<const Grid> grid);
  ~DeviceOperator();

  // all pointers are device pointers; everything is ordered on `stream`
  void residual(double t, double wM, double wA, const double* x, double* r);
  // part 0: all cells; 1 / 2 (only if can_split_apply()): the cells whose vertices are all owned /
  // the two cell layers along the slab axis that read ghost planes -- lets the caller overlap the
  // halo exchange of z with the interior cells
  void jacobian_apply(double t, double wM, double wA, const double* x, const double* z, double* y, int part = 0);
  bool can_split_apply() const;
  // Jacobi folded into the apply: the next jacobian_apply reads its direction as relax * dinv .* z (structured
  // per-cell driver only: apply_scale_ready()); null switches it off again
  bool apply_scale_ready() const;
  double struct_simplex_volume() const;
  void set_apply_scale(const double* dinv, double relax) { zscale_ = dinv; zrelax_ = relax; }
  void jacobian_csr(double t, double wM, double wA, const double* x, double* vals);
  void block_diag(double t, double wM, double wA, const double* x, double* bdiag);
  // scalar diagonal straight into a dof-indexed vector (structured scheme without facet terms);
  // returns false when unsupported -- callers then derive it from block_diag
  bool scalar_diag(double t, double wM, double wA, const double* x, double* diag);

  // ---- tile-marching drivers (kernels/assembly_tile.cuh): owner computes, plain stores, the result is
  // written (not accumulated), optional fused BiCGSTAB update in front and reductions behind.
  struct TileFused {
    int pro = 0, epi = 0, first = 0;
    double relax = 1.0;
    const double *r_in = nullptr, *p_in = nullptr, *v_in = nullptr, *dinv = nullptr, *w = nullptr;
    double *r_out = nullptr, *p_out = nullptr;
    const double *rho_new = nullptr, *rho = nullptr, *hptr = nullptr, *trtt = nullptr;
    double* out = nullptr;   // device scalars: out[q] = reduction q for every bit q of out_mask
    int out_mask = 0;
  };
  // structured single-compartment lattices without facet terms whose staged planes fit shared memory
  bool tile_ready() const { return tile_ok_; }
  // ... and whose vectors the kernels can move with 16-byte accesses (even species counts)
  bool tile_aligned(std::initializer_list<const void*> ptrs) const;
  // y = J(x) z  (fused: see TileFused; the direction is then formed from r_in / p_in / v_in / dinv)
  void tile_apply(double t, double wM, double wA, const double* x, const double* z, double* y, const TileFused* f = nullptr);
  // r += wM M(x) + wA A(t, x) without atomics
  void tile_residual(double t, double wM, double wA, const double* x, double* r);

  // sparsity pattern on the device (built on first use)
  void ensure_csr();
  int64_t nnz() const { return nnz_; }
  int64_t bdiag_size() const;   // doubles in the block diagonal
  // blocks of compartment c live at bdiag[bdiag_shift(c) + dof*ns_c + j] (dof = global row)
  int64_t bdiag_shift(int c) const;

  std::shared_ptr<const Model> model;
  std::shared_ptr<const Grid> grid;
  cudaStream_t stream = nullptr;
  int64_t ndofs = 0;
  std::string scheme;   // "patch" | "atomic"
  OperatorStats stats;

  // CSR pattern
  DeviceBuffer<int64_t> rowptr;
  DeviceBuffer<int32_t> rowptr32, colidx;
  std::vector<int64_t> h_rowptr;
  std::vector<int32_t> h_colidx;
  // constraints
  int64_t ncons = 0;
  DeviceBuffer<int32_t> cdofs;
  DeviceBuffer<double> cvals;
  DeviceBuffer<unsigned char> cmask;   // empty when there are no constraints
  std::vector<int32_t> h_cdofs;
  std::vector<double> h_cvals;
  // owned dof ranges (multi-GPU: set by the partition; default: everything)
  la::Ranges owned;

  // per-kernel-kind device timing with CUDA events on `stream` (off by default)
  void profile_enable(bool on);
  void prof_begin(const char* kind);
  void prof_end();
  // synchronises; kind -> (accumulated ms, launches); clears the record
  std::map<std::string, std::pair<double, long long>> profile_collect();
  // host-side time (ms) next to the device events: how long the host waited in synchronisations
  // ("host_wait") against the whole step ("host_step") tells whether the device ever starves
  void host_add(const char* kind, double ms) {
    if (!profiling_) return;
    host_prof_[kind].first += ms;
    host_prof_[kind].second += 1;
  }
  struct HostTimer {
    DeviceOperator* op;
    const char* kind;
    std::chrono::steady_clock::time_point t0;
    HostTimer(DeviceOperator* o, const char* k) : op(o), kind(k), t0(std::chrono::steady_clock::now()) {}
    ~HostTimer() {
      op->host_add(kind, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
  };
  struct ProfScope {
    DeviceOperator* op;
    ProfScope(DeviceOperator* o, const char* kind) : op(o) { op->prof_begin(kind); }
    ~ProfScope() { op->prof_end(); }
  };

 private:
  friend class Reducer;   // [model.reduce] functionals read the device-resident mesh
  struct ProfRec { std::string kind; cudaEvent_t a = nullptr, b = nullptr; };
  bool profiling_ = false;
  std::vector<ProfRec> prof_;
  std::vector<cudaEvent_t> prof_pool_;
  std::map<std::string, std::pair<double, long long>> host_prof_;
  void launch_volume(const char* kind, int mode, double t, double wM, double wA, const double* x,
                     const double* z, double* r, double* vals, double* bdiag);
  void launch_facets(const char* kind, double t, double wA, const double* x, const double* z,
                     double* r, double* vals, double* bdiag);
  void build_patches();
  void ensure_element_order();
  bool elem_order_ready_ = false;
  std::vector<std::pair<uint64_t, int32_t>> morton_order(int c) const;
  cudaKernel_t kernel(JitGroup group, const std::string& name);
  std::map<int, std::unique_ptr<JitModule>> jit_;
  int struct_march_ = 8, struct_march_apply_ = 0;
  long long struct_march_fill_ = 0;
  std::string jit_defines_;
  DeviceBuffer<double> coords_, coords4_, cell_, cell_patch_;
  const double* zscale_ = nullptr;   // set_apply_scale: consumed by the structured per-cell apply
  double zrelax_ = 1.0;
  bool vector_gather_ = false;
  bool dofs_even_ = false;   // every compartment's dof block starts at an even offset (16-byte gathers)
  DeviceBuffer<int> elems_;
  std::vector<DeviceBuffer<int>> comp_elem_ids_, comp_vdof_;
  std::vector<DeviceBuffer<int>> comp_pverts_, comp_pdofs_;   // packed connectivity in thread order (kernel_args.h)
  bool packed_conn_ = true;
  bool struct_nomask_ = true;   // apply instantiation without the Dirichlet mask when the operator has no constrained dofs
  std::vector<int64_t> comp_nelem_;
  std::vector<PatchSet> patches_;
  std::vector<FacetList> facets_;
  int64_t nnz_ = 0;
  int64_t ne_patch_total_ = 0;
  size_t patch_smem(const PatchSet& P, int ns, int mode) const;
  // gather form of the CSR fill (kernels/assembly_element.cuh): vertices of a compartment and the
  // elements around them, built with the pattern on first use
  struct GatherSet {
    DeviceBuffer<int> verts, vptr, vel;
    int64_t nverts = 0;
    int maxlen = 0;        // longest CSR row of the compartment
    bool usable = false;
  };
  std::vector<GatherSet> gather_;
  std::string csr_fill_ = "scatter";
  void ensure_gather();
  bool tile_ok_ = false, tile_residual_ = true;
  int tile_w_ = 4, tile_r_ = 2, tile_lz_ = 0, tile_minb_ = 3;
  DeviceBuffer<double> tile_slots_, tile_partials_;
  la::ReduceWorkspace tile_ws_;
  void launch_tile(int mode, double t, double wM, double wA, const double* x, const double* z, double* y, const TileFused* f,
                   bool accumulate);
  int struct_comp_ = -1;   // compartment handled by the structured kernels
  int struct_part_ = 0;    // cell range selector of the next structured launch (jacobian_apply)
  int patch_pn_ = 256, patch_pe_ = 512, patch_threads_ = 256, patch_smem_kb_ = 64;
};

}  // namespace dcb
