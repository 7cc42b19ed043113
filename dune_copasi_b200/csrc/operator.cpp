#include "operator.hpp"

#include <parallel/algorithm>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>

namespace dcb {

namespace {

inline uint64_t spread3(uint64_t x) {   // 21 bits -> every third bit
  x &= 0x1fffff;
  x = (x | x << 32) & 0x1f00000000ffffULL;
  x = (x | x << 16) & 0x1f0000ff0000ffULL;
  x = (x | x << 8) & 0x100f00f00f00f00fULL;
  x = (x | x << 4) & 0x10c30c30c30c30c3ULL;
  x = (x | x << 2) & 0x1249249249249249ULL;
  return x;
}
inline uint64_t spread2(uint64_t x) {   // 31 bits -> every second bit
  x &= 0x7fffffff;
  x = (x | x << 16) & 0x0000ffff0000ffffULL;
  x = (x | x << 8) & 0x00ff00ff00ff00ffULL;
  x = (x | x << 4) & 0x0f0f0f0f0f0f0f0fULL;
  x = (x | x << 2) & 0x3333333333333333ULL;
  x = (x | x << 1) & 0x5555555555555555ULL;
  return x;
}

}  // namespace

DeviceOperator::DeviceOperator(std::shared_ptr<const Model> m, std::shared_ptr<const Grid> g)
    : model(std::move(m)), grid(std::move(g)) {
  require_device();
  if (grid->elem_comp.size() != (size_t)grid->ne) fail("grid is not bound to a model");
  ndofs = grid->ndofs;
  owned = la::Ranges::all(ndofs);
  const PTree& acfg = model->cfg.sub("model.assembly.b200");
  scheme = acfg.get("scheme", std::string("auto"));
  if (scheme != "auto" && scheme != "structured" && scheme != "patch" && scheme != "atomic")
    fail("model.assembly.b200.scheme must be 'auto', 'structured', 'patch' or 'atomic'");
  {
    // the implicit-geometry kernels apply to structured simplex grids whose cells all belong to
    // one compartment, without cell data and with point-independent diffusion coefficients
    bool ok = grid->is_structured && grid->cell_keys.empty();
    int full = -1;
    for (int c = 0; c < model->ncomp() && ok; ++c) {
      int64_t n = 0;
      for (int64_t e = 0; e < grid->ne; ++e) n += grid->elem_comp[e] == c;
      if (n == grid->ne && model->comp_nspec[c] > 0) full = c;
      else if (n != 0 && model->comp_nspec[c] > 0) ok = false;
    }
    int species_comps = 0;   // the structured kernels are generated for single-compartment models only (jit.cpp)
    for (int c = 0; c < model->ncomp(); ++c) species_comps += model->comp_nspec[c] > 0;
    ok = ok && species_comps == 1 && full >= 0 && model->diffusion_is_constant(full) &&
         (int64_t)grid->comp_vertices[full].size() == grid->nv;
    if (scheme == "structured" && !ok)
      fail("model.assembly.b200.scheme = structured needs a structured single-compartment grid without cell data");
    if (grid->elem_kind == 1) {
      // Q1 cube grids exist on the implicit-geometry kernels only (kernels/assembly_q1.cuh)
      if (!ok || (scheme != "auto" && scheme != "structured"))
        fail("Q1 cube grids need one compartment over the whole lattice, no cell data, scalar point-independent "
             "diffusion and model.assembly.b200.scheme = auto|structured");
      if (model->numerical_jacobian) fail("model.jacobian.type = numerical is not built for Q1 cube grids");
      if (model->has_outflow()) fail("outflow / transmission terms on Q1 cube grids are out of scope");
    }
    // measured on B200 (profiles/): element-per-thread + fp64 RED atomics beats the patch kernels
    if (scheme == "auto") scheme = ok ? "structured" : "atomic";
    struct_comp_ = ok ? full : -1;
  }
  // measured on B200 (256^3, profiles/README.md): marching pays for the residual (-7 % P1, -20 % Q1)
  // but not for the apply, whose second field doubles the carried state (+5 %): off there by default
  struct_march_ = acfg.get("struct_march", 8);
  struct_march_apply_ = acfg.get("struct_march_apply", 0);
  if (struct_march_ < 0 || struct_march_ > 64 || struct_march_apply_ < 0 || struct_march_apply_ > 64)
    fail("model.assembly.b200.struct_march / struct_march_apply out of range");
  // threads a launch should keep (4 waves of 148 SMs x 6 CTAs x 64 threads) before columns get shorter
  struct_march_fill_ = acfg.get("struct_march_fill", 148 * 6 * 64 * 4);
  // CSR fill: scatter = element per thread + fp64 atomics (default); gather = one thread per vertex
  // re-integrating the elements around it, rows accumulated in shared memory, no atomics.  Measured on
  // B200 (128^3 Gray-Scott): scatter 4.7 ms, gather 13.8 ms per fill -- the (d+1)-fold redundant
  // element loads and the long serial loop per thread cost more than the atomics (profiles/README.md)
  csr_fill_ = acfg.get("csr_fill", std::string("scatter"));
  if (csr_fill_ != "gather" && csr_fill_ != "scatter") fail("model.assembly.b200.csr_fill must be 'gather' or 'scatter'");
  patch_pn_ = acfg.get("patch_vertices", 256);
  patch_pe_ = acfg.get("patch_elements", 512);
  patch_threads_ = acfg.get("patch_threads", 256);
  patch_smem_kb_ = acfg.get("patch_smem_kb", 64);
  if (patch_pn_ < 16 || patch_pn_ > 16384) fail("model.assembly.b200.patch_vertices out of range");
  if (patch_pe_ < 16 || patch_pe_ > 16383) fail("model.assembly.b200.patch_elements out of range");
  DCB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  // tile-marching drivers: same eligibility as the structured kernels, no facet terms, analytic
  // Jacobian of the standard terms, staged planes within the shared-memory budget
  {
    const int ns = struct_comp_ >= 0 ? model->comp_nspec[struct_comp_] : 0;
    int ns_max = 1;
    for (int c = 0; c < model->ncomp(); ++c) ns_max = std::max(ns_max, model->comp_nspec[c]);
    tile_w_ = grid->dim == 3 ? acfg.get("tile_w", 4) : 1;
    tile_r_ = grid->dim == 3 ? acfg.get("tile_r", 1) : 1;
    tile_minb_ = acfg.get("tile_min_blocks", ns_max <= 2 ? 3 : ns_max <= 4 ? 2 : 1);
    tile_lz_ = acfg.get("tile_lz", 0);
    tile_residual_ = acfg.get("tile_residual", true);
    // off by default: measured on B200 (256^3 Gray-Scott, profiles/README.md) the tile kernels run the apply at
    // 56-59 % of the fp64 pipe (two barriers per layer, 12 warps per SM of which a third stage / finish) against
    // 82 % for the per-cell kernels -- 1.17 ms against 0.84 ms per launch --, which the saved BLAS-1 sweeps
    // (31 -> 11 ms per step) do not win back: 109 against 98.5 ms per time step
    const bool want = acfg.get("tile", false);
    tile_ok_ = want && scheme == "structured" && struct_comp_ >= 0 && !model->has_outflow() &&
               !(model->numerical_jacobian || model->has_extended_terms(struct_comp_)) &&
               (ns % 2 != 0 || grid->comp_offset[struct_comp_] % 2 == 0) &&
               tile_smem_bytes(ns, tile_w_, tile_r_, grid->dim) <= 200 * 1024;
    if (tile_ok_) la::reduce_workspace_create(&tile_ws_, 148 * 256);
  }

  // ---- kernels for this model
  jit_defines_ = jit_defines(*model);
  // the hot group is compiled up front, the others on first use
  kernel(scheme == "patch" ? JitGroup::Patch
         : scheme == "structured" ? (grid->elem_kind == 1 ? JitGroup::StructuredQ1 : JitGroup::Structured)
                                  : JitGroup::Element, "");

  // ---- mesh on the device
  const int ncomp = model->ncomp();
  coords_.upload(grid->coords, stream);
  vector_gather_ = acfg.get("vector_gather", false);
  packed_conn_ = acfg.get("packed_conn", true);
  // same-box A/B (tools/gpu_run_r02n.sh): P1 apply 0.830 -> 0.805 ms, Q1 apply 0.735 -> 0.652 ms
  struct_nomask_ = acfg.get("struct_nomask", true);
  if (vector_gather_ && grid->dim == 3 && !grid->coords.empty()) {
    // padded copy for 16-byte gathers (kernels/assembly_element.cuh)
    std::vector<double> c4((size_t)grid->nv * 4, 0.0);
    for (int64_t v = 0; v < grid->nv; ++v)
      for (int k = 0; k < 3; ++k) c4[v * 4 + k] = grid->coords[v * 3 + k];
    coords4_.upload(c4, stream);
  }
  dofs_even_ = true;
  for (int c = 0; c < model->ncomp(); ++c)
    if (model->comp_nspec[c] > 0 && grid->comp_offset[c] % 2 != 0) dofs_even_ = false;
  elems_.upload(grid->elems, stream);
  if (!grid->cell_data.empty()) cell_.upload(grid->cell_data, stream);
  comp_elem_ids_.resize(ncomp);
  comp_vdof_.resize(ncomp);
  comp_nelem_.assign(ncomp, 0);
  for (int c = 0; c < ncomp; ++c) {
    int64_t n = 0;
    for (int64_t e = 0; e < grid->ne; ++e) n += grid->elem_comp[e] == c;
    comp_nelem_[c] = n;
    // vertex -> dof map; skipped when it is the closed form offset + v*ns
    bool identity_dofs = (int64_t)grid->comp_vertices[c].size() == grid->nv;
    if (!identity_dofs) comp_vdof_[c].upload(grid->comp_vdof[c], stream);
  }
  // ---- facet lists per directional outflow pair
  auto opairs = model->outflow_pairs();
  facets_.resize(opairs.size());
  for (size_t p = 0; p < opairs.size(); ++p) {
    int cs = opairs[p].first, ct = opairs[p].second;
    std::vector<long long> fs, fo;
    std::vector<int> ls, lo;
    for (size_t f = 0; f < grid->f_in.size(); ++f) {
      int64_t ei = grid->f_in[f], eo = grid->f_out[f];
      int ci = grid->elem_comp[ei], co = eo >= 0 ? grid->elem_comp[eo] : -1;
      if (cs == ct) {   // boundary outflow
        if (eo < 0 && ci == cs) { fs.push_back(ei); fo.push_back(-1); ls.push_back(grid->f_lin[f]); lo.push_back(-1); }
      } else if (eo >= 0) {
        if (ci == cs && co == ct) { fs.push_back(ei); fo.push_back(eo); ls.push_back(grid->f_lin[f]); lo.push_back(grid->f_lout[f]); }
        if (co == cs && ci == ct) { fs.push_back(eo); fo.push_back(ei); ls.push_back(grid->f_lout[f]); lo.push_back(grid->f_lin[f]); }
      }
    }
    FacetList& F = facets_[p];
    F.cs = cs; F.ct = ct; F.n = (int64_t)fs.size();
    F.f_self.upload(fs, stream); F.f_other.upload(fo, stream);
    F.f_lself.upload(ls, stream); F.f_lother.upload(lo, stream);
  }
  // ---- constraints
  grid->constraints(*model, h_cdofs, h_cvals);
  ncons = (int64_t)h_cdofs.size();
  if (ncons) {
    cdofs.upload(h_cdofs, stream);
    cvals.upload(h_cvals, stream);
    std::vector<unsigned char> mask(ndofs, 0);
    for (auto d : h_cdofs) mask[d] = 1;
    cmask.upload(mask, stream);
  }
  if (scheme == "patch") build_patches();
  DCB_CUDA(cudaStreamSynchronize(stream));
}

// element lists of the element-per-thread kernels, built on first use.  Default: the mesh's own
// element order -- measured on B200 (128^3 Kuhn grid, apply kernel): mesh order 0.47 ms, Morton
// order 0.74 ms, because consecutive lanes then gather consecutive vertices (coalesced) whereas
// Morton order scatters the lanes of a warp.  model.assembly.b200.element_order = morton is kept
// for meshes whose file order has no locality at all.
void DeviceOperator::ensure_element_order() {
  if (elem_order_ready_) return;
  elem_order_ready_ = true;
  const bool morton = model->cfg.sub("model.assembly.b200").get("element_order", std::string("natural")) == "morton";
  for (int c = 0; c < model->ncomp(); ++c) {
    if (model->comp_nspec[c] == 0 || comp_nelem_[c] == 0) continue;
    std::vector<int> ids;
    if (morton) {
      auto order = morton_order(c);
      ids.resize(order.size());
      for (size_t t = 0; t < order.size(); ++t) ids[t] = order[t].second;
    } else if (comp_nelem_[c] != grid->ne) {
      for (int64_t e = 0; e < grid->ne; ++e)
        if (grid->elem_comp[e] == c) ids.push_back((int)e);
    }
    if (!ids.empty()) comp_elem_ids_[c].upload(ids, stream);
    // packed connectivity in thread order (kernel_args.h): vertex ids and dof bases of the four corners
    if (packed_conn_ && grid->dim == 3 && grid->elem_kind == 0) {
      const int64_t n = comp_nelem_[c];
      const int ns = model->comp_nspec[c];
      std::vector<int> pv((size_t)n * 4), pd((size_t)n * 4);
      const bool identity = comp_vdof_[c].p == nullptr;
      for (int64_t t = 0; t < n; ++t) {
        const int64_t e = ids.empty() ? t : ids[t];
        for (int k = 0; k < 4; ++k) {
          const int v = grid->elems[e * 4 + k];
          pv[t * 4 + k] = v;
          pd[t * 4 + k] = identity ? (int)grid->comp_offset[c] + v * ns : grid->comp_vdof[c][v];
        }
      }
      if (comp_pverts_.size() < (size_t)model->ncomp()) { comp_pverts_.resize(model->ncomp()); comp_pdofs_.resize(model->ncomp()); }
      comp_pverts_[c].upload(pv, stream);
      comp_pdofs_[c].upload(pd, stream);
    }
  }
  DCB_CUDA(cudaStreamSynchronize(stream));
}

cudaKernel_t DeviceOperator::kernel(JitGroup group, const std::string& name) {
  auto& mod = jit_[(int)group];
  if (!mod) {
    mod = std::make_unique<JitModule>();
    const std::string src = jit_source(*model, jit_defines_, group);
    try {
      mod->load(jit_compile_cached(src));
    } catch (const std::exception&) {
      // a damaged cache entry: evict, compile again, and let a second failure surface
      jit_cache_evict(src);
      mod->load(jit_compile_cached(src));
    }
  }
  return name.empty() ? nullptr : mod->kernel(name);
}

DeviceOperator::~DeviceOperator() {
  for (auto& r : prof_) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (auto e : prof_pool_) cudaEventDestroy(e);
  la::reduce_workspace_destroy(&tile_ws_);
  if (stream) cudaStreamDestroy(stream);
}

void DeviceOperator::profile_enable(bool on) {
  profiling_ = on;
  if (!on) profile_collect();
}
void DeviceOperator::prof_begin(const char* kind) {
  if (!profiling_) return;
  ProfRec r;
  r.kind = kind;
  // events are pooled: creating two per kernel would show up as host-side gaps on launch-bound workloads
  for (cudaEvent_t* e : {&r.a, &r.b}) {
    if (!prof_pool_.empty()) { *e = prof_pool_.back(); prof_pool_.pop_back(); }
    else DCB_CUDA(cudaEventCreate(e));
  }
  DCB_CUDA(cudaEventRecord(r.a, stream));
  prof_.push_back(r);
}
void DeviceOperator::prof_end() {
  if (!profiling_ || prof_.empty()) return;
  DCB_CUDA(cudaEventRecord(prof_.back().b, stream));
}
std::map<std::string, std::pair<double, long long>> DeviceOperator::profile_collect() {
  std::map<std::string, std::pair<double, long long>> out = host_prof_;
  host_prof_.clear();
  if (prof_.empty()) return out;
  DCB_CUDA(cudaStreamSynchronize(stream));
  for (auto& r : prof_) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      out[r.kind].first += ms;
      out[r.kind].second += 1;
    }
    prof_pool_.push_back(r.a);
    prof_pool_.push_back(r.b);
  }
  prof_.clear();
  return out;
}

int64_t DeviceOperator::bdiag_size() const {
  int64_t n = 0;
  for (int c = 0; c < model->ncomp(); ++c)
    n += (grid->comp_offset[c + 1] - grid->comp_offset[c]) * model->comp_nspec[c];
  return n;
}

int64_t DeviceOperator::bdiag_shift(int c) const {
  int64_t base = 0;
  for (int k = 0; k < c; ++k) base += (grid->comp_offset[k + 1] - grid->comp_offset[k]) * model->comp_nspec[k];
  return base - grid->comp_offset[c] * model->comp_nspec[c];
}

// ---------------------------------------------------------------------------------- patches
// elements of compartment c sorted along the Morton (Z-order) curve of their centroids: neighbours
// in the list are neighbours in space, which is what keeps vertex data in L1/L2 between elements
std::vector<std::pair<uint64_t, int32_t>> DeviceOperator::morton_order(int c) const {
  const int nd = grid->nd(), dim = grid->dim;
  const int64_t ne = grid->ne;
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int64_t v = 0; v < grid->nv; ++v)
    for (int k = 0; k < dim; ++k) {
      lo[k] = std::min(lo[k], grid->coords[v * dim + k]);
      hi[k] = std::max(hi[k], grid->coords[v * dim + k]);
    }
  const double bits = dim == 3 ? 2097151.0 : 2147483647.0;
  std::vector<std::pair<uint64_t, int32_t>> order;
  order.reserve(comp_nelem_[c]);
  for (int64_t e = 0; e < ne; ++e)
    if (grid->elem_comp[e] == c) order.push_back({0, (int32_t)e});
#pragma omp parallel for schedule(static)
  for (int64_t t = 0; t < (int64_t)order.size(); ++t) {
    int64_t e = order[t].second;
    uint64_t q[3] = {0, 0, 0};
    for (int k = 0; k < dim; ++k) {
      double cen = 0;
      for (int a = 0; a < nd; ++a) cen += grid->coords[(int64_t)grid->elems[e * nd + a] * dim + k];
      cen /= nd;
      double span = hi[k] - lo[k];
      q[k] = (uint64_t)(span > 0 ? (cen - lo[k]) / span * bits : 0.0);
    }
    order[t].first = dim == 3 ? (spread3(q[0]) | spread3(q[1]) << 1 | spread3(q[2]) << 2)
                              : (spread2(q[0]) | spread2(q[1]) << 1);
  }
  __gnu_parallel::sort(order.begin(), order.end());
  return order;
}

void DeviceOperator::build_patches() {
  const int nd = grid->nd(), dim = grid->dim, ncomp = model->ncomp();
  const int64_t ne = grid->ne;
  patches_.resize(ncomp);
  std::vector<double> cell_sorted;
  const size_t nkeys = grid->cell_keys.size();
  // total elements with a compartment (patch element order = compartments concatenated)
  int64_t total = 0;
  for (int c = 0; c < ncomp; ++c) total += model->comp_nspec[c] > 0 ? comp_nelem_[c] : 0;
  ne_patch_total_ = total;
  if (nkeys) cell_sorted.resize(nkeys * (size_t)total);
  int64_t ebase = 0;
  std::vector<int32_t> mark(grid->nv, -1);
  for (int c = 0; c < ncomp; ++c) {
    PatchSet& P = patches_[c];
    P.comp = c;
    if (model->comp_nspec[c] == 0 || comp_nelem_[c] == 0) continue;
    // budgets: the element-result buffer of the apply kernel (nd*ns doubles per element) plus the
    // staged vertex data must fit the shared-memory budget that keeps several CTAs per SM
    {
      const int ns = model->comp_nspec[c];
      int pe = patch_pe_, pn = patch_pn_;
      auto bytes = [&](int pe_, int pn_) {
        return 8.0 * (pn_ * (dim + 2 * ns) + (double)nd * ns * pe_) + 2.0 * (pe_ * nd + pn_ + 1);
      };
      while (pe > 32 && bytes(pe, pn) > patch_smem_kb_ * 1024.0) { pe = pe * 3 / 4; pn = std::max(32, pn * 3 / 4); }
      P.max_elems = pe;
      P.max_nodes = pn;
    }
    const int pe_max = P.max_elems, pn_max = P.max_nodes;
    // 1. Morton order of the element centroids
    std::vector<std::pair<uint64_t, int32_t>> order = morton_order(c);
    const int64_t n = (int64_t)order.size();
    // 2. greedy cuts under the vertex / element budgets
    std::vector<int> elem_ptr{0};
    {
      int cur_nodes = 0, pid = 0;
      std::fill(mark.begin(), mark.end(), -1);
      for (int64_t t = 0; t < n; ++t) {
        int64_t e = order[t].second;
        int fresh = 0;
        for (int a = 0; a < nd; ++a) fresh += mark[grid->elems[e * nd + a]] != pid;
        // duplicates inside one element do not occur (simplex vertices are distinct)
        if (cur_nodes + fresh > pn_max || (t - elem_ptr.back()) >= pe_max) {
          elem_ptr.push_back((int)t);
          ++pid;
          cur_nodes = 0;
          fresh = nd;
        }
        for (int a = 0; a < nd; ++a) mark[grid->elems[e * nd + a]] = pid;
        cur_nodes += fresh;
      }
      elem_ptr.push_back((int)n);
    }
    const int np = (int)elem_ptr.size() - 1;
    // 3. per patch: vertex list (ascending), local connectivity, vertex -> element adjacency
    std::vector<int> node_cnt(np + 1, 0);
    std::vector<std::vector<int>> pnodes(np);
    std::vector<unsigned short> lconn((size_t)n * 4, 0), adj((size_t)n * nd);
#pragma omp parallel for schedule(dynamic, 16)
    for (int p = 0; p < np; ++p) {
      auto& nodes = pnodes[p];
      nodes.reserve((size_t)(elem_ptr[p + 1] - elem_ptr[p]) * nd);
      for (int t = elem_ptr[p]; t < elem_ptr[p + 1]; ++t)
        for (int a = 0; a < nd; ++a) nodes.push_back(grid->elems[(int64_t)order[t].second * nd + a]);
      std::sort(nodes.begin(), nodes.end());
      nodes.erase(std::unique(nodes.begin(), nodes.end()), nodes.end());
      for (int t = elem_ptr[p]; t < elem_ptr[p + 1]; ++t)
        for (int a = 0; a < nd; ++a) {
          int v = grid->elems[(int64_t)order[t].second * nd + a];
          lconn[(size_t)t * 4 + a] =
              (unsigned short)(std::lower_bound(nodes.begin(), nodes.end(), v) - nodes.begin());
        }
      node_cnt[p + 1] = (int)nodes.size();
    }
    std::vector<int> node_ptr(np + 1, 0);
    for (int p = 0; p < np; ++p) node_ptr[p + 1] = node_ptr[p] + node_cnt[p + 1];
    std::vector<int> nodes_flat(node_ptr[np]);
    std::vector<int> adj_ptr((size_t)node_ptr[np] + 1, 0);
#pragma omp parallel for schedule(dynamic, 16)
    for (int p = 0; p < np; ++p) {
      std::copy(pnodes[p].begin(), pnodes[p].end(), nodes_flat.begin() + node_ptr[p]);
      // counts per local vertex (stored shifted by one, prefix-summed globally below)
      for (int t = elem_ptr[p]; t < elem_ptr[p + 1]; ++t)
        for (int a = 0; a < nd; ++a) adj_ptr[(size_t)node_ptr[p] + lconn[(size_t)t * 4 + a] + 1]++;
    }
    for (size_t i = 0; i + 1 < adj_ptr.size(); ++i) adj_ptr[i + 1] += adj_ptr[i];
#pragma omp parallel for schedule(dynamic, 16)
    for (int p = 0; p < np; ++p) {
      std::vector<int> cur(adj_ptr.begin() + node_ptr[p], adj_ptr.begin() + node_ptr[p + 1]);
      for (int t = elem_ptr[p]; t < elem_ptr[p + 1]; ++t)     // ascending local element index
        for (int a = 0; a < nd; ++a) {
          int ln = lconn[(size_t)t * 4 + a];
          adj[cur[ln]++] = (unsigned short)(((t - elem_ptr[p]) << 2) | a);
        }
    }
    // patch-ordered cell data
    for (size_t k = 0; k < nkeys; ++k)
      for (int64_t t = 0; t < n; ++t)
        cell_sorted[k * (size_t)total + ebase + t] = grid->cell_data[k * ne + order[t].second];
    // element ranges are stored relative to the global patch element order
    for (auto& x : elem_ptr) x += (int)ebase;
    P.npatch = np;
    P.elem_begin = ebase;
    P.nelem = n;
    P.total_nodes = node_ptr[np];
    P.node_ptr.upload(node_ptr, stream);
    P.nodes.upload(nodes_flat, stream);
    P.elem_ptr.upload(elem_ptr, stream);
    P.adj_ptr.upload(adj_ptr, stream);
    P.lconn.upload(lconn, stream);
    P.adj.upload(adj, stream);
    DCB_CUDA(cudaStreamSynchronize(stream));
    ebase += n;
  }
  if (nkeys) cell_patch_.upload(cell_sorted, stream);
  DCB_CUDA(cudaStreamSynchronize(stream));
}

size_t DeviceOperator::patch_smem(const PatchSet& P, int ns, int mode) const {
  const int nv = mode == 2 ? ns * ns : ns, nd = grid->nd();
  size_t doubles = (size_t)P.max_nodes * (grid->dim + ns + (mode == 1 ? ns : 0)) + (size_t)nd * nv * P.max_elems;
  size_t shorts = (size_t)P.max_elems * nd + P.max_nodes + 1;
  return doubles * 8 + ((shorts * 2 + 15) / 16) * 16;
}

// ---------------------------------------------------------------------------------- launches
void DeviceOperator::launch_volume(const char* kind, int mode, double t, double wM, double wA,
                                   const double* x, const double* z, double* r, double* vals,
                                   double* bdiag) {
  const int ncomp = model->ncomp();
  const bool use_patch = scheme == "patch" && mode != 3;   // mode 3 = CSR fill (element kernels)
  for (int c = 0; c < ncomp; ++c) {
    const int ns = model->comp_nspec[c];
    if (ns == 0 || comp_nelem_[c] == 0) continue;
    if (mode == 4 && !(scheme == "structured" && c == struct_comp_)) continue;   // scalar diagonal: structured only
    // finite-difference Jacobians (model.jacobian.type = numerical) live in the element kernels
    // ... and so do the general analytic Jacobians of advection / tensor / dD/du terms
    const bool fd = (model->numerical_jacobian || model->has_extended_terms(c)) && mode != 0;
    const bool q1 = grid->elem_kind == 1;
    if (scheme == "structured" && (mode != 3 || q1) && c == struct_comp_ && (!fd || q1)) {
      DcStructArgs a{};
      a.ncells = 1;
      for (int k = 0; k < 3; ++k) {
        a.n[k] = k < grid->dim ? grid->s_cells[k] : 1;
        a.h[k] = grid->s_h[k];
        a.rh[k] = 1.0 / grid->s_h[k];
        a.origin[k] = grid->s_origin[k];
        a.ncells *= a.n[k];
      }
      a.vol = struct_simplex_volume();
      a.dof_offset = (int)grid->comp_offset[c];
      a.time = t; a.wM = wM; a.wA = wA; a.x = x; a.z = z; a.r = r;
      a.bdiag = bdiag ? (mode == 4 ? bdiag : bdiag + bdiag_shift(c)) : nullptr;
      a.cmask = cmask.p;
      a.zscale = nullptr; a.zrelax = 1.0;
      a.rowptr = (const long long*)rowptr.p; a.colidx = colidx.p; a.vals = vals;
      static const char* sn[5] = {"dc_k_struct_residual_", "dc_k_struct_apply_", "dc_k_struct_bdiag_", "", "dc_k_struct_diag_"};
      static const char* qn[5] = {"dc_k_q1_residual_", "dc_k_q1_apply_", "dc_k_q1_bdiag_", "dc_k_q1_csr_", "dc_k_q1_diag_"};
      static const char* sk[5] = {"struct_residual", "struct_apply", "struct_bdiag", "struct_csr", "struct_diag"};
      // residual / apply: register marching along the last axis (model.assembly.b200.struct_march
      // cells per thread, 0 = one thread per cell); shortened when the box is too thin to fill the GPU
      int march = mode == 0 ? struct_march_ : mode == 1 ? struct_march_apply_ : 0;
      const long long total = a.ncells, layer = total / a.n[grid->dim - 1];
      while (march > 1 && layer * ((a.n[grid->dim - 1] + march - 1) / march) < struct_march_fill_) march /= 2;
      a.march = march;
      if (mode == 1 && zscale_) {
        if (march > 0) fail("internal: scaled Jacobian apply with a marching driver");
        a.zscale = zscale_; a.zrelax = zrelax_;
      }
      const std::string kname = march > 0 ? std::string(q1 ? "dc_k_q1_march_" : "dc_k_struct_march_") + (mode == 0 ? "residual_" : "apply_")
                                : a.zscale ? std::string(q1 ? "dc_k_q1_apply_scaled_" : "dc_k_struct_apply_scaled_")
                                : (mode == 1 && ncons == 0 && struct_nomask_) ? std::string(q1 ? "dc_k_q1_apply_nomask_" : "dc_k_struct_apply_nomask_")
                                          : std::string(q1 ? qn[mode] : sn[mode]);
      cudaKernel_t k = kernel(q1 ? JitGroup::StructuredQ1 : JitGroup::Structured, kname + std::to_string(c));
      const int sth = model->cfg.sub("model.assembly.b200").get("struct_threads", 32);
      // cell ranges of this launch: everything, or (multi-GPU overlap) the interior layers /
      // the two layers along the slab axis that touch ghost planes
      long long ranges[2][2] = {{0, total}, {0, 0}};
      int nranges = 1;
      if (struct_part_ == 1) { ranges[0][0] = layer; ranges[0][1] = total - layer; }
      if (struct_part_ == 2) { ranges[0][1] = layer; ranges[1][0] = total - layer; ranges[1][1] = total; nranges = 2; }
      for (int q = 0; q < nranges; ++q) {
        a.cell_begin = ranges[q][0];
        a.ncells = ranges[q][1];
        if (a.ncells <= a.cell_begin) continue;
        long long nthreads = a.ncells - a.cell_begin;
        if (march > 0) {
          const long long layers = nthreads / layer;
          nthreads = layer * ((layers + march - 1) / march);
        }
        ProfScope ps(this, struct_part_ == 2 ? "struct_apply_halo_layers" : sk[mode]);
        jit_launch(k, (unsigned)((nthreads + sth - 1) / sth), sth, 0, stream, a);
        stats.launches++;
      }
      continue;
    }
    // the block-diagonal buffer grows with ns^2: fall back to the element kernel when it cannot be staged
    const bool patch_here = use_patch && !fd && patch_smem(patches_[c], ns, mode) <= 200 * 1024;
    if (patch_here) {
      const PatchSet& P = patches_[c];
      DcPatchArgs a{};
      a.coords = coords_.p;
      a.patch_node_ptr = P.node_ptr.p; a.patch_nodes = P.nodes.p; a.patch_elem_ptr = P.elem_ptr.p;
      a.lconn = P.lconn.p - 0; a.adj = P.adj.p; a.adj_ptr = P.adj_ptr.p;
      // lconn/adj are indexed by global patch element ids: shift the base pointers of this set
      a.lconn = P.lconn.p - (size_t)P.elem_begin * 4;
      a.vdof = comp_vdof_[c].p;
      a.cell = cell_patch_.p; a.ne_total = ne_patch_total_;
      a.npatch = P.npatch; a.dof_offset = (int)grid->comp_offset[c];
      a.time = t; a.wM = wM; a.wA = wA;
      a.x = x; a.z = z; a.r = r; a.bdiag = bdiag ? bdiag + bdiag_shift(c) : nullptr; a.cmask = cmask.p;
      a.max_nodes = P.max_nodes; a.max_elems = P.max_elems;
      const size_t smem = patch_smem(P, ns, mode);
      static const char* names[3] = {"dc_k_patch_residual_", "dc_k_patch_apply_", "dc_k_patch_bdiag_"};
      cudaKernel_t k = kernel(JitGroup::Patch, std::string(names[mode]) + std::to_string(c));
      DCB_CUDA(cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      // persistent grid: a multiple of the 148 SMs x resident CTAs (shared memory / thread limits)
      int per_sm = std::max(1, (int)(220 * 1024 / (smem + 1024)));
      per_sm = std::min(per_sm, std::max(1, 2048 / patch_threads_));
      unsigned gridsz = (unsigned)std::min<int64_t>(P.npatch, (int64_t)148 * per_sm);
      static const char* pk[3] = {"patch_residual", "patch_apply", "patch_bdiag"};
      ProfScope ps(this, pk[mode]);
      jit_launch(k, gridsz, patch_threads_, smem, stream, a);
    } else {
      ensure_element_order();
      if (mode == 1 && zscale_) fail("internal: scaled Jacobian apply reached the element kernels");
      DcVolArgs a{};
      a.coords = coords_.p; a.elems = elems_.p; a.elem_ids = comp_elem_ids_[c].p; a.vdof = comp_vdof_[c].p;
      a.cell = cell_.p; a.ne_total = grid->ne; a.n = comp_nelem_[c];
      a.dof_offset = (int)grid->comp_offset[c];
      a.time = t; a.wM = wM; a.wA = wA; a.x = x; a.z = z; a.r = r;
      a.rowptr = (const long long*)rowptr.p; a.colidx = colidx.p; a.vals = vals;
      a.bdiag = bdiag ? bdiag + bdiag_shift(c) : nullptr;
      a.cmask = cmask.p;
      // 16-byte gathers (padded coordinates, double2 dof loads): an option, off by default -- measured on B200 (cell
      // model on the nested mesh, 96^3) neutral while the plain scatter bound the kernel, -2 % with the transposed
      // scatter at 128 registers, +19 % at 164 registers (profiles/r02_cell10.md)
      if (c < (int)comp_pverts_.size() && comp_pverts_[c].p) { a.pverts = comp_pverts_[c].p; a.pdofs = comp_pdofs_[c].p; }
      a.coords4 = vector_gather_ ? coords4_.p : nullptr;
      a.vec = vector_gather_ && ns % 2 == 0 && dofs_even_ &&
              !((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(z)) & 15u);
      if (mode == 3 && csr_fill_ == "gather" && !fd) {
        ensure_gather();
        const GatherSet& G = gather_[c];
        if (G.usable) {
          a.verts = G.verts.p; a.vptr = G.vptr.p; a.vel = G.vel.p; a.gather_maxlen = G.maxlen;
          a.n = G.nverts;
          const int th = 64;
          const size_t smem = (size_t)G.maxlen * ns * sizeof(double) * th;
          cudaKernel_t k = kernel(JitGroup::Csr, "dc_k_jacobian_gather_" + std::to_string(c));
          DCB_CUDA(cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          ProfScope ps(this, "csr_fill");
          jit_launch(k, (unsigned)((a.n + th - 1) / th), th, smem, stream, a);
          stats.launches++;
          continue;
        }
      }
      const bool nomask = mode == 1 && ncons == 0 && struct_nomask_;   // direction used as loaded (no Dirichlet rows)
      cudaKernel_t k = kernel(mode == 3 ? JitGroup::Csr : JitGroup::Element,
                              std::string(kind) + (nomask ? "nomask_" : "") + std::to_string(c));
      static const char* ek[4] = {"elem_residual", "elem_apply", "elem_bdiag", "csr_fill"};
      ProfScope ps(this, ek[mode]);
      jit_launch(k, (unsigned)((a.n + 127) / 128), 128, 0, stream, a);
    }
    stats.launches++;
  }
}

void DeviceOperator::launch_facets(const char* kind, double t, double wA, const double* x,
                                   const double* z, double* r, double* vals, double* bdiag) {
  // one fused launch over the facet lists of all directional pairs; the argument block mirrors the
  // generated `struct DcFacetArgsAll { DcFacetArgs a[np]; int first[np + 1]; }` (jit.cpp)
  const size_t np = facets_.size();
  if (np == 0) return;
  std::vector<char> blob(np * sizeof(DcFacetArgs) + (np + 1) * sizeof(int), 0);
  auto* args = reinterpret_cast<DcFacetArgs*>(blob.data());
  auto* first = reinterpret_cast<int*>(blob.data() + np * sizeof(DcFacetArgs));
  int blocks = 0;
  for (size_t p = 0; p < np; ++p) {
    const FacetList& F = facets_[p];
    DcFacetArgs a{};
    a.coords = coords_.p; a.elems = elems_.p;
    a.f_self = F.f_self.p; a.f_other = F.f_other.p; a.f_lself = F.f_lself.p; a.f_lother = F.f_lother.p;
    a.vdof_s = comp_vdof_[F.cs].p; a.vdof_t = comp_vdof_[F.ct].p;
    a.cell = cell_.p; a.ne_total = grid->ne; a.n = F.n;
    a.dof_offset_s = (int)grid->comp_offset[F.cs]; a.dof_offset_t = (int)grid->comp_offset[F.ct];
    a.block_offset = blocks;
    a.time = t; a.wA = wA; a.x = x; a.z = z; a.r = r;
    a.rowptr = (const long long*)rowptr.p; a.colidx = colidx.p; a.vals = vals;
    a.bdiag = bdiag ? bdiag + bdiag_shift(F.cs) : nullptr;
    a.cmask = cmask.p;
    args[p] = a;
    first[p] = blocks;
    blocks += (int)((F.n + 63) / 64);
  }
  first[np] = blocks;
  if (blocks == 0) return;
  cudaKernel_t k = kernel(JitGroup::Skeleton, kind);
  ProfScope ps(this, "facets");
  void* params[] = {(void*)blob.data()};
  DCB_CUDA(cudaLaunchKernel((const void*)k, dim3((unsigned)blocks), dim3(64), params, 0, stream));
  stats.launches++;
}

// ---------------------------------------------------------------------------------- tile marching
size_t tile_smem_bytes(int ns, int tile_w, int tile_r, int dim) {
  // 14 planes of 33 x (rows + 1) vertices (3 u, 3 direction, 2 epilogue operand, 4 raw, 2 accumulators) + spill rows
  const size_t rows = dim == 3 ? (size_t)tile_w * tile_r + 1 : 1, w = dim == 3 ? tile_w : 1;
  return (14 * 33 * rows + 2 * w * 33) * ns * sizeof(double);
}

// vectors of the tile kernels move as 16-byte accesses when the species count is even
bool DeviceOperator::tile_aligned(std::initializer_list<const void*> ptrs) const {
  if (!tile_ok_) return false;
  if (model->comp_nspec[struct_comp_] % 2 != 0) return true;
  for (const void* p : ptrs)
    if (reinterpret_cast<uintptr_t>(p) & 15u) return false;
  return true;
}

void DeviceOperator::launch_tile(int mode, double t, double wM, double wA, const double* x, const double* z, double* y,
                                 const TileFused* f, bool accumulate) {
  if (!tile_ok_) fail("tile-marching kernels are not available for this operator");
  const int c = struct_comp_, dim = grid->dim, L = dim - 1;
  const int ns = model->comp_nspec[c];
  const bool q1 = grid->elem_kind == 1;
  DcTileArgs A{};
  DcStructArgs& a = A.s;
  a.ncells = 1;
  for (int k = 0; k < 3; ++k) {
    a.n[k] = k < dim ? grid->s_cells[k] : 1;
    a.h[k] = grid->s_h[k];
    a.rh[k] = 1.0 / grid->s_h[k];
    a.origin[k] = grid->s_origin[k];
    a.ncells *= a.n[k];
  }
  a.vol = struct_simplex_volume();
  a.dof_offset = (int)grid->comp_offset[c];
  a.time = t; a.wM = wM; a.wA = wA; a.x = x; a.z = z; a.r = y;
  a.cmask = cmask.p;
  const int tile_x_ = 32, tile_y_ = tile_w_ * tile_r_;
  A.ntx = (a.n[0] + tile_x_ - 1) / tile_x_;
  A.nty = dim == 3 ? (a.n[1] + tile_y_ - 1) / tile_y_ : 1;
  // chunks along the marching axis: enough CTAs for ~8 waves of 148 SMs x resident CTAs, but chunks
  // of at least 8 layers (every chunk stages one extra plane and leaves one more cut plane)
  const int nL = a.n[L];
  int lz = tile_lz_;
  if (lz <= 0) {
    const long long tiles = (long long)A.ntx * A.nty;
    const long long want = (148LL * tile_minb_ * 8 + tiles - 1) / tiles;
    lz = (int)std::max<long long>(8, (nL + want - 1) / want);
  }
  lz = std::min(lz, nL);
  A.lz = lz;
  const int nchunk = (nL + lz - 1) / lz;
  const long long plane = (long long)(a.n[0] + 1) * (dim == 3 ? a.n[1] + 1 : 1);
  A.own_lo = 0; A.own_hi = nL + 1;
  if (grid->n_owned >= 0) {
    A.own_lo = (int)(grid->owned_begin / plane);
    A.own_hi = A.own_lo + (int)(grid->n_owned / plane);
  }
  if (f) {
    A.pro = f->pro; A.epi = f->epi; A.first = f->first; A.relax = f->relax;
    A.r_in = f->r_in; A.p_in = f->p_in; A.v_in = f->v_in; A.dinv = f->dinv; A.w = f->w;
    A.r_out = f->r_out; A.p_out = f->p_out;
    A.rho_new = f->rho_new; A.rho = f->rho; A.hptr = f->hptr; A.trtt = f->trtt;
  }
  A.accumulate = accumulate;
  A.identity = mode == 1 && !accumulate && A.pro == 0 && ncons > 0;
  const unsigned nblocks = (unsigned)((long long)A.ntx * A.nty * nchunk);
  if (tile_slots_.n < 7 * ndofs) tile_slots_.alloc(7 * ndofs);
  if (tile_partials_.n < (int64_t)nblocks * 4) tile_partials_.alloc((int64_t)nblocks * 4);
  A.slots = tile_slots_.p; A.slot_stride = ndofs; A.partials = tile_partials_.p;
  const size_t smem = tile_smem_bytes(ns, tile_w_, tile_r_, dim);
  if (!tile_aligned({x, z, y, A.r_in, A.p_in, A.v_in, A.dinv, A.w, A.r_out, A.p_out}))
    fail("tile-marching kernels: vectors of an even number of species must be 16-byte aligned");
  const std::string kname = std::string("dc_k_tile_") + (q1 ? "q1_" : "") + (mode == 0 ? "residual_" : "apply_") + std::to_string(c);
  cudaKernel_t k = kernel(q1 ? JitGroup::TileQ1 : JitGroup::Tile, kname);
  DCB_CUDA(cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  {
    ProfScope ps(this, mode == 0 ? "tile_residual" : "tile_apply");
    jit_launch(k, nblocks, (unsigned)(32 * (dim == 3 ? tile_w_ : 1)), smem, stream, A);
    stats.launches++;
  }
  la::TileFixup F{};
  F.n[0] = a.n[0]; F.n[1] = dim == 3 ? a.n[1] : 0; F.n[2] = nL;
  F.tile[0] = tile_x_; F.tile[1] = dim == 3 ? tile_y_ : 1; F.tile[2] = lz;
  F.ns = ns; F.dof_offset = a.dof_offset; F.own_lo = A.own_lo; F.own_hi = A.own_hi;
  F.epi = A.epi; F.y = y; F.slots = tile_slots_.p; F.slot_stride = ndofs;
  F.w = A.w; F.aux = A.r_out;
  F.cmask = A.identity ? cmask.p : nullptr; F.zraw = z;
  F.main_partials = tile_partials_.p; F.nmain = (int)nblocks;
  F.out = f ? f->out : nullptr; F.out_mask = f && f->out ? f->out_mask : 0;
  {
    ProfScope ps(this, "tile_fixup");
    la::tile_fixup(F, tile_ws_, stream);
    stats.launches++;
  }
}

void DeviceOperator::tile_apply(double t, double wM, double wA, const double* x, const double* z, double* y,
                                const TileFused* f) {
  launch_tile(1, t, wM, wA, x, z, y, f, false);
}

void DeviceOperator::tile_residual(double t, double wM, double wA, const double* x, double* r) {
  launch_tile(0, t, wM, wA, x, nullptr, r, nullptr, true);
}

void DeviceOperator::residual(double t, double wM, double wA, const double* x, double* r) {
  if (tile_residual_ && tile_aligned({x, r})) { tile_residual(t, wM, wA, x, r); return; }
  launch_volume("dc_k_residual_volume_", 0, t, wM, wA, x, nullptr, r, nullptr, nullptr);
  if (wA != 0.0) launch_facets("dc_k_skeleton_residual", t, wA, x, nullptr, r, nullptr, nullptr);
}

bool DeviceOperator::can_split_apply() const {
  int with_species = 0;
  for (int c = 0; c < model->ncomp(); ++c) with_species += model->comp_nspec[c] > 0;
  return scheme == "structured" && facets_.empty() && with_species == 1 && struct_comp_ >= 0 &&
         !model->numerical_jacobian && grid->s_cells[grid->dim - 1] >= 3;
}

// prod(h) / dim!, in the kernels' operation order (h[0] * h[1] * h[2], then one division)
double DeviceOperator::struct_simplex_volume() const {
  double adet = 1.0;
  for (int k = 0; k < grid->dim; ++k) adet *= grid->s_h[k];
  return adet / (grid->dim == 3 ? 6.0 : 2.0);
}

// the per-cell structured apply can form its direction as relax * dinv .* z while it loads the corners
bool DeviceOperator::apply_scale_ready() const {
  // (finite-difference and extended-term Jacobians are applied by the element kernels even on structured grids)
  return scheme == "structured" && struct_comp_ >= 0 && struct_march_apply_ == 0 && !tile_ready() && facets_.empty() &&
         ncons == 0 && model->ncomp() == 1 && !model->numerical_jacobian && !model->has_extended_terms(struct_comp_);
}

void DeviceOperator::jacobian_apply(double t, double wM, double wA, const double* x, const double* z, double* y,
                                    int part) {
  if (part != 0 && !can_split_apply()) fail("jacobian_apply: this operator cannot be split into interior / halo layers");
  if (part == 0 && tile_aligned({x, z, y})) { launch_tile(1, t, wM, wA, x, z, y, nullptr, true); return; }   // y += J z
  struct_part_ = part;
  launch_volume("dc_k_jacobian_apply_volume_", 1, t, wM, wA, x, z, y, nullptr, nullptr);
  struct_part_ = 0;
  if (part != 0) return;   // no facet terms on a splittable operator
  if (wA != 0.0) launch_facets("dc_k_skeleton_apply", t, wA, x, z, y, nullptr, nullptr);
}

void DeviceOperator::block_diag(double t, double wM, double wA, const double* x, double* bdiag) {
  launch_volume("dc_k_bdiag_volume_", 2, t, wM, wA, x, nullptr, nullptr, nullptr, bdiag);
  if (wA != 0.0) launch_facets("dc_k_skeleton_bdiag", t, wA, x, nullptr, nullptr, nullptr, bdiag);
}

bool DeviceOperator::scalar_diag(double t, double wM, double wA, const double* x, double* diag) {
  if (scheme != "structured" || !facets_.empty() || model->ncomp() != 1 || model->numerical_jacobian) return false;
  launch_volume("", 4, t, wM, wA, x, nullptr, nullptr, nullptr, diag);
  return true;
}

void DeviceOperator::jacobian_csr(double t, double wM, double wA, const double* x, double* vals) {
  ensure_csr();
  launch_volume("dc_k_jacobian_volume_", 3, t, wM, wA, x, nullptr, nullptr, vals, nullptr);
  if (wA != 0.0) launch_facets("dc_k_skeleton_jacobian", t, wA, x, nullptr, nullptr, vals, nullptr);
}

void DeviceOperator::ensure_gather() {
  if (!gather_.empty()) return;
  ensure_csr();
  const int ncomp = model->ncomp(), nd = grid->nd();
  gather_.resize(ncomp);
  for (int c = 0; c < ncomp; ++c) {
    GatherSet& G = gather_[c];
    const int ns = model->comp_nspec[c];
    if (ns == 0 || comp_nelem_[c] == 0 || grid->elem_kind != 0) continue;
    if (grid->ne * 4 > (int64_t)INT32_MAX) continue;   // (element << 2 | local index) in 32 bits
    const auto& verts = grid->comp_vertices[c];
    const int64_t n = (int64_t)verts.size();
    // longest row of the compartment -> shared-memory slots per thread
    int maxlen = 0;
    for (int64_t lv = 0; lv < n; ++lv)
      for (int i = 0; i < ns; ++i) {
        const int64_t row = grid->comp_vdof[c][verts[lv]] + i;
        maxlen = std::max(maxlen, (int)(h_rowptr[row + 1] - h_rowptr[row]));
      }
    if ((size_t)maxlen * ns * sizeof(double) * 64 > 200 * 1024) continue;
    std::vector<int> local(grid->nv, -1);
    for (int64_t lv = 0; lv < n; ++lv) local[verts[lv]] = (int)lv;
    std::vector<int> vptr(n + 1, 0);
    for (int64_t e = 0; e < grid->ne; ++e)
      if (grid->elem_comp[e] == c)
        for (int a = 0; a < nd; ++a) vptr[local[grid->elems[e * nd + a]] + 1]++;
    for (int64_t lv = 0; lv < n; ++lv) vptr[lv + 1] += vptr[lv];
    std::vector<int> vel(vptr[n]), cur(vptr.begin(), vptr.end() - 1);
    for (int64_t e = 0; e < grid->ne; ++e)
      if (grid->elem_comp[e] == c)
        for (int a = 0; a < nd; ++a) vel[cur[local[grid->elems[e * nd + a]]]++] = (int)(e << 2 | a);
    if (n != grid->nv) G.verts.upload(std::vector<int>(verts.begin(), verts.end()), stream);
    G.vptr.upload(vptr, stream);
    G.vel.upload(vel, stream);
    G.nverts = n;
    G.maxlen = maxlen;
    G.usable = true;
    DCB_CUDA(cudaStreamSynchronize(stream));
  }
}

void DeviceOperator::ensure_csr() {
  if (rowptr.n) return;
  grid->pattern(*model, h_rowptr, h_colidx);
  nnz_ = h_rowptr[ndofs];
  rowptr.upload(h_rowptr, stream);
  colidx.upload(h_colidx, stream);
  if (nnz_ < INT32_MAX) {
    std::vector<int32_t> rp32(h_rowptr.begin(), h_rowptr.end());
    rowptr32.upload(rp32, stream);
  }
  DCB_CUDA(cudaStreamSynchronize(stream));
}

}  // namespace dcb
