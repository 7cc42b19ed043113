#include "grid.hpp"

#include <algorithm>
#include <array>
#include <cfloat>
#include <cmath>
#include <numeric>

namespace dcb {

Grid Grid::structured(int dim, const int* cells, const double* origin, const double* extent, int elem_kind) {
  if (dim < 2 || dim > 3) fail("structured grid: dim must be 2 or 3");
  if (elem_kind != 0 && elem_kind != 1) fail("structured grid: element kind must be 0 (simplex) or 1 (cube)");
  for (int a = 0; a < dim; ++a)
    if (cells[a] < 1) fail("structured grid: cells must be >= 1");
  Grid g = structured_box(dim, cells, origin, extent, 0, cells[dim - 1], elem_kind);
  for (int a = 0; a < dim; ++a) { g.s_origin_exact_[a] = origin[a]; g.s_extent_exact_[a] = extent[a]; }
  return g;
}

// cube layers [layer_lo, layer_hi) along the last axis of the global lattice `cells`
Grid Grid::structured_box(int dim, const int* cells, const double* origin, const double* extent,
                          int layer_lo, int layer_hi, int elem_kind) {
  Grid g;
  g.dim = dim;
  g.is_structured = true;
  g.elem_kind = elem_kind;
  g.s_layer_lo = layer_lo;
  g.s_layers_global = cells[dim - 1];
  int64_t ncg[3] = {1, 1, 1}, nc[3] = {1, 1, 1}, nvs[3] = {1, 1, 1}, off[3] = {0, 0, 0};
  for (int a = 0; a < dim; ++a) ncg[a] = nc[a] = cells[a];
  nc[dim - 1] = layer_hi - layer_lo;
  off[dim - 1] = layer_lo;
  for (int a = 0; a < dim; ++a) {
    nvs[a] = nc[a] + 1;
    g.s_cells[a] = (int)nc[a];
    g.s_h[a] = extent[a] / (double)ncg[a];
    g.s_origin[a] = origin[a] + extent[a] * ((double)off[a] / (double)ncg[a]);
  }
  g.nv = nvs[0] * nvs[1] * nvs[2];
  int nperm = elem_kind == 1 ? 1 : dim == 2 ? 2 : 6;
  g.ne = nc[0] * nc[1] * nc[2] * nperm;
  if (g.nv > INT32_MAX || g.ne > (int64_t)INT32_MAX) fail("structured grid too large for one device");
  g.coords.resize(g.nv * dim);
#pragma omp parallel for schedule(static)
  for (int64_t k = 0; k < nvs[2]; ++k)
    for (int64_t j = 0; j < nvs[1]; ++j)
      for (int64_t i = 0; i < nvs[0]; ++i) {
        int64_t v = i + nvs[0] * (j + nvs[1] * k);
        int64_t idx[3] = {i, j, k};
        // the global formula, so that a slab carries bit-identical coordinates
        for (int a = 0; a < dim; ++a)
          g.coords[v * dim + a] = origin[a] + extent[a] * ((double)(idx[a] + off[a]) / (double)ncg[a]);
      }
  // Kuhn simplices: walk from the lowest corner along the axes in the order of the
  // lexicographically enumerated permutations of (0..dim-1)
  static const int perm2[2][2] = {{0, 1}, {1, 0}};
  static const int perm3[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
  int64_t stride[3] = {1, nvs[0], nvs[0] * nvs[1]};
  int nd = g.nd();
  g.elems.resize(g.ne * nd);
#pragma omp parallel for schedule(static)
  for (int64_t k = 0; k < nc[2]; ++k)
    for (int64_t j = 0; j < nc[1]; ++j)
      for (int64_t i = 0; i < nc[0]; ++i) {
        int64_t cube = i + nc[0] * (j + nc[1] * k);
        int64_t base = i * stride[0] + j * stride[1] + k * stride[2];
        if (elem_kind == 1) {   // the cell itself, corners by bit pattern
          for (int m = 0; m < nd; ++m) {
            int64_t v = base;
            for (int a = 0; a < dim; ++a) v += ((m >> a) & 1) * stride[a];
            g.elems[cube * nd + m] = (int32_t)v;
          }
          continue;
        }
        for (int p = 0; p < nperm; ++p) {
          int64_t cur = base;
          int32_t* el = &g.elems[(cube * nperm + p) * nd];
          el[0] = (int32_t)cur;
          for (int s = 0; s < dim; ++s) {
            cur += stride[dim == 2 ? perm2[p][s] : perm3[p][s]];
            el[s + 1] = (int32_t)cur;
          }
        }
      }
  return g;
}

Grid Grid::from_arrays(int dim, int64_t nv, const double* coords, int64_t ne, const int32_t* elems,
                       const std::vector<std::string>& keys, const double* cell_data) {
  if (dim < 2 || dim > 3) fail("mesh: dim must be 2 or 3");
  Grid g;
  g.dim = dim; g.nv = nv; g.ne = ne;
  g.coords.assign(coords, coords + nv * dim);
  g.elems.assign(elems, elems + ne * (dim + 1));
  for (auto v : g.elems)
    if (v < 0 || v >= nv) fail("mesh: element vertex index out of range");
  g.cell_keys = keys;
  if (!keys.empty()) g.cell_data.assign(cell_data, cell_data + (int64_t)keys.size() * ne);
  return g;
}

namespace {
struct FaceRec {
  int32_t v[3];
  int32_t l;
  int64_t e;
  bool same(const FaceRec& o) const { return v[0] == o.v[0] && v[1] == o.v[1] && v[2] == o.v[2]; }
  bool operator<(const FaceRec& o) const {
    if (v[0] != o.v[0]) return v[0] < o.v[0];
    if (v[1] != o.v[1]) return v[1] < o.v[1];
    if (v[2] != o.v[2]) return v[2] < o.v[2];
    return e < o.e;
  }
};
}  // namespace

void Grid::bind(const Model& model) {
  if (model.dim != dim) fail("model and grid dimensions differ");
  const int ndl = nd();
  // ---- compartments: expression on the cell centre != 0 (make_multi_domain_grid.hh:118-155)
  elem_comp.assign(ne, -1);
  int ncomp = model.ncomp();
  bool overlap = false;
  for (int c = 0; c < ncomp; ++c) {
    double cv;
    if (is_constant(model.comp_expr[c], &cv)) {
      if (cv != 0.0) {
        for (int64_t e = 0; e < ne; ++e) { overlap |= elem_comp[e] >= 0; elem_comp[e] = c; }
      }
      continue;
    }
#pragma omp parallel for schedule(static) reduction(|| : overlap)
    for (int64_t e = 0; e < ne; ++e) {
      double cen[3] = {0, 0, 0}, cell[32];
      for (int a = 0; a < ndl; ++a)
        for (int k = 0; k < dim; ++k) cen[k] += coords[(int64_t)elems[e * ndl + a] * dim + k];
      for (int k = 0; k < dim; ++k) cen[k] /= ndl;
      for (size_t k = 0; k < cell_keys.size() && k < 32; ++k) cell[k] = cell_data[k * ne + e];
      double val = model.eval_host(model.comp_expr[c], cen, 0.0, cell, 1.0, 0.0);
      if (std::fabs(val) > 1e-8 * std::max(1.0, std::fabs(val))) {  // FloatCmp::ne(val, 0.)
        if (elem_comp[e] >= 0) overlap = true;
        elem_comp[e] = c;
      }
    }
  }
  if (overlap) fail("overlapping compartments (a cell in several sub-domains) are out of scope");
  comp_nspec.assign(model.comp_nspec.begin(), model.comp_nspec.end());

  // ---- DOF map: compartments concatenated; vertex-major / species-minor inside a compartment;
  //      sub-domain vertices in ascending global vertex id
  comp_vertices.assign(ncomp, {});
  comp_vdof.assign(ncomp, {});
  comp_offset.assign(ncomp + 1, 0);
  for (int c = 0; c < ncomp; ++c) {
    std::vector<uint8_t> used(nv, 0);
    for (int64_t e = 0; e < ne; ++e)
      if (elem_comp[e] == c)
        for (int a = 0; a < ndl; ++a) used[elems[e * ndl + a]] = 1;
    auto& vd = comp_vdof[c];
    vd.assign(nv, -1);
    int64_t lv = 0;
    for (int64_t v = 0; v < nv; ++v)
      if (used[v]) {
        comp_vertices[c].push_back((int32_t)v);
        int64_t d = comp_offset[c] + lv * comp_nspec[c];
        if (d + comp_nspec[c] > INT32_MAX) fail("more than 2^31 dofs on one device");
        vd[v] = (int32_t)d;
        ++lv;
      }
    comp_offset[c + 1] = comp_offset[c] + lv * comp_nspec[c];
  }
  ndofs = comp_offset[ncomp];

  // ---- facets (only when something lives on them)
  bool need_facets = model.has_outflow();
  for (auto& s : model.species) need_facets |= !expr_is_absent(s.constrain_boundary);
  f_in.clear(); f_out.clear(); f_lin.clear(); f_lout.clear(); boundary_vertices.clear();
  if (!need_facets) return;
  if (elem_kind == 1) {
    // Q1 lattices: no facet terms; Dirichlet data lives on the outer vertices of the global lattice
    // (a slab's cut planes are not boundary)
    if (model.has_outflow()) fail("outflow / transmission terms on Q1 cube grids are out of scope");
    int64_t nvs[3] = {1, 1, 1};
    for (int a = 0; a < dim; ++a) nvs[a] = s_cells[a] + 1;
    const int L = dim - 1;
    for (int64_t v = 0; v < nv; ++v) {
      int64_t idx[3] = {v % nvs[0], (v / nvs[0]) % nvs[1], v / (nvs[0] * nvs[1])};
      bool onb = false;
      for (int a = 0; a < dim; ++a) {
        int64_t gi = idx[a] + (a == L ? s_layer_lo : 0), gn = a == L ? s_layers_global : s_cells[a];
        onb |= gi == 0 || gi == gn;
      }
      if (onb) boundary_vertices.push_back(v);
    }
    return;
  }
  std::vector<FaceRec> recs((size_t)ne * ndl);
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < ne; ++e)
    for (int a = 0; a < ndl; ++a) {
      FaceRec& r = recs[e * ndl + a];
      int n = 0;
      r.v[2] = -1;
      for (int b = 0; b < ndl; ++b)
        if (b != a) r.v[n++] = elems[e * ndl + b];
      std::sort(r.v, r.v + dim);
      r.l = a;
      r.e = e;
    }
  std::sort(recs.begin(), recs.end());
  struct F { int64_t in, out; int32_t lin, lout; };
  std::vector<F> fs;
  std::vector<uint8_t> isb(nv, 0);
  for (size_t i = 0; i < recs.size();) {
    if (i + 1 < recs.size() && recs[i].same(recs[i + 1])) {
      const FaceRec &a = recs[i], &b = recs[i + 1];   // a.e < b.e by the sort
      if (elem_comp[a.e] != elem_comp[b.e]) fs.push_back({a.e, b.e, a.l, b.l});
      i += 2;
    } else {
      // On a partitioned (local) grid a face without a local neighbour is a true boundary face
      // only if it has an owned vertex (then all elements around it are local); faces made of
      // ghosts only lie on the partition cut and feed ghost rows, which are never used.
      bool real = n_owned < 0;
      for (int k = 0; k < dim && !real; ++k) real = owns(recs[i].v[k]);
      if (real) {
        for (int k = 0; k < dim; ++k) isb[recs[i].v[k]] = 1;
        if (elem_comp[recs[i].e] >= 0) fs.push_back({recs[i].e, -1, recs[i].l, -1});
      }
      i += 1;
    }
  }
  std::sort(fs.begin(), fs.end(), [](const F& a, const F& b) { return a.in != b.in ? a.in < b.in : a.lin < b.lin; });
  for (auto& f : fs) {
    f_in.push_back(f.in); f_out.push_back(f.out); f_lin.push_back(f.lin); f_lout.push_back(f.lout);
  }
  for (int64_t v = 0; v < nv; ++v)
    if (isb[v]) boundary_vertices.push_back(v);
}

void Grid::pattern(const Model& model, std::vector<int64_t>& rowptr, std::vector<int32_t>& colidx) const {
  const int ndl = nd();
  // species coupling lists per species
  std::vector<std::vector<int>> act(model.nspec());
  for (auto& p : model.species_pairs()) act[p.first].push_back(model.species[p.second].local);
  // skeleton links (row, col), local_operator.hh:340-391
  std::vector<std::pair<int64_t, int32_t>> extra;
  for (auto& t : model.terms) {
    if (t.kind != Term::OutflowJac) continue;
    int ci = model.species[t.i].comp, ck = model.species[t.k].comp, l = t.j;
    for (size_t f = 0; f < f_in.size(); ++f) {
      for (int side = 0; side < 2; ++side) {
        int64_t e = side == 0 ? f_in[f] : f_out[f], eo = side == 0 ? f_out[f] : f_in[f];
        if (e < 0 || elem_comp[e] != ci) continue;
        int target = f_out[f] >= 0 ? elem_comp[eo] : ci;
        if (target != l) continue;
        int64_t ew = ck == ci ? e : ((f_out[f] >= 0 && elem_comp[eo] == ck) ? eo : -1);
        if (ew < 0) continue;
        int opp = side == 0 ? f_lin[f] : f_lout[f];
        for (int a = 0; a < ndl; ++a) {
          if (a == opp) continue;
          int64_t row = elem_dof(e, a) + model.species[t.i].local;
          for (int b = 0; b < ndl; ++b) {
            if (b == opp) continue;
            // the other element's vertex that carries phi_b of this element: the same vertex, or
            // (reference_compat, local_operator.hh:1133-1143) the one with the same local index
            int32_t gv = (model.reference_compat && ew != e) ? elems[ew * ndl + b] : elems[e * ndl + b];
            int32_t col = comp_vdof[ck][gv] + model.species[t.k].local;
            extra.push_back({row, col});
          }
        }
      }
    }
  }
  std::sort(extra.begin(), extra.end());
  // vertex -> incident elements (CSR)
  std::vector<int64_t> vptr(nv + 1, 0);
  for (int64_t e = 0; e < ne; ++e)
    if (elem_comp[e] >= 0)
      for (int a = 0; a < ndl; ++a) vptr[elems[e * ndl + a] + 1]++;
  for (int64_t v = 0; v < nv; ++v) vptr[v + 1] += vptr[v];
  std::vector<int64_t> vel(vptr[nv]);
  {
    std::vector<int64_t> cur(vptr.begin(), vptr.end() - 1);
    for (int64_t e = 0; e < ne; ++e)
      if (elem_comp[e] >= 0)
        for (int a = 0; a < ndl; ++a) vel[cur[elems[e * ndl + a]]++] = e;
  }
  // row of dof (c, v, i): two passes (count, fill)
  rowptr.assign(ndofs + 1, 0);
  auto row_cols = [&](int c, int32_t v, int i, std::vector<int32_t>& buf) {
    buf.clear();
    int g = model.comp_first[c] + i;
    int64_t row = comp_vdof[c][v] + i;
    if (!act[g].empty())
      for (int64_t p = vptr[v]; p < vptr[v + 1]; ++p) {
        int64_t e = vel[p];
        if (elem_comp[e] != c) continue;
        for (int b = 0; b < ndl; ++b) {
          int32_t base = comp_vdof[c][elems[e * ndl + b]];
          for (int j : act[g]) buf.push_back(base + j);
        }
      }
    auto lo = std::lower_bound(extra.begin(), extra.end(), std::make_pair(row, (int32_t)INT32_MIN));
    for (; lo != extra.end() && lo->first == row; ++lo) buf.push_back(lo->second);
    std::sort(buf.begin(), buf.end());
    buf.erase(std::unique(buf.begin(), buf.end()), buf.end());
  };
  for (int pass = 0; pass < 2; ++pass) {
    if (pass == 1) {
      for (int64_t r = 0; r < ndofs; ++r) rowptr[r + 1] += rowptr[r];
      colidx.resize(rowptr[ndofs]);
    }
    for (int c = 0; c < model.ncomp(); ++c) {
      const auto& verts = comp_vertices[c];
      int ns = comp_nspec[c];
#pragma omp parallel
      {
        std::vector<int32_t> buf;
#pragma omp for schedule(static)
        for (int64_t lv = 0; lv < (int64_t)verts.size(); ++lv)
          for (int i = 0; i < ns; ++i) {
            row_cols(c, verts[lv], i, buf);
            int64_t row = comp_vdof[c][verts[lv]] + i;
            if (pass == 0) rowptr[row + 1] = (int64_t)buf.size();
            else std::copy(buf.begin(), buf.end(), colidx.begin() + rowptr[row]);
          }
      }
    }
  }
}

void Grid::interpolate(const Model& model, double time, std::vector<double>& u) const {
  u.assign(ndofs, 0.0);
  for (int g = 0; g < model.nspec(); ++g) {
    const auto& s = model.species[g];
    if (expr_is_absent(s.initial)) continue;
    NodeP ast = model.compile(s.initial);
    const auto& verts = comp_vertices[s.comp];
    int ns = comp_nspec[s.comp];
#pragma omp parallel for schedule(static)
    for (int64_t lv = 0; lv < (int64_t)verts.size(); ++lv) {
      // P1: nodal evaluation (make_initial.hh:26-90, interpolate.hh:22-78)
      u[comp_offset[s.comp] + lv * ns + s.local] =
          model.eval_host(ast, &coords[(int64_t)verts[lv] * dim], time, nullptr, 1.0, 0.0);
    }
  }
}

// Dirichlet translation constraints (constraints.hh:114-195).  The reference walks the intersections and, per
// vertex of a face, keeps the value when the vertex is a boundary vertex exactly if the face is a boundary face
// (`in_boundary == isBoundary(vertex)`, :182): `constrain.boundary` therefore binds the boundary vertices,
// `constrain.skeleton` the vertices that are not on the boundary (each of them lies on a face with a neighbour).
// Restated per vertex: the expression sees the vertex position and in_boundary / in_skeleton; the face-dependent
// symbols (normal_*, entity_volume) read 0 -- with them the reference's value depends on which face is visited last.
void Grid::constraints(const Model& model, std::vector<int32_t>& dofs, std::vector<double>& vals) const {
  dofs.clear();
  vals.clear();
  std::vector<uint8_t> isb(nv, 0);
  for (auto v : boundary_vertices) isb[v] = 1;
  for (int g = 0; g < model.nspec(); ++g) {
    const auto& s = model.species[g];
    const bool hb = !expr_is_absent(s.constrain_boundary), hs = !expr_is_absent(s.constrain_skeleton);
    if (!hb && !hs) continue;
    NodeP ab = hb ? model.compile(s.constrain_boundary) : nullptr;
    NodeP as = hs ? model.compile(s.constrain_skeleton) : nullptr;
    const auto& verts = comp_vertices[s.comp];
    int ns = comp_nspec[s.comp];
    for (int64_t lv = 0; lv < (int64_t)verts.size(); ++lv) {
      const bool onb = isb[verts[lv]];
      const NodeP& ast = onb ? ab : as;
      if (!ast) continue;
      // time is NaN for constraints (constraints.hh:63-66); no_value == DBL_MAX => unconstrained
      double val = model.eval_host(ast, &coords[(int64_t)verts[lv] * dim], std::nan(""), nullptr, 0.0, onb ? 1.0 : 0.0,
                                   onb ? 0.0 : 1.0);
      if (val == DBL_MAX) continue;
      dofs.push_back((int32_t)(comp_offset[s.comp] + lv * ns + s.local));
      vals.push_back(val);
    }
  }
}


// ------------------------------------------------------------------------------------------------
// multi-GPU partition
// Recursive coordinate bisection of the vertices into `size` parts: split the longest
// axis of the bounding box at the weighted median, parts in proportion to the ranks on either side (any
// rank count), ties broken by the vertex id -- every rank computes the same map from the global mesh.
static void rcb_split(const std::vector<double>& coords, int dim, std::vector<int64_t>& ids, size_t b, size_t e,
                      int r0, int nparts, std::vector<int32_t>& owner) {
  if (nparts == 1) {
    for (size_t i = b; i < e; ++i) owner[ids[i]] = r0;
    return;
  }
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (size_t i = b; i < e; ++i)
    for (int k = 0; k < dim; ++k) {
      const double x = coords[ids[i] * dim + k];
      lo[k] = std::min(lo[k], x);
      hi[k] = std::max(hi[k], x);
    }
  int axis = 0;
  for (int k = 1; k < dim; ++k)
    if (hi[k] - lo[k] > hi[axis] - lo[axis]) axis = k;
  const int nl = nparts / 2;
  const size_t mid = b + (size_t)((__int128)(e - b) * nl / nparts);
  std::nth_element(ids.begin() + b, ids.begin() + mid, ids.begin() + e, [&](int64_t p, int64_t q) {
    const double xp = coords[p * dim + axis], xq = coords[q * dim + axis];
    return xp != xq ? xp < xq : p < q;
  });
  rcb_split(coords, dim, ids, b, mid, r0, nl, owner);
  rcb_split(coords, dim, ids, mid, e, r0 + nl, nparts - nl, owner);
}

Grid Grid::partition(int rank, int size, const std::string& method_in) const {
  if (size < 1 || rank < 0 || rank >= size) fail("partition: bad rank/size");
  const int ndl = nd();
  std::string method = method_in.empty() ? "auto" : method_in;
  if (method != "auto" && method != "slab" && method != "range" && method != "rcb")
    fail("partition: method must be auto, slab (structured lattices), range or rcb");
  const bool lattice = is_structured && global_vid.empty();
  if (method == "slab" && !lattice) fail("partition: slabs need a structured lattice");
  if (method == "auto") method = lattice ? "slab" : "rcb";
  if (method == "slab") {
    // slabs of vertex planes along the last axis
    const int L = dim - 1;
    const int64_t nplanes = s_cells[L] + 1;
    if (size > nplanes) fail("partition: more ranks than vertex planes");
    auto pbeg = [&](int r) { return nplanes * r / size; };
    const int64_t p0 = pbeg(rank), p1 = pbeg(rank + 1);
    const int lo = (int)std::max<int64_t>(p0 - 1, 0), hi = (int)std::min<int64_t>(p1, s_cells[L]);
    double extent[3];
    for (int a = 0; a < dim; ++a) extent[a] = s_h[a] * s_cells[a];
    // s_origin/extent of a global grid are the creation arguments
    Grid l = structured_box(dim, s_cells, s_origin_exact_, s_extent_exact_, lo, hi, elem_kind);
    int64_t plane = 1;
    for (int a = 0; a < L; ++a) plane *= s_cells[a] + 1;
    int64_t cubes_per_layer = 1;
    for (int a = 0; a < L; ++a) cubes_per_layer *= s_cells[a];
    l.global_vid.resize(l.nv);
    l.vowner.resize(l.nv);
    for (int64_t v = 0; v < l.nv; ++v) {
      l.global_vid[v] = v + (int64_t)lo * plane;
      int64_t pl = v / plane + lo;
      int r = (int)((pl + 1) * size / nplanes);
      if (r > size - 1) r = size - 1;
      while (r > 0 && pl < pbeg(r)) --r;
      while (r < size - 1 && pl >= pbeg(r + 1)) ++r;
      l.vowner[v] = r;
    }
    l.owned_begin = (p0 - lo) * plane;
    l.n_owned = (p1 - p0) * plane;
    l.global_eid.resize(l.ne);
    const int nperm = elem_kind == 1 ? 1 : dim == 2 ? 2 : 6;
    for (int64_t e = 0; e < l.ne; ++e) l.global_eid[e] = e + (int64_t)lo * cubes_per_layer * nperm;
    (void)extent;
    return l;
  }
  // vertex -> owning rank: contiguous ranges of the global numbering, or recursive coordinate bisection
  std::vector<int32_t> owner(nv, 0);
  if (method == "range") {
    auto vbeg = [&](int r) { return (int64_t)((__int128)nv * r / size); };
    for (int r = 0; r < size; ++r)
      for (int64_t v = vbeg(r); v < vbeg(r + 1); ++v) owner[v] = r;
  } else {
    if (coords.size() != (size_t)nv * dim) fail("partition: rcb needs vertex coordinates");
    std::vector<int64_t> ids(nv);
    std::iota(ids.begin(), ids.end(), (int64_t)0);
    rcb_split(coords, dim, ids, 0, (size_t)nv, 0, size, owner);
  }
  Grid l;
  l.dim = dim;
  l.elem_kind = elem_kind;
  l.cell_keys = cell_keys;
  // local elements: any vertex owned
  std::vector<int64_t> lel;
  for (int64_t e = 0; e < ne; ++e) {
    bool mine = false;
    for (int a = 0; a < ndl; ++a) mine |= owner[elems[e * ndl + a]] == rank;
    if (mine) lel.push_back(e);
  }
  // local vertices: owned range first, ghosts ascending
  std::vector<uint8_t> used(nv, 0);
  for (int64_t e : lel)
    for (int a = 0; a < ndl; ++a) used[elems[e * ndl + a]] = 1;
  std::vector<int32_t> g2l(nv, -1);
  int64_t nl = 0;
  for (int64_t v = 0; v < nv; ++v)
    if (owner[v] == rank) { g2l[v] = (int32_t)nl++; l.global_vid.push_back(v); }   // owned first, ascending global id
  l.n_owned = nl;
  l.owned_begin = 0;
  for (int64_t v = 0; v < nv; ++v)
    if (used[v] && owner[v] != rank) { g2l[v] = (int32_t)nl++; l.global_vid.push_back(v); }   // ghosts last
  l.nv = nl;
  l.ne = (int64_t)lel.size();
  l.coords.resize(nl * dim);
  l.vowner.resize(nl);
  for (int64_t i = 0; i < nl; ++i) {
    for (int k = 0; k < dim; ++k) l.coords[i * dim + k] = coords[l.global_vid[i] * dim + k];
    l.vowner[i] = owner[l.global_vid[i]];
  }
  l.elems.resize(l.ne * ndl);
  l.global_eid = lel;
  for (int64_t i = 0; i < l.ne; ++i)
    for (int a = 0; a < ndl; ++a) l.elems[i * ndl + a] = g2l[elems[lel[i] * ndl + a]];
  if (!cell_keys.empty()) {
    l.cell_data.resize(cell_keys.size() * (size_t)l.ne);
    for (size_t k = 0; k < cell_keys.size(); ++k)
      for (int64_t i = 0; i < l.ne; ++i) l.cell_data[k * l.ne + i] = cell_data[k * ne + lel[i]];
  }
  return l;
}

void Grid::owned_ranges(std::vector<int64_t>& begin, std::vector<int64_t>& end) const {
  begin.clear();
  end.clear();
  for (size_t c = 0; c < comp_vertices.size(); ++c) {
    int64_t first = 0, last = (int64_t)comp_vertices[c].size();
    if (n_owned >= 0) {
      first = std::lower_bound(comp_vertices[c].begin(), comp_vertices[c].end(), (int32_t)owned_begin) -
              comp_vertices[c].begin();
      last = std::lower_bound(comp_vertices[c].begin(), comp_vertices[c].end(), (int32_t)(owned_begin + n_owned)) -
             comp_vertices[c].begin();
    }
    begin.push_back(comp_offset[c] + first * comp_nspec[c]);
    end.push_back(comp_offset[c] + last * comp_nspec[c]);
  }
}

void Grid::halo_plan(int rank, std::vector<int>& peers, std::vector<std::vector<int32_t>>& send,
                     std::vector<std::vector<int32_t>>& recv) const {
  peers.clear(); send.clear(); recv.clear();
  if (n_owned < 0) return;
  const int ndl = nd();
  // (peer, compartment, global vertex) keys
  struct Key { int peer, comp; int64_t gv; int32_t lv; };
  auto less = [](const Key& a, const Key& b) {
    if (a.peer != b.peer) return a.peer < b.peer;
    if (a.comp != b.comp) return a.comp < b.comp;
    return a.gv < b.gv;
  };
  auto same = [](const Key& a, const Key& b) { return a.peer == b.peer && a.comp == b.comp && a.gv == b.gv; };
  std::vector<Key> s, r;
  for (int64_t e = 0; e < ne; ++e) {
    int c = elem_comp[e];
    if (c < 0 || comp_nspec[c] == 0) continue;
    for (int a = 0; a < ndl; ++a) {
      int32_t v = elems[e * ndl + a];
      if (vowner[v] == rank) {
        for (int b = 0; b < ndl; ++b) {
          int q = vowner[elems[e * ndl + b]];
          if (q != rank) s.push_back({q, c, global_vid[v], v});
        }
      } else {
        r.push_back({vowner[v], c, global_vid[v], v});
      }
    }
  }
  for (auto* k : {&s, &r}) {
    std::sort(k->begin(), k->end(), less);
    k->erase(std::unique(k->begin(), k->end(), same), k->end());
  }
  std::vector<int> ps;
  for (auto& k : s) ps.push_back(k.peer);
  for (auto& k : r) ps.push_back(k.peer);
  std::sort(ps.begin(), ps.end());
  ps.erase(std::unique(ps.begin(), ps.end()), ps.end());
  peers = ps;
  send.resize(ps.size());
  recv.resize(ps.size());
  auto fill = [&](const std::vector<Key>& keys, std::vector<std::vector<int32_t>>& out) {
    for (auto& k : keys) {
      size_t pi = std::lower_bound(ps.begin(), ps.end(), k.peer) - ps.begin();
      int32_t d = comp_vdof[k.comp][k.lv];
      for (int sp = 0; sp < comp_nspec[k.comp]; ++sp) out[pi].push_back(d + sp);
    }
  };
  fill(s, send);
  fill(r, recv);
}

}  // namespace dcb
