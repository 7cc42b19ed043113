// Host-side mesh, compartments, interface facets, P1 DOF map and sparsity pattern.
//
// Replaces (for the hot path) what the reference obtains from dune-grid/multidomaingrid + PDELab's
// basis: grid/make_multi_domain_grid.hh:76-100 (structured simplex grid), :118-194 (compartment
// marking by expression on cell centres), model_*_compartment_traits.hh (EntityGrouping /
// Lexicographic DOF order), local_operator.hh:276-399 + make_step_operator.hh:378-384 (pattern).
// Numbering rules are written down in DESIGN.md and restated independently in oracle/mesh.py.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "model.hpp"

namespace dcb {

struct Grid {
  int dim = 0;
  int64_t nv = 0, ne = 0;
  std::vector<double> coords;            // [nv*dim]
  std::vector<int32_t> elems;            // [ne*nd()]
  // 0: simplices (the reference's element type); 1: the cells of a structured lattice as Q1 cubes,
  // corner m of a cell at the bit pattern of m (x = bit 0).  Q1 is BASELINE configs[3]'s element
  // and NOT a reference capability (PkLocalFiniteElementMap is simplex-only): cube
  // grids run on the implicit-geometry kernels only (kernels/assembly_q1.cuh).
  int elem_kind = 0;
  std::vector<std::string> cell_keys;
  std::vector<double> cell_data;         // [nkeys*ne]

  // ---- filled by bind(model)
  std::vector<int32_t> elem_comp;        // [ne] compartment id or -1
  std::vector<int64_t> f_in, f_out;      // interface + boundary facets
  std::vector<int32_t> f_lin, f_lout;    // local index of the vertex opposite to the facet
  std::vector<int64_t> boundary_vertices;
  std::vector<int32_t> comp_nspec;
  std::vector<std::vector<int32_t>> comp_vertices;  // sorted global vertex ids per compartment
  std::vector<std::vector<int32_t>> comp_vdof;      // [nv] dof of species 0 at the vertex, or -1
  std::vector<int64_t> comp_offset;
  int64_t ndofs = 0;

  // ---- structured simplex grids keep their lattice (implicit-geometry kernels)
  bool is_structured = false;
  int s_cells[3] = {1, 1, 1};            // cells of this (local) box
  double s_origin[3] = {0, 0, 0}, s_h[3] = {1, 1, 1};
  double s_origin_exact_[3] = {0, 0, 0}, s_extent_exact_[3] = {1, 1, 1};   // creation arguments (global grid)
  int s_layer_lo = 0, s_layers_global = 0;   // first cell layer of this box / cell layers of the global lattice (last axis)

  // ---- partition data (local grids produced by partition(); empty on a global grid)
  int64_t n_owned = -1;                  // number of owned vertices (-1: not a partitioned grid)
  int64_t owned_begin = 0;               // owned vertices are the local ids [owned_begin, owned_begin + n_owned)
  bool owns(int64_t v) const { return n_owned < 0 || (v >= owned_begin && v < owned_begin + n_owned); }
  std::vector<int64_t> global_vid;       // [nv] global vertex id
  std::vector<int32_t> vowner;           // [nv] owning rank
  std::vector<int64_t> global_eid;       // [ne] global element id

  // Vertex partition: every vertex has one owning rank; a rank's local mesh holds every
  // element touching an owned vertex (owner computes, one layer of ghosts), vertices renumbered owned
  // first, ghosts last (each group ascending in global id).  Methods:
  //   slab   structured lattices: slabs of vertex planes along the last axis, so that every local mesh is
  //          again a structured box (owned planes in the middle, one ghost plane per side)
  //   rcb    recursive coordinate bisection of the vertex coordinates (any rank count)
  //   range  contiguous ranges of the global vertex numbering
  //   auto   slab for structured lattices, rcb otherwise
  Grid partition(int rank, int size, const std::string& method = "auto") const;
  static Grid structured_box(int dim, const int* cells, const double* origin, const double* extent,
                             int layer_lo, int layer_hi, int elem_kind = 0);
  // owned local dof ranges per compartment (valid after bind)
  void owned_ranges(std::vector<int64_t>& begin, std::vector<int64_t>& end) const;
  // halo plan of a bound local grid: per peer the local dofs to send / receive, both ordered by
  // (compartment, global vertex id, species) so the two sides agree without communication
  void halo_plan(int rank, std::vector<int>& peers, std::vector<std::vector<int32_t>>& send,
                 std::vector<std::vector<int32_t>>& recv) const;

  static Grid structured(int dim, const int* cells, const double* origin, const double* extent, int elem_kind = 0);
  static Grid from_arrays(int dim, int64_t nv, const double* coords, int64_t ne, const int32_t* elems,
                          const std::vector<std::string>& keys, const double* cell_data);

  void bind(const Model& model);
  int nd() const { return elem_kind == 1 ? 1 << dim : dim + 1; }
  int64_t elem_dof(int64_t e, int a) const {   // dof of species 0 at local vertex a of element e
    int c = elem_comp[e];
    return c < 0 ? -1 : comp_vdof[c][elems[e * nd() + a]];
  }

  // sorted CSR pattern of the Jacobian (rows = dofs)
  void pattern(const Model& model, std::vector<int64_t>& rowptr, std::vector<int32_t>& colidx) const;
  // initial values / Dirichlet data at the dofs
  void interpolate(const Model& model, double time, std::vector<double>& u) const;
  void constraints(const Model& model, std::vector<int32_t>& dofs, std::vector<double>& vals) const;
};

// VTK output (vtk.cpp): "<path>/<stem>-<compartment>-<00000>.vtu" + "<path>/<stem>-<compartment>.pvd";
// `timesteps` holds the stamps already written under `path` and is cleared unless `append`
void write_vtk(const Grid& grid, const Model& model, const double* u, double time, const std::string& path,
               bool append, std::vector<double>& timesteps);

}  // namespace dcb
