// Model description: what dune/copasi/model/diffusion_reaction/local_equations.hh:617-700 builds
// from the ini ([compartments], [parser_context], [model.scalar_field.*]) -- which terms exist per
// species -- plus the lowering of those terms to CUDA source that NVRTC fuses into the assembly
// kernels (replaces the per-quadrature-point type-erased functor calls of
// local_equations.hh:101-146 / functor_factory_parser.impl.hh:116-182).
#pragma once
#include <string>
#include <vector>

#include "expr.hpp"
#include "ptree.hpp"

namespace dcb {

struct Term {
  // Diff / DiffJac: scalar diffusion D_ij and dD_ij/du_k; DiffT / DiffTJac: one entry [r][c] of a
  // tensor diffusion and of its derivative (make_tensor_apply, functor_factory_parser.impl.hh:65-112);
  // Vel / VelJac: one component of the advection velocity and of its derivative (make_vector :36-61)
  enum Kind { Reaction, ReactionJac, Storage, StorageJac, Diff, DiffJac, Outflow, OutflowJac, Vel, VelJac, DiffT, DiffTJac };
  Kind kind;
  int i = -1;    // species (global id)
  int j = -1;    // wrt species (Jac / Diff*) or target compartment (Outflow*) or axis (Vel)
  int k = -1;    // jac wrt species of DiffJac / OutflowJac; axis of VelJac; 3r+c of DiffT; 9k+3r+c of DiffTJac
  NodeP ast;     // resolved + folded
  std::string text;
};

struct SpeciesInfo {
  std::string name;
  int comp = -1, local = -1;
  std::string initial, constrain_boundary, constrain_skeleton;   // raw expressions ("" if absent)
};

class Model {
 public:
  // dim and the cell-data keys are part of the symbol table
  Model(const PTree& cfg, int dim, const std::vector<std::string>& cell_keys);

  int dim;
  bool is_linear = false;
  // model.jacobian.type = numerical: finite differences with model.jacobian.epsilon
  // (local_operator.hh:193-202, 713-765); linear operators always use the analytic path (:234-235)
  bool numerical_jacobian = false;
  // model.b200.reference_compat (default true): facet terms exactly as local_operator.hh:903-916 / :1298 compute
  // them (cross-side coefficients paired with this side's shape functions by local index)
  bool reference_compat = true;
  bool blocked_scalar_fields = false, blocked_compartments = false;   // container nesting only (model.cpp)
  double fd_epsilon = 1e-7;
  // model.jacobian.type = symbolic (extension; north_star: "analytic Jacobians from SymEngine"): every
  // jacobian entry is derived from its function by expr.cpp's differentiator, the ini's own
  // `jacobian.*` sub-sections are not read.  The reference only knows user-supplied entries
  // (local_equations.hh:553-579).
  bool symbolic_jacobian = false;
  std::vector<std::string> cell_keys;
  std::vector<std::string> comp_names;
  std::vector<NodeP> comp_expr;
  std::vector<SpeciesInfo> species;           // compartment-major
  std::vector<int> comp_nspec, comp_first;    // per compartment
  std::vector<Term> terms;
  ParserContext ctx;
  PTree cfg;

  int ncomp() const { return (int)comp_names.size(); }
  int nspec() const { return (int)species.size(); }
  int species_index(const std::string& name) const;
  bool has_outflow() const;
  // no diffusion coefficient of compartment c depends on the quadrature point (position, fields)
  bool diffusion_is_constant(int c) const;
  // compartment c carries advection, tensor diffusion or diffusion-Jacobian terms: they are
  // assembled by the general element kernels only (experimental in the reference, CHANGELOG !83)
  bool has_extended_terms(int c) const;
  // species couplings (i,j) of the volume sparsity pattern, local_operator.hh:276-338
  std::vector<std::pair<int, int>> species_pairs() const;
  // directional compartment pairs (cs -> ct) that carry an outflow term; ct == cs means boundary
  std::vector<std::pair<int, int>> outflow_pairs() const;

  // host evaluation of a resolved expression with position/time/cell data bound (setup work only)
  double eval_host(const NodeP& ast, const double* pos, double time, const double* cell,
                   double in_volume, double in_boundary, double in_skeleton = 0.0) const;
  NodeP compile(const std::string& text) const;   // parse + resolve against the context

  // CUDA source of the per-model device functions (see kernels/assembly.cuh for the consumers)
  std::string cuda_source() const;
  // CUDA text of one resolved expression in the volume context of compartment c (symbols `c`, `u`,
  // `g` as in DcComp<C>::scalar); species without support there read 0, c outside [0, ncomp) = no compartment
  std::string lower_volume(const NodeP& ast, int c) const;
};

}  // namespace dcb
