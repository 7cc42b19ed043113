// Expression front-end: one tolerant grammar covering what the reference's three parser back-ends
// (muParser / ExprTk / SymEngine, src/dune/copasi/parser/**) accept in the repo's inis
// (SURVEY.md App. D).  An expression is parsed once per model, parser_context constants and inline
// functions (src/dune/copasi/parser/context.cc:56-97) are folded in, and the tree is lowered to CUDA
// C text that NVRTC fuses into the assembly kernels; the same tree can be evaluated on the host
// for one-off setup work (compartment marking, initial values, Dirichlet values).
//
// Symbol table = what dune/copasi/model/functor_factory_parser.impl.hh:133-169 binds by address.
#pragma once
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "ptree.hpp"
#include "tiff.hpp"

namespace dcb {

struct Node;
using NodeP = std::shared_ptr<const Node>;

enum class Op {
  Num, Var, Neg, Not, Add, Sub, Mul, Div, Pow, Mod, Lt, Gt, Le, Ge, Eq, Ne, And, Or, Sel, Call
};

struct Node {
  Op op;
  double num = 0;
  std::string name;         // Var / Call
  std::vector<NodeP> kids;
};

// Tabulated one-argument context functions (src/dune/copasi/parser/context.cc):
//   kind 0  `type = interpolation` (:72-97): sorted `domain`, `range`; lower_bound, constant outside, std::lerp inside;
//   kind 1  `type = function` with `interpolate = true` (:237-283): the function sampled on `intervals` equal
//           intervals of `interpolation.domain.<arg>`, `interpolation.out_of_bounds = error | clamp`.  Restated
//           literally: the reference forms i = max(0, k) and j = min(k, intervals + 1) from the same interval
//           number k, i.e. it returns sample k (piecewise constant, lerp(g[k], g[k], t)); `clamp` is
//           std::clamp(domain[0], pos, domain[1]) = max(domain[0], pos) (arguments in that order).  Where the
//           reference would read past the table (pos > domain[1] under `clamp`) the last sample is used;
//           `error` yields NaN on the device (the reference throws) and throws on the host.
//   kind 2  `type = tiff` (:66-71): a grayscale image as a function of (x, y), nearest pixel (tiff.hpp).  Available to
//           every host-evaluated expression (initial values, constraints, compartments); device code embeds images
//           of up to kMaxDevicePixels pixels as constant arrays, larger ones fail loudly when a kernel needs them.
struct Table {
  static constexpr size_t kMaxDevicePixels = 1u << 18;
  TiffImage img;                       // kind 2
  int kind = 0;
  std::vector<double> domain, range;   // kind 1: domain = {d0, d1}, range = the intervals + 1 samples
  bool clamp = false;
  std::string name, key;               // context name; unique symbol used in trees and generated code
  double eval(double x, bool* out_of_bounds = nullptr) const;
};

struct ParserContext {
  std::map<std::string, double> constants;
  struct Fn { std::vector<std::string> args; std::string body; };
  std::map<std::string, Fn> functions;
  std::map<std::string, std::shared_ptr<const Table>> tables;   // context name -> table
  static ParserContext from_config(const PTree& parser_context);
  // CUDA C definitions of the tables (constant arrays + evaluation functions named after Table::key)
  std::string cuda_tables() const;
};

// "a, b: body" (ParserContext::parse_function_expression, src/dune/copasi/parser/context.cc)
ParserContext::Fn parse_function_expression(const std::string& text, const std::string& what);

// true when the expression is empty or a literal zero: the term does not exist
// (functor_factory_parser.impl.hh:122-124)
bool expr_is_absent(const std::string& text);

NodeP parse_expr(const std::string& text);
// tree construction for callers that combine parsed expressions
NodeP make_node(Op op, std::vector<NodeP> kids);
NodeP make_var(const std::string& name);
// inline context constants/functions, fold constants
NodeP resolve_expr(const NodeP& ast, const ParserContext& ctx);
bool is_constant(const NodeP& ast, double* value = nullptr);
void collect_vars(const NodeP& ast, std::vector<std::string>& out);

// d(ast)/d(var) of a resolved expression, simplified (0 + x, 1 * x, constant folding ...).  Comparisons,
// logical operators, floor/ceil/round/sgn are piecewise constant: their derivative is 0 and the
// condition of a selection is kept as it is (derivative of the taken branch).  This is what north_star
// calls "analytic Jacobians from SymEngine": the reference itself takes its Jacobian entries from the
// ini (local_equations.hh:553-579); model.jacobian.type = symbolic derives them instead.
NodeP differentiate(const NodeP& ast, const std::string& var);
bool is_zero(const NodeP& ast);
// infix text of a tree (diagnostics, generated-source comments)
std::string to_text(const NodeP& ast);

// host evaluation; `lookup` maps a variable name to its value (throws for unknown names)
double eval_expr(const NodeP& ast, const std::function<double(const std::string&)>& lookup);

// Polynomial normal form of a set of expressions over common atoms (variables and every
// sub-expression that is not a sum / product / small integer power / division by a constant).  The
// generated kernels evaluate the reaction terms at every quadrature point of every element and the
// fp64 pipe is what bounds them: in this form an entry costs one product per distinct monomial
// (shared by all entries of a function) and one fused multiply-add per term, with the Runge-Kutta
// weights folded into the coefficients.  Same value up to the rounding of the reassociated products.
struct PolyForm {
  std::vector<NodeP> atoms;                                    // atom id -> sub-expression
  std::vector<std::map<std::vector<int>, double>> polys;       // per expression: monomial (sorted atom ids) -> coefficient
};
// false when an expansion exceeds max_terms terms or max_degree factors (callers then emit the tree as it is)
bool expand_polynomials(const std::vector<NodeP>& exprs, PolyForm& out, size_t max_terms = 24, size_t max_degree = 6);

// CUDA C text; `symbol` maps a variable name to a C expression (returns "" for unknown -> error)
std::string to_cuda(const NodeP& ast, const std::function<std::string(const std::string&)>& symbol);

}  // namespace dcb
