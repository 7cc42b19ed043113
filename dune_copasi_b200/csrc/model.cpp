#include "model.hpp"

#include <functional>

#include <algorithm>
#include <cmath>
#include <set>
#include <sstream>

namespace dcb {

static const char* kAxis[3] = {"x", "y", "z"};

Model::Model(const PTree& cfg_, int dim_, const std::vector<std::string>& keys)
    : dim(dim_), cell_keys(keys) {
  cfg.parse_ini(cfg_.dump());
  if (dim < 2 || dim > 3) fail("only dimensions 2 and 3 are built (got ", dim, ")");
  ctx = ParserContext::from_config(cfg.sub("parser_context"));
  const PTree& mcfg = cfg.sub("model");
  is_linear = mcfg.get("is_linear", false);
  if (mcfg.get("order", 1) != 1) fail("model.order = ", mcfg.get("order", 1), ": only P1 is built");
  {
    std::string jt = mcfg.get("jacobian.type", std::string("analytical"));
    if (jt != "analytical" && jt != "numerical" && jt != "symbolic")
      fail("The option 'model.jacobian.type' must be either 'analytical' or 'numerical' (or the extension 'symbolic')");
    numerical_jacobian = jt == "numerical" && !is_linear;
    symbolic_jacobian = jt == "symbolic";
    fd_epsilon = mcfg.get("jacobian.epsilon", 1e-7);
    reference_compat = mcfg.get("b200.reference_compat", true);
    // model.blocked_layout.{scalar_fields, compartments} (factory.hh:74-75, 96-131) select the *nesting* of the dune-istl
    // containers -- PDELab::EntityGrouping<ES, Blocked> per vertex, PDELab::Lexicographic<Blocked> over the
    // compartments (model_single_compartment_traits.hh:26-27, model_multi_compartment_traits.hh:22) -- not the order of
    // the scalars: compartment-major, vertex, species in all four combinations.  The C ABI takes flat arrays in that
    // order, so the flags are read and change nothing here (the DUNE-side shim copies between the nested container
    // and the flat array either way).
    blocked_scalar_fields = mcfg.get("blocked_layout.scalar_fields", false);
    blocked_compartments = mcfg.get("blocked_layout.compartments", false);

  }
  const PTree& comps = cfg.sub("compartments");
  for (auto& name : comps.sub_keys()) {
    const PTree& c = comps.sub(name);
    std::string type = c.get("type", std::string("expression"));
    if (type != "expression") fail("compartments.", name, ".type = '", type, "' is not known");
    comp_names.push_back(name);
    comp_expr.push_back(compile(c.get("expression", std::string("0"))));
  }
  if (comp_names.empty()) fail("config has no [compartments]");
  const PTree& fields = mcfg.sub("scalar_field");
  comp_nspec.assign(ncomp(), 0);
  comp_first.assign(ncomp(), 0);
  std::vector<const PTree*> scfg;
  for (int c = 0; c < ncomp(); ++c) {
    comp_first[c] = nspec();
    for (auto& name : fields.sub_keys()) {
      const PTree& f = fields.sub(name);
      if (f.get("compartment", std::string()) != comp_names[c]) continue;
      SpeciesInfo s;
      s.name = name;
      s.comp = c;
      s.local = comp_nspec[c]++;
      s.initial = f.get("initial.expression", std::string());
      s.constrain_boundary = f.get("constrain.boundary.expression", std::string());
      s.constrain_skeleton = f.get("constrain.skeleton.expression", std::string());
      // constrain.volume binds the dofs attached to the cell itself (codim 0, constraints.hh:93-112): P1 / Q1
      // elements have none, so the section is read and has no effect -- as in the reference
      species.push_back(s);
      scfg.push_back(&f);
    }
  }
  if (species.empty())
    fail("Basis has dimension 0, make sure to have at least one 'scalar_field' with a non-empty 'compartment'");
  {
    std::set<std::string> seen;
    for (auto& s : species)
      if (!seen.insert(s.name).second) fail("Variable with name '", s.name, "' is repeated");
  }
  auto add = [&](Term::Kind kind, int i, int j, int k, const std::string& text) {
    if (expr_is_absent(text)) return false;
    Term t;
    t.kind = kind; t.i = i; t.j = j; t.k = k; t.text = text;
    t.ast = compile(text);
    terms.push_back(t);
    return true;
  };
  // symbolic mode: d(term)/d(species k) for every species k of the given compartments, kept when it
  // is not identically zero
  auto derive = [&](Term::Kind jkind, const Term& fn, std::vector<int> comps,
                    const std::function<void(Term&, int)>& place) {
    std::sort(comps.begin(), comps.end());
    comps.erase(std::unique(comps.begin(), comps.end()), comps.end());
    for (int c : comps)
      for (int k = comp_first[c]; k < comp_first[c] + comp_nspec[c]; ++k) {
        NodeP d = differentiate(fn.ast, species[k].name);
        if (is_zero(d)) continue;
        Term t;
        t.kind = jkind; t.i = fn.i; t.ast = d;
        t.text = to_text(d);
        place(t, k);
        terms.push_back(t);
      }
  };
  for (int g = 0; g < nspec(); ++g) {
    const PTree& f = *scfg[g];
    const int cg = species[g].comp;
    {
      // velocity.<axis>.expression, velocity.jacobian.<wrt>.<axis>.expression (local_equations.hh:663-669)
      const PTree& v = f.sub("velocity");
      bool active = false;
      for (int a = 0; a < dim; ++a) active |= add(Term::Vel, g, a, -1, v.sub(kAxis[a]).get("expression", std::string()));
      if (active && symbolic_jacobian) {
        std::vector<Term> fns;
        for (auto& t : terms)
          if (t.kind == Term::Vel && t.i == g) fns.push_back(t);
        for (auto& fn : fns)
          derive(Term::VelJac, fn, {cg}, [&](Term& t, int k) { t.j = k; t.k = fn.j; });
      } else if (active)
        for (auto& wrt : v.sub("jacobian").sub_keys()) {
          int k = species_index(wrt);
          if (k < 0) continue;
          for (int a = 0; a < dim; ++a)
            add(Term::VelJac, g, k, a, v.sub("jacobian").sub(wrt).sub(kAxis[a]).get("expression", std::string()));
        }
    }
    struct { Term::Kind k, jk; const char* key; } two[] = {
        {Term::Reaction, Term::ReactionJac, "reaction"}, {Term::Storage, Term::StorageJac, "storage"}};
    for (auto& tk : two) {
      const PTree& t = f.sub(tk.key);
      if (!add(tk.k, g, -1, -1, t.get("expression", std::string()))) continue;
      if (symbolic_jacobian) {
        const Term fn = terms.back();
        derive(tk.jk, fn, {cg}, [&](Term& d, int k) { d.j = k; d.k = -1; });
      } else
        for (auto& wrt : t.sub("jacobian").sub_keys()) {
          int j = species_index(wrt);
          if (j >= 0) add(tk.jk, g, j, -1, t.sub("jacobian").sub(wrt).get("expression", std::string()));
        }
    }
    const PTree& cd = f.sub("cross_diffusion");
    for (auto& wrt : cd.sub_keys()) {
      int j = species_index(wrt);
      if (j < 0) continue;
      const PTree& d = cd.sub(wrt);
      // `type` = scalar | tensor is read for the function and again for each jacobian entry
      auto diffusion = [&](const PTree& t, Term::Kind ks, Term::Kind kt, int kk) {
        std::string type = t.get("type", std::string("scalar"));
        if (type == "scalar") return add(ks, g, j, kk, t.get("expression", std::string()));
        if (type != "tensor") fail("not known type 'scalar_value.", species[g].name, ".cross_diffusion.type = ", type, "'");
        bool active = false;
        for (int r = 0; r < dim; ++r)
          for (int c = 0; c < dim; ++c) {
            std::string key = std::string(kAxis[r]) + kAxis[c];
            if (t.has_sub(key))
              active |= add(kt, g, j, 3 * r + c + (kk >= 0 ? 9 * kk : 0), t.sub(key).get("expression", std::string()));
          }
        return active;
      };
      const size_t first_new = terms.size();
      const bool have = diffusion(d, Term::Diff, Term::DiffT, -1);
      if (have && symbolic_jacobian) {
        std::vector<Term> fns(terms.begin() + first_new, terms.end());
        for (auto& fn : fns)
          derive(fn.kind == Term::Diff ? Term::DiffJac : Term::DiffTJac, fn, {cg}, [&](Term& t, int k) {
            t.j = fn.j;
            t.k = fn.kind == Term::Diff ? k : fn.k + 9 * k;
          });
      } else if (have)
        for (auto& kk : d.sub("jacobian").sub_keys()) {
          int k = species_index(kk);
          if (k >= 0) diffusion(d.sub("jacobian").sub(kk), Term::DiffJac, Term::DiffTJac, k);
        }
    }
    const PTree& of = f.sub("outflow");
    for (auto& cname : of.sub_keys()) {
      auto it = std::find(comp_names.begin(), comp_names.end(), cname);
      if (it == comp_names.end()) continue;
      int l = (int)(it - comp_names.begin());
      const PTree& o = of.sub(cname);
      if (!add(Term::Outflow, g, l, -1, o.get("expression", std::string()))) continue;
      if (symbolic_jacobian) {
        const Term fn = terms.back();
        derive(Term::OutflowJac, fn, {cg, l}, [&](Term& t, int k) { t.j = l; t.k = k; });
      } else
        for (auto& kk : o.sub("jacobian").sub_keys()) {
          int k = species_index(kk);
          if (k >= 0) add(Term::OutflowJac, g, l, k, o.sub("jacobian").sub(kk).get("expression", std::string()));
        }
    }
  }
}

int Model::species_index(const std::string& name) const {
  for (int g = 0; g < nspec(); ++g)
    if (species[g].name == name) return g;
  return -1;
}

bool Model::has_outflow() const {
  for (auto& t : terms)
    if (t.kind == Term::Outflow) return true;
  return false;
}

static bool depends_on_point(const NodeP& ast);

bool Model::has_extended_terms(int c) const {
  for (auto& t : terms)
    if (species[t.i].comp == c && (t.kind == Term::Vel || t.kind == Term::VelJac || t.kind == Term::DiffT ||
                                   t.kind == Term::DiffTJac || t.kind == Term::DiffJac))
      return true;
  return false;
}

bool Model::diffusion_is_constant(int c) const {
  if (has_extended_terms(c)) return false;
  for (auto& t : terms)
    if (t.kind == Term::Diff && species[t.i].comp == c && species[t.j].comp == c && depends_on_point(t.ast)) return false;
  return true;
}

std::vector<std::pair<int, int>> Model::species_pairs() const {
  std::set<std::pair<int, int>> s;
  for (auto& t : terms) {
    if (t.kind == Term::ReactionJac || t.kind == Term::StorageJac || t.kind == Term::Diff || t.kind == Term::DiffT ||
        t.kind == Term::VelJac)
      s.insert({t.i, t.j});
    else if (t.kind == Term::Storage || t.kind == Term::Vel) s.insert({t.i, t.i});   // velocity: :304-307
    else if (t.kind == Term::DiffJac) s.insert({t.i, t.k});
    else if (t.kind == Term::DiffTJac) s.insert({t.i, t.k / 9});
  }
  std::vector<std::pair<int, int>> out;
  for (auto& p : s)
    if (species[p.first].comp == species[p.second].comp) out.push_back(p);
  return out;
}

std::vector<std::pair<int, int>> Model::outflow_pairs() const {
  std::set<std::pair<int, int>> s;
  for (auto& t : terms)
    if (t.kind == Term::Outflow) s.insert({species[t.i].comp, t.j});
  return {s.begin(), s.end()};
}

NodeP Model::compile(const std::string& text) const { return resolve_expr(parse_expr(text), ctx); }

double Model::eval_host(const NodeP& ast, const double* pos, double time, const double* cell,
                        double in_volume, double in_boundary, double in_skeleton) const {
  return eval_expr(ast, [&](const std::string& n) -> double {
    if (n == "time") return time;
    if (n == "in_volume") return in_volume;
    if (n == "in_boundary") return in_boundary;
    if (n == "in_skeleton") return in_skeleton;
    if (n == "integration_factor" || n == "entity_volume") return 0.0;
    for (int a = 0; a < 3; ++a) {
      if (n == std::string("position_") + kAxis[a]) return a < dim ? pos[a] : 0.0;
      if (n == std::string("normal_") + kAxis[a]) return 0.0;
    }
    for (size_t k = 0; k < cell_keys.size(); ++k)
      if (n == cell_keys[k]) return cell ? cell[k] : 0.0;
    fail("unknown symbol '", n, "' in a setup expression");
  });
}

// ------------------------------------------------------------------------------------------------
// CUDA lowering.  One struct per compartment (volume terms) and one per directional outflow pair.
namespace {

struct SymbolMap {
  const Model& m;
  int cs, ct;          // own compartment; other-side compartment (-1: volume context)
  bool codim1;
  std::string operator()(const std::string& n) const {
    if (n == "time") return "c.time";
    if (n == "integration_factor") return "c.integration_factor";
    if (n == "entity_volume") return "c.entity_volume";
    if (n == "in_volume") return "c.in_volume";
    if (n == "in_boundary") return "c.in_boundary";
    if (n == "in_skeleton") return "c.in_skeleton";
    for (int a = 0; a < 3; ++a) {
      if (n == std::string("position_") + kAxis[a]) return a < m.dim ? "c.pos[" + std::to_string(a) + "]" : "0.0";
      if (n == std::string("normal_") + kAxis[a]) {
        if (!codim1) return "";
        return a < m.dim ? "c.nrm[" + std::to_string(a) + "]" : "0.0";
      }
    }
    for (size_t k = 0; k < m.cell_keys.size(); ++k)
      if (n == m.cell_keys[k]) return "c.cell[" + std::to_string(k) + "]";
    for (int g = 0; g < m.nspec(); ++g) {
      const auto& s = m.species[g];
      auto side = [&](const char* u) { return std::string(u) + "[" + std::to_string(s.local) + "]"; };
      if (n == s.name) {
        if (s.comp == cs) return side(codim1 ? "us" : "u");
        if (codim1 && s.comp == ct) return side("ut");
        return "0.0";  // species without support here: value stays 0 (LocalEquations::clear)
      }
      for (int a = 0; a < 3; ++a)
        if (n == "grad_" + s.name + "_" + kAxis[a]) {
          if (a >= m.dim) return "0.0";
          std::string ax = "[" + std::to_string(a) + "]";
          if (s.comp == cs) return side(codim1 ? "gs" : "g") + ax;
          if (codim1 && s.comp == ct) return side("gt") + ax;
          return "0.0";
        }
    }
    return "";
  }
};

}  // namespace

std::string Model::lower_volume(const NodeP& ast, int c) const {
  SymbolMap sym{*this, (c >= 0 && c < ncomp()) ? c : -2, -1, false};
  return to_cuda(ast, sym);
}

static bool depends_on_point(const NodeP& ast) {
  std::vector<std::string> v;
  collect_vars(ast, v);
  for (auto& n : v)
    if (n != "time" && n != "entity_volume" && n != "in_volume" && n != "in_boundary" && n != "in_skeleton")
      return true;   // position, cell data are per element but harmless; species values are per point
  return false;
}



namespace {
// One point function of the generated kernels: lhs = wA * a + wM * m for every entry, in polynomial
// normal form over common atoms (see expand_polynomials).  Returns false when an expression does
// not expand within the limits; nothing has been written then.
struct WeightedEntry { std::string lhs; NodeP a, m; };

bool emit_polynomial_entries(std::ostream& o, const std::vector<WeightedEntry>& entries,
                             const std::function<std::string(const std::string&)>& sym) {
  std::vector<NodeP> exprs;
  for (auto& e : entries) {
    if (e.a) exprs.push_back(e.a);
    if (e.m) exprs.push_back(e.m);
  }
  PolyForm pf;
  if (!expand_polynomials(exprs, pf)) return false;
  std::ostringstream out;
  char buf[64];
  auto lit = [&](double v) { snprintf(buf, sizeof buf, "%.17g", v); std::string t = buf;
                             if (t.find_first_of(".eEn") == std::string::npos) t += ".0"; return t; };
  for (size_t k = 0; k < pf.atoms.size(); ++k) out << "    const double q" << k << " = " << to_cuda(pf.atoms[k], sym) << ";\n";
  // products: every monomial of degree >= 2 and its prefixes, shortest first
  std::map<std::vector<int>, std::string> name;
  std::vector<std::vector<int>> order;
  std::function<void(const std::vector<int>&)> need = [&](const std::vector<int>& m) {
    if (m.size() < 2 || name.count(m)) return;
    std::vector<int> prefix(m.begin(), m.end() - 1);
    need(prefix);
    name[m] = "m" + std::to_string(order.size());
    order.push_back(m);
  };
  for (auto& p : pf.polys)
    for (auto& t : p) need(t.first);
  auto mono_name = [&](const std::vector<int>& m) { return m.size() == 1 ? "q" + std::to_string(m[0]) : name.at(m); };
  for (auto& m : order) {
    std::vector<int> prefix(m.begin(), m.end() - 1);
    out << "    const double " << name[m] << " = " << mono_name(prefix) << " * q" << m.back() << ";\n";
  }
  size_t at = 0;
  for (auto& e : entries) {
    // coefficient of a monomial: the weighted sum over the two forms, e.g. (wA * 0.042 + wM)
    std::map<std::vector<int>, std::string> coef;
    auto take = [&](const std::map<std::vector<int>, double>& p, const char* wname) {
      for (auto& t : p) {
        const std::string c = t.second == 1.0 ? std::string(wname) : t.second == -1.0 ? "-" + std::string(wname)
                                                                                      : std::string(wname) + " * " + lit(t.second);
        std::string& dst = coef[t.first];
        dst += (dst.empty() ? "" : " + ") + c;
      }
    };
    if (e.a) take(pf.polys[at++], "wA");
    if (e.m) take(pf.polys[at++], "wM");
    if (coef.empty()) { out << "    " << e.lhs << " = 0.0;\n"; continue; }
    std::string text;
    for (auto& kv : coef)
      text += (text.empty() ? "" : " + ") + ("(" + kv.second + ")") + (kv.first.empty() ? "" : " * " + mono_name(kv.first));
    out << "    " << e.lhs << " = " << text << ";\n";
  }
  o << out.str();
  return true;
}
}  // namespace

std::string Model::cuda_source() const {
  std::ostringstream o;
  o << "// generated by dune_copasi_b200 Model::cuda_source()\n";
  o << "#define DC_DIM " << dim << "\n#define DC_NKEYS " << cell_keys.size() << "\n#define DC_NCOMP " << ncomp() << "\n";
  {
    char eps[64];
    snprintf(eps, sizeof eps, "%.17g", fd_epsilon);
    o << "#define DC_NUMJAC " << (numerical_jacobian ? 1 : 0) << "\n#define DC_FD_EPS " << eps << "\n";
    o << "#define DC_REF_COMPAT " << (reference_compat ? 1 : 0) << "\n";
  }
  o << "struct DcCtx { double time, entity_volume, integration_factor, in_volume, in_boundary, in_skeleton;"
       " double pos[3]; double nrm[3]; double cell[" << std::max<size_t>(1, cell_keys.size()) << "]; };\n";
  o << "template <int N> __device__ __forceinline__ double dc_powi(double x) { double r = x;\n"
       "#pragma unroll\n  for (int i = 1; i < N; ++i) r *= x; return r; }\n";
  o << "__device__ __forceinline__ double dc_sgn(double x) { return (double)((x > 0.0) - (x < 0.0)); }\n";
  o << "__device__ __forceinline__ double dc_min(double a, double b) { return a < b ? a : b; }\n";
  o << "__device__ __forceinline__ double dc_max(double a, double b) { return a > b ? a : b; }\n";
  o << ctx.cuda_tables();   // tabulated parser_context functions (expr.hpp)
  o << "template <int C> struct DcComp;\ntemplate <int P> struct DcOutflow;\n";
  auto pairs = species_pairs();
  for (int c = 0; c < ncomp(); ++c) {
    int ns = comp_nspec[c], g0 = comp_first[c];
    SymbolMap sym{*this, c, -1, false};
    auto code = [&](const Term& t) { return to_cuda(t.ast, sym); };
    auto find = [&](Term::Kind k, int i) -> const Term* {
      for (auto& t : terms) if (t.kind == k && t.i == i) return &t;
      return nullptr;
    };
    bool has_mass = false, has_stiff = false, has_diff = false, diff_const = true;
    for (auto& t : terms) {
      if (species[t.i].comp != c) continue;
      if (t.kind == Term::Storage) has_mass = true;
      if (t.kind == Term::Reaction) has_stiff = true;
      if (t.kind == Term::Diff && species[t.j].comp == c) {
        has_stiff = has_diff = true;
        if (depends_on_point(t.ast)) diff_const = false;
      }
      if (t.kind == Term::Vel || (t.kind == Term::DiffT && species[t.j].comp == c)) has_stiff = has_diff = true;
    }
    const bool has_ext = has_extended_terms(c);
    if (has_ext) diff_const = false;   // fluxes are evaluated point by point
    o << "template <> struct DcComp<" << c << "> {\n";
    o << "  static constexpr int NS = " << std::max(ns, 1) << ";\n  static constexpr int NS_REAL = " << ns << ";\n";
    o << "  static constexpr bool HAS_MASS = " << has_mass << ", HAS_STIFF = " << has_stiff
      << ", HAS_DIFF = " << has_diff << ", DIFF_CONST = " << diff_const << ", HAS_EXT = " << has_ext << ";\n";
    // pattern mask (which species pairs exist in the sparsity pattern)
    {
      int n = std::max(ns, 1);
      std::vector<unsigned long long> words((n * n + 63) / 64, 0ull);
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j)
          if (std::find(pairs.begin(), pairs.end(), std::make_pair(g0 + i, g0 + j)) != pairs.end())
            words[(i * n + j) >> 6] |= 1ull << ((i * n + j) & 63);
      o << "  __host__ __device__ static constexpr bool pair(int i, int j) {\n    const unsigned long long m[" << words.size() << "] = {";
      for (size_t w = 0; w < words.size(); ++w) o << words[w] << "ull" << (w + 1 < words.size() ? "," : "");
      o << "};\n    return (m[(i * NS + j) >> 6] >> ((i * NS + j) & 63)) & 1ull;\n  }\n";
    }
    // diffusion mask (which D_ij are written in the ini): lets the kernels skip the zero couplings
    {
      int n = std::max(ns, 1);
      std::vector<unsigned long long> words((n * n + 63) / 64, 0ull);
      for (auto& t : terms)
        if ((t.kind == Term::Diff || t.kind == Term::DiffT) && species[t.i].comp == c && species[t.j].comp == c) {
          int i = species[t.i].local, j = species[t.j].local;
          words[(i * n + j) >> 6] |= 1ull << ((i * n + j) & 63);
        }
      o << "  __host__ __device__ static constexpr bool dpair(int i, int j) {\n    const unsigned long long m[" << words.size() << "] = {";
      for (size_t w = 0; w < words.size(); ++w) o << words[w] << "ull" << (w + 1 < words.size() ? "," : "");
      o << "};\n    return (m[(i * NS + j) >> 6] >> ((i * NS + j) & 63)) & 1ull;\n  }\n";
    }
    // ---- scalar part of the residual at a point
    o << "  // sc[i] = wA*(-R_i) + wM*(u_i*storage_i)   (local_operator.hh:479-480)\n";
    o << "  __device__ __forceinline__ static void scalar(const DcCtx& c, const double* u, const double (*g)[DC_DIM], double wM, double wA, double* sc) {\n";
    const bool poly = cfg.sub("model.assembly.b200").get("polynomial_terms", true);
    {
      std::vector<WeightedEntry> entries;
      for (int i = 0; i < ns; ++i) {
        const Term* r = find(Term::Reaction, g0 + i);
        const Term* s = find(Term::Storage, g0 + i);
        WeightedEntry e;
        e.lhs = "sc[" + std::to_string(i) + "]";
        if (r) e.a = make_node(Op::Neg, {r->ast});
        if (s) e.m = make_node(Op::Mul, {make_var(species[g0 + i].name), s->ast});
        entries.push_back(e);
      }
      if (!(poly && emit_polynomial_entries(o, entries, sym)))
        for (int i = 0; i < ns; ++i) {
          const Term* r = find(Term::Reaction, g0 + i);
          const Term* s = find(Term::Storage, g0 + i);
          o << "    sc[" << i << "] = ";
          if (!r && !s) o << "0.0";
          if (r) o << "wA * (-(" << code(*r) << "))";
          if (s) o << (r ? " + " : "") << "wM * (u[" << i << "] * (" << code(*s) << "))";
          o << ";\n";
        }
    }
    o << "    (void)c; (void)u; (void)g; (void)wM; (void)wA; (void)sc;\n  }\n";
    // ---- diffusive flux
    o << "  // fl[i][k] = -wA * sum_j D_ij * grad(u_j)[k]   (local_operator.hh:481-483)\n";
    o << "  __device__ __forceinline__ static void flux(const DcCtx& c, const double* u, const double (*g)[DC_DIM], double wA, double (*fl)[DC_DIM]) {\n";
    for (int i = 0; i < ns; ++i) {
      o << "    {";
      for (int k = 0; k < dim; ++k) o << " fl[" << i << "][" << k << "] = 0.0;";
      o << "\n";
      for (auto& t : terms)
        if (t.kind == Term::Diff && t.i == g0 + i && species[t.j].comp == c) {
          int j = species[t.j].local;
          o << "      { const double D = wA * (" << code(t) << ");";
          for (int k = 0; k < dim; ++k) o << " fl[" << i << "][" << k << "] -= D * g[" << j << "][" << k << "];";
          o << " }\n";
        } else if (t.kind == Term::DiffT && t.i == g0 + i && species[t.j].comp == c) {
          // tensor diffusion: out[r] = sum_c D[r][c] in[c]
          o << "      fl[" << i << "][" << t.k / 3 << "] -= wA * (" << code(t) << ") * g[" << species[t.j].local << "]["
            << t.k % 3 << "];\n";
        } else if (t.kind == Term::Vel && t.i == g0 + i) {
          // advection: flux += velocity * u_i  (local_operator.hh:481)
          o << "      fl[" << i << "][" << t.j << "] += wA * (" << code(t) << ") * u[" << i << "];\n";
        }
      o << "    }\n";
    }
    o << "    (void)c; (void)u; (void)g; (void)wA; (void)fl;\n  }\n";
    // ---- Jacobian coefficients
    o << "  // jm[i][j]: coefficient of phi_a*phi_b = wA*(-dR_i/du_j) + wM*(stg_i*delta_ij + dstg_i/du_j*u_i)\n"
         "  //   (local_operator.hh:605-641)\n";
    o << "  __device__ __forceinline__ static void jac_mass(const DcCtx& c, const double* u, const double (*g)[DC_DIM], double wM, double wA, double (*jm)[NS]) {\n";
    o << "    for (int i = 0; i < NS; ++i) for (int j = 0; j < NS; ++j) jm[i][j] = 0.0;\n";
    {
      std::map<std::pair<int, int>, WeightedEntry> by_entry;
      auto sum = [](const NodeP& acc, const NodeP& x) { return acc ? make_node(Op::Add, {acc, x}) : x; };
      for (auto& t : terms) {
        if (species[t.i].comp != c) continue;
        int i = species[t.i].local;
        if (t.kind == Term::ReactionJac && species[t.j].comp == c) {
          auto& e = by_entry[{i, species[t.j].local}];
          e.a = sum(e.a, make_node(Op::Neg, {t.ast}));
        }
        if (t.kind == Term::Storage) {
          auto& e = by_entry[{i, i}];
          e.m = sum(e.m, t.ast);
        }
        if (t.kind == Term::StorageJac && species[t.j].comp == c) {
          auto& e = by_entry[{i, species[t.j].local}];
          e.m = sum(e.m, make_node(Op::Mul, {t.ast, make_var(species[t.i].name)}));
        }
      }
      std::vector<WeightedEntry> entries;
      for (auto& kv : by_entry) {
        kv.second.lhs = "jm[" + std::to_string(kv.first.first) + "][" + std::to_string(kv.first.second) + "]";
        entries.push_back(kv.second);
      }
      if (!(poly && emit_polynomial_entries(o, entries, sym)))
        for (auto& t : terms) {
          if (species[t.i].comp != c) continue;
          int i = species[t.i].local;
          if (t.kind == Term::ReactionJac && species[t.j].comp == c)
            o << "    jm[" << i << "][" << species[t.j].local << "] += wA * (-(" << code(t) << "));\n";
          if (t.kind == Term::Storage)
            o << "    jm[" << i << "][" << i << "] += wM * (" << code(t) << ");\n";
          if (t.kind == Term::StorageJac && species[t.j].comp == c)
            o << "    jm[" << i << "][" << species[t.j].local << "] += wM * ((" << code(t) << ") * u[" << i << "]);\n";
        }
    }
    o << "    (void)c; (void)u; (void)g; (void)wM; (void)wA;\n  }\n";
    o << "  // jd[i][j]: coefficient of grad(phi_a).grad(phi_b) = wA*D_ij   (local_operator.hh:674-685)\n";
    o << "  __device__ __forceinline__ static void jac_diff(const DcCtx& c, const double* u, const double (*g)[DC_DIM], double wA, double (*jd)[NS]) {\n";
    o << "    for (int i = 0; i < NS; ++i) for (int j = 0; j < NS; ++j) jd[i][j] = 0.0;\n";
    for (auto& t : terms)
      if (t.kind == Term::Diff && species[t.i].comp == c && species[t.j].comp == c)
        o << "    jd[" << species[t.i].local << "][" << species[t.j].local << "] += wA * (" << code(t) << ");\n";
    o << "    (void)c; (void)u; (void)g; (void)wA;\n  }\n";
    // ---- general (extended) Jacobian coefficients, consumed by dc_ext_jacobian only
    o << "  // DT[i][j][r][c]: (DT grad(phi_a)) . grad(phi_b), scalar coefficients on the diagonal (local_operator.hh:674-685)\n";
    o << "  __device__ __noinline__ static void jac_diff_t(const DcCtx& c, const double* u, const double (*g)[DC_DIM], double wA, double (*DT)[NS][DC_DIM][DC_DIM]) {\n";
    o << "    for (int i = 0; i < NS; ++i) for (int j = 0; j < NS; ++j) for (int r = 0; r < DC_DIM; ++r) for (int s = 0; s < DC_DIM; ++s) DT[i][j][r][s] = 0.0;\n";
    if (has_ext)
      for (auto& t : terms) {
        if (species[t.i].comp != c || (t.kind != Term::Diff && t.kind != Term::DiffT) || species[t.j].comp != c) continue;
        int i = species[t.i].local;
        if (t.kind == Term::Diff) {
          o << "    { const double D = wA * (" << code(t) << ");";
          for (int k = 0; k < dim; ++k) o << " DT[" << i << "][" << species[t.j].local << "][" << k << "][" << k << "] += D;";
          o << " }\n";
        } else if (t.kind == Term::DiffT) {
          o << "    DT[" << i << "][" << species[t.j].local << "][" << t.k / 3 << "][" << t.k % 3 << "] += wA * (" << code(t) << ");\n";
        }
      }
    o << "    (void)c; (void)u; (void)g; (void)wA;\n  }\n";
    o << "  // W[i][k][r]: coefficient of phi_a * d(phi_b)/dx_r -- advection (local_operator.hh:643-671) and\n"
         "  //   dD_ij/du_k grad(u_k) (:688-700), with the index roles exactly as written there\n";
    o << "  __device__ __noinline__ static void jac_ext(const DcCtx& c, const double* u, const double (*g)[DC_DIM], double wA, double (*W)[NS][DC_DIM]) {\n";
    o << "    for (int i = 0; i < NS; ++i) for (int k = 0; k < NS; ++k) for (int r = 0; r < DC_DIM; ++r) W[i][k][r] = 0.0;\n";
    if (has_ext)
      for (auto& t : terms) {
        if (species[t.i].comp != c) continue;
        int i = species[t.i].local;
        if (t.kind == Term::Vel) {
          o << "    W[" << i << "][" << i << "][" << t.j << "] -= wA * (" << code(t) << ");\n";
        } else if (t.kind == Term::VelJac && species[t.j].comp == c) {
          o << "    W[" << i << "][" << species[t.j].local << "][" << t.k << "] -= wA * (" << code(t) << ") * u[" << i << "];\n";
        } else if (t.kind == Term::DiffJac && species[t.k].comp == c) {
          int k = species[t.k].local;
          o << "    { const double dD = wA * (" << code(t) << ");";
          for (int r = 0; r < dim; ++r) o << " W[" << i << "][" << k << "][" << r << "] += dD * g[" << k << "][" << r << "];";
          o << " }\n";
        } else if (t.kind == Term::DiffTJac && species[t.k / 9].comp == c) {
          int k = species[t.k / 9].local, r = (t.k % 9) / 3, cc = t.k % 3;
          o << "    W[" << i << "][" << k << "][" << r << "] += wA * (" << code(t) << ") * g[" << k << "][" << cc << "];\n";
        }
      }
    o << "    (void)c; (void)u; (void)g; (void)wA;\n  }\n";
    o << "};\n";
  }
  // ---- outflow (skeleton / boundary) per directional compartment pair
  auto opairs = outflow_pairs();
  o << "#define DC_NOUTFLOW " << opairs.size() << "\n";
  for (size_t p = 0; p < opairs.size(); ++p) {
    int cs = opairs[p].first, ct = opairs[p].second;
    bool boundary = cs == ct;
    int nss = comp_nspec[cs], nst = boundary ? 0 : comp_nspec[ct];
    SymbolMap sym{*this, cs, boundary ? -1 : ct, true};
    o << "template <> struct DcOutflow<" << p << "> {\n";
    o << "  static constexpr int CS = " << cs << ", CT = " << ct << ", NSS = " << nss << ", NST = " << std::max(nst, 1)
      << ", NST_REAL = " << nst << ";\n  static constexpr bool BOUNDARY = " << boundary << ";\n";
    std::vector<std::vector<int>> ps(nss, std::vector<int>(nss, 0)), pt(nss, std::vector<int>(std::max(nst, 1), 0));
    std::ostringstream fl, jc;
    for (int i = 0; i < nss; ++i) fl << "    T[" << i << "] = 0.0;\n";
    for (auto& t : terms) {
      if (species[t.i].comp != cs || t.j != ct) continue;
      int i = species[t.i].local;
      if (t.kind == Term::Outflow) fl << "    T[" << i << "] = " << to_cuda(t.ast, sym) << ";\n";
      if (t.kind == Term::OutflowJac) {
        const auto& sk = species[t.k];
        if (sk.comp == cs) {
          ps[i][sk.local] = 1;
          jc << "    js[" << i << "][" << sk.local << "] += " << to_cuda(t.ast, sym) << ";\n";
        } else if (!boundary && sk.comp == ct) {
          pt[i][sk.local] = 1;
          jc << "    jt[" << i << "][" << sk.local << "] += " << to_cuda(t.ast, sym) << ";\n";
        }
      }
    }
    auto mask = [&](const char* name, const std::vector<std::vector<int>>& m, int cols) {
      std::vector<unsigned long long> words((m.size() * cols + 63) / 64 + 1, 0ull);
      for (size_t i = 0; i < m.size(); ++i)
        for (int j = 0; j < cols; ++j)
          if (m[i][j]) words[(i * cols + j) >> 6] |= 1ull << ((i * cols + j) & 63);
      o << "  __host__ __device__ static constexpr bool " << name << "(int i, int j) {\n    const unsigned long long m[" << words.size() << "] = {";
      for (size_t w = 0; w < words.size(); ++w) o << words[w] << "ull" << (w + 1 < words.size() ? "," : "");
      o << "};\n    return (m[(i * " << cols << " + j) >> 6] >> ((i * " << cols << " + j) & 63)) & 1ull;\n  }\n";
    };
    mask("pair_s", ps, nss);
    mask("pair_t", pt, std::max(nst, 1));
    o << "  // T[i]: outflow of own species i through the facet (local_operator.hh:920-927)\n";
    o << "  __device__ __forceinline__ static void flux(const DcCtx& c, const double* us, const double (*gs)[DC_DIM], const double* ut, const double (*gt)[DC_DIM], double* T) {\n"
      << fl.str() << "    (void)c; (void)us; (void)gs; (void)ut; (void)gt;\n  }\n";
    o << "  // js/jt: dT_i/du_k for k on the own / other side (local_operator.hh:1127-1142)\n";
    o << "  __device__ __forceinline__ static void jacobian(const DcCtx& c, const double* us, const double (*gs)[DC_DIM], const double* ut, const double (*gt)[DC_DIM], double (*js)[NSS], double (*jt)[NST]) {\n"
      << "    for (int i = 0; i < NSS; ++i) { for (int j = 0; j < NSS; ++j) js[i][j] = 0.0; for (int j = 0; j < NST; ++j) jt[i][j] = 0.0; }\n"
      << jc.str() << "    (void)c; (void)us; (void)gs; (void)ut; (void)gt;\n  }\n";
    o << "};\n";
  }
  return o.str();
}

}  // namespace dcb
