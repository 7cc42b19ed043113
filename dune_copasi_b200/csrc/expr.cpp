#include "expr.hpp"

#include <mutex>
#include <set>
#include <sstream>

#include <algorithm>

#include <cfloat>
#include <cmath>
#include <cstring>

namespace dcb {

namespace {

struct Tok {
  enum Kind { Num, Id, Op, End } kind;
  std::string text;
  double num = 0;
};

std::vector<Tok> lex(const std::string& s) {
  std::vector<Tok> out;
  size_t i = 0, n = s.size();
  while (i < n) {
    char c = s[i];
    if (isspace((unsigned char)c)) { ++i; continue; }
    if (isdigit((unsigned char)c) || (c == '.' && i + 1 < n && isdigit((unsigned char)s[i + 1]))) {
      size_t j = i;
      while (j < n && isdigit((unsigned char)s[j])) ++j;
      if (j < n && s[j] == '.') { ++j; while (j < n && isdigit((unsigned char)s[j])) ++j; }
      if (j < n && (s[j] == 'e' || s[j] == 'E')) {
        size_t k = j + 1;
        if (k < n && (s[k] == '+' || s[k] == '-')) ++k;
        if (k < n && isdigit((unsigned char)s[k])) {
          while (k < n && isdigit((unsigned char)s[k])) ++k;
          j = k;
        }
      }
      Tok t{Tok::Num, s.substr(i, j - i)};
      t.num = std::strtod(t.text.c_str(), nullptr);
      out.push_back(t);
      i = j;
      continue;
    }
    if (isalpha((unsigned char)c) || c == '_') {
      size_t j = i;
      while (j < n && (isalnum((unsigned char)s[j]) || s[j] == '_')) ++j;
      out.push_back({Tok::Id, s.substr(i, j - i)});
      i = j;
      continue;
    }
    static const char* two[] = {"**", "<=", ">=", "==", "!=", "&&", "||"};
    bool hit = false;
    for (auto t : two)
      if (i + 1 < n && s[i] == t[0] && s[i + 1] == t[1]) {
        out.push_back({Tok::Op, t});
        i += 2;
        hit = true;
        break;
      }
    if (hit) continue;
    if (strchr("+-*/^%<>()?:,&|!", c)) {
      out.push_back({Tok::Op, std::string(1, c)});
      ++i;
      continue;
    }
    fail("expression '", s, "': unexpected character '", c, "' at ", i);
  }
  out.push_back({Tok::End, ""});
  return out;
}

NodeP mk(Op op, std::vector<NodeP> kids) {
  auto n = std::make_shared<Node>();
  n->op = op;
  n->kids = std::move(kids);
  return n;
}
NodeP mk_num(double v) {
  auto n = std::make_shared<Node>();
  n->op = Op::Num;
  n->num = v;
  return n;
}

// recursive descent; precedence low->high: ?: | or | and | == != | < > <= >= | + - | * / % | unary | ^
struct P {
  std::vector<Tok> t;
  size_t i = 0;
  const std::string& src;
  explicit P(const std::string& s) : t(lex(s)), src(s) {}
  bool is(const char* v) const { return (t[i].kind == Tok::Op || t[i].kind == Tok::Id) && t[i].text == v; }
  bool eat(const char* v) { if (is(v)) { ++i; return true; } return false; }
  void need(const char* v) { if (!eat(v)) fail("expression '", src, "': expected '", v, "' near token ", i); }

  NodeP ternary() {
    NodeP c = lor();
    if (eat("?")) {
      NodeP a = ternary();
      need(":");
      NodeP b = ternary();
      return mk(Op::Sel, {c, a, b});
    }
    return c;
  }
  NodeP lor() {
    NodeP a = land();
    while (is("or") || is("||") || is("|")) { ++i; a = mk(Op::Or, {a, land()}); }
    return a;
  }
  NodeP land() {
    NodeP a = equality();
    while (is("and") || is("&&") || is("&")) { ++i; a = mk(Op::And, {a, equality()}); }
    return a;
  }
  NodeP equality() {
    NodeP a = relational();
    for (;;) {
      if (eat("==")) a = mk(Op::Eq, {a, relational()});
      else if (eat("!=")) a = mk(Op::Ne, {a, relational()});
      else return a;
    }
  }
  NodeP relational() {
    NodeP a = additive();
    for (;;) {
      if (eat("<=")) a = mk(Op::Le, {a, additive()});
      else if (eat(">=")) a = mk(Op::Ge, {a, additive()});
      else if (eat("<")) a = mk(Op::Lt, {a, additive()});
      else if (eat(">")) a = mk(Op::Gt, {a, additive()});
      else return a;
    }
  }
  NodeP additive() {
    NodeP a = term();
    for (;;) {
      if (eat("+")) a = mk(Op::Add, {a, term()});
      else if (eat("-")) a = mk(Op::Sub, {a, term()});
      else return a;
    }
  }
  NodeP term() {
    NodeP a = unary();
    for (;;) {
      if (eat("*")) a = mk(Op::Mul, {a, unary()});
      else if (eat("/")) a = mk(Op::Div, {a, unary()});
      else if (eat("%")) a = mk(Op::Mod, {a, unary()});
      else return a;
    }
  }
  NodeP unary() {
    if (eat("-")) return mk(Op::Neg, {unary()});
    if (eat("+")) return unary();
    if (eat("!") || eat("not")) return mk(Op::Not, {unary()});
    return power();
  }
  NodeP power() {
    NodeP b = atom();
    if (eat("^") || eat("**")) return mk(Op::Pow, {b, unary()});  // right assoc, signed exponent
    return b;
  }
  NodeP atom() {
    Tok k = t[i++];
    if (k.kind == Tok::Num) return mk_num(k.num);
    if (k.kind == Tok::Id) {
      auto n = std::make_shared<Node>();
      n->name = k.text;
      if (eat("(")) {
        n->op = Op::Call;
        if (!eat(")")) {
          for (;;) {
            n->kids.push_back(ternary());
            if (eat(")")) break;
            need(",");
          }
        }
      } else
        n->op = Op::Var;
      return n;
    }
    if (k.kind == Tok::Op && k.text == "(") {
      NodeP e = ternary();
      need(")");
      return e;
    }
    fail("expression '", src, "': unexpected token '", k.text, "'");
  }
};

NodeP subst(const NodeP& a, const std::map<std::string, NodeP>& env) {
  if (a->op == Op::Num) return a;
  if (a->op == Op::Var) {
    auto it = env.find(a->name);
    return it == env.end() ? a : it->second;
  }
  auto n = std::make_shared<Node>(*a);
  for (auto& k : n->kids) k = subst(k, env);
  return n;
}

// tables by their unique key: trees carry only the key (Op::Call), evaluation and lowering look the data up here
std::map<std::string, std::shared_ptr<const Table>>& table_registry() {
  static std::map<std::string, std::shared_ptr<const Table>> r;
  return r;
}
std::mutex& table_mutex() { static std::mutex m; return m; }
std::shared_ptr<const Table> find_table(const std::string& key) {
  if (key.rfind("dc_tab_", 0) != 0) return nullptr;
  std::lock_guard<std::mutex> lock(table_mutex());
  auto it = table_registry().find(key);
  return it == table_registry().end() ? nullptr : it->second;
}

double apply1(const std::string& f, double a, bool* ok) {
  *ok = true;
  if (auto t = find_table(f)) {
    if (t->kind == 2) fail("function '", t->name, "' expects 2 arguments");
    bool oob = false;
    double v = t->eval(a, &oob);
    if (oob) fail("interpolation of function '", t->name, "' is out of bounds: ", a, " not in [", t->domain[0], ", ", t->domain[1], "]");
    return v;
  }
  if (f == "sqrt") return std::sqrt(a);
  if (f == "exp") return std::exp(a);
  if (f == "log" || f == "ln") return std::log(a);
  if (f == "sin") return std::sin(a);
  if (f == "cos") return std::cos(a);
  if (f == "tan") return std::tan(a);
  if (f == "abs") return std::fabs(a);
  if (f == "floor") return std::floor(a);
  if (f == "ceil") return std::ceil(a);
  if (f == "tanh") return std::tanh(a);
  if (f == "sinh") return std::sinh(a);
  if (f == "cosh") return std::cosh(a);
  if (f == "asin") return std::asin(a);
  if (f == "acos") return std::acos(a);
  if (f == "atan") return std::atan(a);
  if (f == "log10") return std::log10(a);
  if (f == "log2") return std::log2(a);
  if (f == "exp2") return std::exp2(a);
  if (f == "round") return std::round(a);
  if (f == "sgn" || f == "sign") return (a > 0) - (a < 0);
  *ok = false;
  return 0;
}

double apply2(const std::string& f, double a, double b, bool* ok) {
  *ok = true;
  if (auto t = find_table(f)) {
    if (t->kind == 2) return t->img(a, b);
  }
  if (f == "min") return a < b ? a : b;
  if (f == "max") return a > b ? a : b;
  if (f == "atan2") return std::atan2(a, b);
  if (f == "pow") return std::pow(a, b);
  *ok = false;
  return 0;
}

double binop(Op op, double a, double b) {
  switch (op) {
    case Op::Add: return a + b;
    case Op::Sub: return a - b;
    case Op::Mul: return a * b;
    case Op::Div: return a / b;
    case Op::Pow: return std::pow(a, b);
    case Op::Mod: return std::fmod(a, b);
    case Op::Lt: return a < b;
    case Op::Gt: return a > b;
    case Op::Le: return a <= b;
    case Op::Ge: return a >= b;
    case Op::Eq: return a == b;
    case Op::Ne: return a != b;
    case Op::And: return (a != 0.0) && (b != 0.0);
    case Op::Or: return (a != 0.0) || (b != 0.0);
    default: fail("internal: not a binary op");
  }
}

NodeP fold(const NodeP& a) {
  // children are already folded
  bool all = !a->kids.empty();
  for (auto& k : a->kids) all = all && k->op == Op::Num;
  if (!all) return a;
  switch (a->op) {
    case Op::Neg: return mk_num(-a->kids[0]->num);
    case Op::Not: return mk_num(a->kids[0]->num == 0.0);
    case Op::Sel: return a->kids[0]->num != 0.0 ? a->kids[1] : a->kids[2];
    case Op::Call: {
      bool ok = false;
      double v = 0;
      if (a->name.rfind("dc_tab_", 0) == 0) return a;   // tabulated functions are folded in resolve_rec
      if (a->kids.size() == 1) v = apply1(a->name, a->kids[0]->num, &ok);
      else if (a->kids.size() >= 2) {
        v = a->kids[0]->num;
        ok = true;
        for (size_t i = 1; i < a->kids.size() && ok; ++i) v = apply2(a->name, v, a->kids[i]->num, &ok);
      }
      return ok ? mk_num(v) : a;
    }
    case Op::Num: case Op::Var: return a;
    default: return mk_num(binop(a->op, a->kids[0]->num, a->kids[1]->num));
  }
}

NodeP resolve_rec(const NodeP& a, const ParserContext& ctx, int depth) {
  if (depth > 16) fail("parser_context functions nested too deep (recursion?)");
  if (a->op == Op::Num) return a;
  if (a->op == Op::Var) {
    auto it = ctx.constants.find(a->name);
    if (it != ctx.constants.end()) return mk_num(it->second);
    if (a->name == "no_value") return mk_num(DBL_MAX);
    if (a->name == "pi") return mk_num(3.14159265358979323846);
    return a;
  }
  auto n = std::make_shared<Node>(*a);
  for (auto& k : n->kids) k = resolve_rec(k, ctx, depth);
  if (n->op == Op::Call) {
    auto it = ctx.functions.find(n->name);
    if (it != ctx.functions.end() && !ctx.tables.count(n->name)) {
      const auto& fn = it->second;
      if (fn.args.size() != n->kids.size())
        fail("function '", n->name, "' expects ", fn.args.size(), " arguments, got ", n->kids.size());
      std::map<std::string, NodeP> env;
      for (size_t i = 0; i < fn.args.size(); ++i) env[fn.args[i]] = n->kids[i];
      return resolve_rec(subst(parse_expr(fn.body), env), ctx, depth + 1);
    }
    auto tb = ctx.tables.find(n->name);
    if (tb != ctx.tables.end() && tb->second->kind == 2) {
      if (n->kids.size() != 2) fail("function '", n->name, "' expects 2 arguments, got ", n->kids.size());
      n->name = tb->second->key;
      if (n->kids[0]->op == Op::Num && n->kids[1]->op == Op::Num) return mk_num(tb->second->img(n->kids[0]->num, n->kids[1]->num));
      return n;
    }
    if (tb != ctx.tables.end()) {
      if (n->kids.size() != 1) fail("function '", n->name, "' expects 1 argument, got ", n->kids.size());
      n->name = tb->second->key;
      if (n->kids[0]->op == Op::Num) {   // constant argument: fold unless it is out of bounds (left to run time)
        bool oob = false;
        double v = tb->second->eval(n->kids[0]->num, &oob);
        if (!oob) return mk_num(v);
      }
      return n;
    }
    if (n->name == "if" && n->kids.size() == 3) return fold(mk(Op::Sel, n->kids));
  }
  return fold(n);
}

std::string num_lit(double v) {
  if (std::isinf(v)) return v > 0 ? "(1.0/0.0)" : "(-1.0/0.0)";
  if (std::isnan(v)) return "(0.0/0.0)";
  char buf[64];
  snprintf(buf, sizeof buf, "%.17g", v);
  std::string s = buf;
  if (s.find_first_of(".eE") == std::string::npos) s += ".0";
  if (v < 0) s = "(" + s + ")";
  return s;
}

bool is_cmp(Op op) {
  return op == Op::Lt || op == Op::Gt || op == Op::Le || op == Op::Ge || op == Op::Eq ||
         op == Op::Ne || op == Op::And || op == Op::Or || op == Op::Not;
}

}  // namespace

ParserContext::Fn parse_function_expression(const std::string& e, const std::string& what) {
  auto colon = e.find(':');
  if (colon == std::string::npos) fail(what, ": function needs 'args: body'");
  ParserContext::Fn fn;
  std::string head = e.substr(0, colon);
  size_t p = 0;
  while (p <= head.size()) {
    size_t q = head.find(',', p);
    if (q == std::string::npos) q = head.size();
    std::string a = trim(head.substr(p, q - p));
    if (!a.empty()) fn.args.push_back(a);
    p = q + 1;
  }
  fn.body = trim(e.substr(colon + 1));
  return fn;
}

// std::lerp as libstdc++ implements it (the reference calls std::lerp): exact at the ends, monotonic
static double lerp_std(double a, double b, double t) {
  if ((a <= 0 && b >= 0) || (a >= 0 && b <= 0)) return t * b + (1 - t) * a;
  if (t == 1) return b;
  const double x = a + t * (b - a);
  return (t > 1) == (b > a) ? (b < x ? x : b) : (x < b ? x : b);
}

double Table::eval(double x, bool* out_of_bounds) const {
  if (kind == 0) {   // context.cc:83-95
    const size_t d = std::lower_bound(domain.begin(), domain.end(), x) - domain.begin();
    if (d == 0) return range.front();
    if (d == domain.size()) return range.back();
    return lerp_std(range[d - 1], range[d], (x - domain[d - 1]) / (domain[d] - domain[d - 1]));
  }
  // context.cc:256-277
  const double d0 = domain[0], d1 = domain[1];
  const size_t n = range.size() - 1;   // intervals
  if (clamp) x = d0 < x ? x : d0;
  else if (x < d0 || x > d1 || !(x == x)) {
    if (out_of_bounds) *out_of_bounds = true;
    return std::nan("");
  }
  double interval = (x - d0) * ((double)n / (d1 - d0));
  double whole;
  std::modf(interval, &whole);
  size_t k = whole <= 0.0 ? 0 : (whole >= (double)n ? n : (size_t)whole);
  return range[k];
}

ParserContext ParserContext::from_config(const PTree& pc) {
  ParserContext ctx;
  std::vector<std::string> sampled;   // functions with interpolate = true: tabulated once the context is complete
  for (auto& name : pc.sub_keys()) {
    const PTree& s = pc.sub(name);
    std::string type = s.get("type", std::string());
    if (type == "constant") {
      ctx.constants[name] = s.get("value", 0.0);
    } else if (type == "function") {
      ctx.functions[name] = parse_function_expression(s.get("expression", std::string()), "parser_context." + name);
      if (s.get("interpolate", false)) sampled.push_back(name);
    } else if (type == "interpolation") {
      auto t = std::make_shared<Table>();
      t->kind = 0;
      t->name = name;
      t->domain = s.get_vec("domain", {});
      t->range = s.get_vec("range", {});
      if (!std::is_sorted(t->domain.begin(), t->domain.end())) fail("parser_context.", name, ": the interpolation domain must be sorted");
      if (t->domain.size() < 2 || t->domain.size() != t->range.size())
        fail("parser_context.", name, ": interpolation range and domain must have at least two points and be the same size");
      ctx.tables[name] = t;
    } else if (type == "tiff") {
      auto t = std::make_shared<Table>();
      t->kind = 2;
      t->name = name;
      t->img = read_tiff(s.get("path", std::string()));
      ctx.tables[name] = t;
    } else if (!type.empty()) {
      // random_field needs parafields (absent from the image); it only feeds initial conditions
      fail("parser_context.", name, ": type '", type, "' is not supported by this build");
    }
  }
  for (auto& name : sampled) {
    const PTree& s = pc.sub(name);
    const Fn fn = ctx.functions.at(name);
    if (fn.args.size() != 1) fail("parser_context.", name, ": cannot interpolate a function with ", fn.args.size(), " arguments");
    const long long intervals = s.get("interpolation.intervals", 1000);
    if (intervals > 100000) fail("parser_context.", name, ": number of interpolation intervals is too big");
    if (intervals < 1) fail("parser_context.", name, ": at least one interval is required");
    auto t = std::make_shared<Table>();
    t->kind = 1;
    t->name = name;
    t->domain = s.get_vec("interpolation.domain." + fn.args[0], {0.0, 1.0});
    if (t->domain.size() != 2 || t->domain[0] >= t->domain[1]) fail("parser_context.", name, ": domain arguments are not ordered");
    const std::string ooo = s.get("interpolation.out_of_bounds", std::string("error"));
    if (ooo != "clamp" && ooo != "error") fail("parser_context.", name, ": not known interpolation.out_of_bounds = ", ooo);
    t->clamp = ooo == "clamp";
    // sample the function itself (it may use the context's constants and its other, untabulated functions)
    ParserContext plain = ctx;
    plain.tables.clear();
    NodeP body = resolve_expr(parse_expr(fn.body), plain);
    const double width = (t->domain[1] - t->domain[0]) / (double)intervals;
    t->range.resize((size_t)intervals + 1);
    for (size_t i = 0; i < t->range.size(); ++i) {
      const double x = t->domain[0] + (double)i * width;
      t->range[i] = eval_expr(body, [&](const std::string& v) -> double {
        if (v != fn.args[0]) fail("parser_context.", name, ": unknown symbol '", v, "' while sampling the function");
        return x;
      });
    }
    ctx.tables[name] = t;
  }
  // unique keys: the data decides, so equal tables of different models share a symbol and different ones never clash
  for (auto& kv : ctx.tables) {
    auto t = std::const_pointer_cast<Table>(kv.second);
    unsigned long long h = 1469598103934665603ull;
    auto mix = [&](const void* p, size_t n) {
      for (size_t i = 0; i < n; ++i) { h ^= ((const unsigned char*)p)[i]; h *= 1099511628211ull; }
    };
    mix(&t->kind, sizeof t->kind); mix(&t->clamp, sizeof t->clamp);
    mix(t->domain.data(), t->domain.size() * sizeof(double));
    mix(t->range.data(), t->range.size() * sizeof(double));
    if (t->kind == 2) {
      const float geo[4] = {t->img.x_res, t->img.y_res, t->img.x_off, t->img.y_off};
      mix(&t->img.rows, sizeof t->img.rows); mix(&t->img.cols, sizeof t->img.cols); mix(geo, sizeof geo);
      mix(t->img.values.data(), t->img.values.size() * sizeof(double));
    }
    char buf[32];
    snprintf(buf, sizeof buf, "%016llx", h);
    std::string id;
    for (char ch : kv.first) id += (isalnum((unsigned char)ch) ? ch : '_');
    t->key = "dc_tab_" + id + "_" + buf;
    std::lock_guard<std::mutex> lock(table_mutex());
    table_registry()[t->key] = t;
  }
  return ctx;
}

std::string ParserContext::cuda_tables() const {
  std::ostringstream o;
  if (!tables.empty())
    o << "__device__ __forceinline__ double dc_lerp(double a, double b, double t) {\n"
         "  if ((a <= 0 && b >= 0) || (a >= 0 && b <= 0)) return t * b + (1 - t) * a;\n"
         "  if (t == 1) return b;\n"
         "  const double x = a + t * (b - a);\n"
         "  return (t > 1) == (b > a) ? (b < x ? x : b) : (x < b ? x : b);\n}\n";
  std::set<std::string> done;
  for (auto& kv : tables) {
    const Table& t = *kv.second;
    if (!done.insert(t.key).second) continue;
    auto array = [&](const char* what, const std::vector<double>& v) {
      o << "__device__ const double " << t.key << "_" << what << "[" << v.size() << "] = {";
      for (size_t i = 0; i < v.size(); ++i) o << (i ? ", " : "") << num_lit(v[i]);
      o << "};\n";
    };
    if (t.kind == 2) {
      if (t.img.values.size() > Table::kMaxDevicePixels) continue;   // host-only (to_cuda refuses to reference it)
      array("r", t.img.values);
      auto flit = [](float v) { char b[48]; snprintf(b, sizeof b, "%.9gf", (double)v); std::string s = b; if (s.find_first_of(".e") == std::string::npos) s.insert(s.size() - 1, ".0"); return s; };
      o << "__device__ __noinline__ double " << t.key << "(double x, double y) {\n"
        << "  const float fx = " << flit(t.img.x_res) << " * ((float)x - " << flit(t.img.x_off) << "), fy = " << flit(t.img.y_res)
        << " * ((float)y - " << flit(t.img.y_off) << ");\n"
        << "  unsigned px = fx <= 0.0f ? 0u : (fx >= 4294967040.0f ? 4294967295u : (unsigned)fx);\n"
        << "  const unsigned py = fy <= 0.0f ? 0u : (fy >= 4294967040.0f ? 4294967295u : (unsigned)fy);\n"
        << "  unsigned line = " << t.img.rows << "u - py - 1u;\n"
        << "  if (px > " << t.img.cols - 1 << "u) px = " << t.img.cols - 1 << "u;\n"
        << "  if (line > " << t.img.rows - 1 << "u) line = " << t.img.rows - 1 << "u;\n"
        << "  return " << t.key << "_r[(size_t)line * " << t.img.cols << " + px];\n}\n";
      continue;
    }
    array("r", t.range);
    if (t.kind == 0) {
      array("d", t.domain);
      const size_t n = t.domain.size();
      // std::lower_bound: first index with domain[index] >= x
      o << "__device__ __noinline__ double " << t.key << "(double x) {\n"
        << "  int lo = 0, hi = " << n << ";\n"
        << "  while (lo < hi) { const int mid = (lo + hi) >> 1; if (" << t.key << "_d[mid] < x) lo = mid + 1; else hi = mid; }\n"
        << "  if (lo == 0) return " << t.key << "_r[0];\n"
        << "  if (lo == " << n << ") return " << t.key << "_r[" << n - 1 << "];\n"
        << "  return dc_lerp(" << t.key << "_r[lo - 1], " << t.key << "_r[lo], (x - " << t.key << "_d[lo - 1]) / (" << t.key
        << "_d[lo] - " << t.key << "_d[lo - 1]));\n}\n";
    } else {
      const size_t n = t.range.size() - 1;
      o << "__device__ __noinline__ double " << t.key << "(double x) {\n"
        << "  const double d0 = " << num_lit(t.domain[0]) << ", d1 = " << num_lit(t.domain[1]) << ";\n";
      if (t.clamp) o << "  x = d0 < x ? x : d0;\n";
      else o << "  if (x < d0 || x > d1 || !(x == x)) return 0.0 / 0.0;\n";
      o << "  double whole;\n  modf((x - d0) * (" << num_lit((double)n) << " / (d1 - d0)), &whole);\n"
        << "  const int k = whole <= 0.0 ? 0 : (whole >= " << num_lit((double)n) << " ? " << n << " : (int)whole);\n"
        << "  return " << t.key << "_r[k];\n}\n";
    }
  }
  return o.str();
}

bool expr_is_absent(const std::string& text) {
  std::string s = trim(text);
  if (s.empty()) return true;
  // literal zero: optional sign, zeros with optional point and exponent
  size_t i = 0;
  if (s[i] == '+' || s[i] == '-') ++i;
  while (i < s.size() && isspace((unsigned char)s[i])) ++i;
  size_t digits = 0;
  bool dot = false;
  for (; i < s.size(); ++i) {
    if (s[i] == '0') ++digits;
    else if (s[i] == '.' && !dot) dot = true;
    else break;
  }
  if (digits == 0) return false;
  if (i < s.size() && (s[i] == 'e' || s[i] == 'E')) {
    ++i;
    if (i < s.size() && (s[i] == '+' || s[i] == '-')) ++i;
    size_t d = 0;
    while (i < s.size() && isdigit((unsigned char)s[i])) { ++i; ++d; }
    if (d == 0) return false;
  }
  return i == s.size();
}

NodeP parse_expr(const std::string& text) {
  P p(text);
  NodeP e = p.ternary();
  if (p.t[p.i].kind != Tok::End) fail("expression '", text, "': trailing input '", p.t[p.i].text, "'");
  return e;
}

NodeP resolve_expr(const NodeP& ast, const ParserContext& ctx) { return resolve_rec(ast, ctx, 0); }

NodeP make_node(Op op, std::vector<NodeP> kids) { return mk(op, std::move(kids)); }
NodeP make_var(const std::string& name) {
  auto n = std::make_shared<Node>();
  n->op = Op::Var;
  n->name = name;
  return n;
}

bool is_constant(const NodeP& ast, double* value) {
  if (ast->op != Op::Num) return false;
  if (value) *value = ast->num;
  return true;
}

void collect_vars(const NodeP& a, std::vector<std::string>& out) {
  if (a->op == Op::Var) out.push_back(a->name);
  for (auto& k : a->kids) collect_vars(k, out);
}

double eval_expr(const NodeP& a, const std::function<double(const std::string&)>& lookup) {
  switch (a->op) {
    case Op::Num: return a->num;
    case Op::Var: return lookup(a->name);
    case Op::Neg: return -eval_expr(a->kids[0], lookup);
    case Op::Not: return eval_expr(a->kids[0], lookup) == 0.0;
    case Op::Sel:
      return eval_expr(a->kids[0], lookup) != 0.0 ? eval_expr(a->kids[1], lookup)
                                                  : eval_expr(a->kids[2], lookup);
    case Op::Call: {
      bool ok = false;
      double v = 0;
      if (a->kids.size() == 1) v = apply1(a->name, eval_expr(a->kids[0], lookup), &ok);
      else if (a->kids.size() >= 2) {
        v = eval_expr(a->kids[0], lookup);
        ok = true;
        for (size_t i = 1; i < a->kids.size() && ok; ++i)
          v = apply2(a->name, v, eval_expr(a->kids[i], lookup), &ok);
      }
      if (!ok) fail("unknown function '", a->name, "' with ", a->kids.size(), " argument(s)");
      return v;
    }
    default: return binop(a->op, eval_expr(a->kids[0], lookup), eval_expr(a->kids[1], lookup));
  }
}

static std::string cond_text(const NodeP& a, const std::function<std::string(const std::string&)>& sym);

std::string to_cuda(const NodeP& a, const std::function<std::string(const std::string&)>& sym) {
  auto rec = [&](const NodeP& n) { return to_cuda(n, sym); };
  switch (a->op) {
    case Op::Num: return num_lit(a->num);
    case Op::Var: {
      std::string s = sym(a->name);
      if (s.empty()) fail("unknown symbol '", a->name, "' in expression");
      return s;
    }
    case Op::Neg: return "(-" + rec(a->kids[0]) + ")";
    case Op::Add: return "(" + rec(a->kids[0]) + " + " + rec(a->kids[1]) + ")";
    case Op::Sub: return "(" + rec(a->kids[0]) + " - " + rec(a->kids[1]) + ")";
    case Op::Mul: return "(" + rec(a->kids[0]) + " * " + rec(a->kids[1]) + ")";
    case Op::Div: return "(" + rec(a->kids[0]) + " / " + rec(a->kids[1]) + ")";
    case Op::Mod: return "fmod(" + rec(a->kids[0]) + ", " + rec(a->kids[1]) + ")";
    case Op::Pow: {
      double e;
      if (is_constant(a->kids[1], &e) && e == std::floor(e) && std::fabs(e) <= 8) {
        // small integer powers become multiplications (pow() costs hundreds of fp64 instructions)
        int n = (int)std::fabs(e);
        if (n == 0) return "1.0";
        std::string b = rec(a->kids[0]);
        std::string r = "dc_powi<" + std::to_string(n) + ">(" + b + ")";
        return e < 0 ? "(1.0 / " + r + ")" : r;
      }
      if (is_constant(a->kids[1], &e) && e == 0.5) return "sqrt(" + rec(a->kids[0]) + ")";
      return "pow(" + rec(a->kids[0]) + ", " + rec(a->kids[1]) + ")";
    }
    case Op::Sel: return "(" + cond_text(a->kids[0], sym) + " ? " + rec(a->kids[1]) + " : " + rec(a->kids[2]) + ")";
    case Op::Call: {
      const std::string& f = a->name;
      static const std::map<std::string, std::string> one = {
          {"sqrt", "sqrt"}, {"exp", "exp"}, {"log", "log"}, {"ln", "log"}, {"sin", "sin"}, {"cos", "cos"},
          {"tan", "tan"}, {"abs", "fabs"}, {"floor", "floor"}, {"ceil", "ceil"}, {"tanh", "tanh"},
          {"sinh", "sinh"}, {"cosh", "cosh"}, {"asin", "asin"}, {"acos", "acos"}, {"atan", "atan"},
          {"log10", "log10"}, {"log2", "log2"}, {"exp2", "exp2"}, {"round", "round"},
          {"sgn", "dc_sgn"}, {"sign", "dc_sgn"}};
      static const std::map<std::string, std::string> two = {
          {"min", "dc_min"}, {"max", "dc_max"}, {"atan2", "atan2"}, {"pow", "pow"}};
      if (a->kids.size() == 1 && one.count(f)) return one.at(f) + "(" + rec(a->kids[0]) + ")";
      if (a->kids.size() == 1 && f.rfind("dc_tab_", 0) == 0) return f + "(" + rec(a->kids[0]) + ")";
      if (a->kids.size() == 2 && f.rfind("dc_tab_", 0) == 0) {
        auto t = find_table(f);
        if (t && t->img.values.size() > Table::kMaxDevicePixels)
          fail("image function '", t->name, "' (", t->img.rows, " x ", t->img.cols, " pixels) is too large for device code: "
               "use it in initial / constrain / compartment expressions, or an image of at most ", Table::kMaxDevicePixels, " pixels");
        return f + "(" + rec(a->kids[0]) + ", " + rec(a->kids[1]) + ")";
      }
      if (a->kids.size() >= 2 && two.count(f)) {
        std::string r = rec(a->kids[0]);
        for (size_t i = 1; i < a->kids.size(); ++i) r = two.at(f) + "(" + r + ", " + rec(a->kids[i]) + ")";
        return r;
      }
      fail("unknown function '", f, "' with ", a->kids.size(), " argument(s)");
    }
    default:  // comparisons and logic produce 0/1 doubles (ExprTk semantics)
      return "(" + cond_text(a, sym) + " ? 1.0 : 0.0)";
  }
}

// boolean-valued C text of a node used as a condition
static std::string cond_text(const NodeP& a, const std::function<std::string(const std::string&)>& sym) {
  auto rec = [&](const NodeP& n) { return to_cuda(n, sym); };
  auto b = [&](const NodeP& n) { return cond_text(n, sym); };
  switch (a->op) {
    case Op::Lt: return "(" + rec(a->kids[0]) + " < " + rec(a->kids[1]) + ")";
    case Op::Gt: return "(" + rec(a->kids[0]) + " > " + rec(a->kids[1]) + ")";
    case Op::Le: return "(" + rec(a->kids[0]) + " <= " + rec(a->kids[1]) + ")";
    case Op::Ge: return "(" + rec(a->kids[0]) + " >= " + rec(a->kids[1]) + ")";
    case Op::Eq: return "(" + rec(a->kids[0]) + " == " + rec(a->kids[1]) + ")";
    case Op::Ne: return "(" + rec(a->kids[0]) + " != " + rec(a->kids[1]) + ")";
    case Op::And: return "(" + b(a->kids[0]) + " && " + b(a->kids[1]) + ")";
    case Op::Or: return "(" + b(a->kids[0]) + " || " + b(a->kids[1]) + ")";
    case Op::Not: return "(!" + b(a->kids[0]) + ")";
    default: return "(" + rec(a) + " != 0.0)";
  }
  (void)is_cmp;
}


// ------------------------------------------------------------------------------------------------
// polynomial normal form
namespace {
using Mono = std::vector<int>;
using Poly = std::map<Mono, double>;

struct PolyBuilder {
  std::vector<NodeP>& atoms;
  std::map<std::string, int> index;   // text of the sub-expression -> atom id
  size_t max_terms, max_degree;

  int atom(const NodeP& a) {
    const std::string key = to_text(a);
    auto it = index.find(key);
    if (it != index.end()) return it->second;
    atoms.push_back(a);
    return index[key] = (int)atoms.size() - 1;
  }
  static void add(Poly& p, const Mono& m, double c) {
    if (c == 0.0) return;
    double& v = p[m];
    v += c;
    if (v == 0.0) p.erase(m);
  }
  bool mul(const Poly& a, const Poly& b, Poly& out) {
    out.clear();
    for (auto& ta : a)
      for (auto& tb : b) {
        Mono m(ta.first);
        m.insert(m.end(), tb.first.begin(), tb.first.end());
        std::sort(m.begin(), m.end());
        if (m.size() > max_degree) return false;
        add(out, m, ta.second * tb.second);
        if (out.size() > max_terms) return false;
      }
    return true;
  }
  bool build(const NodeP& a, Poly& out) {
    out.clear();
    switch (a->op) {
      case Op::Num: add(out, {}, a->num); return true;
      case Op::Neg: {
        Poly p;
        if (!build(a->kids[0], p)) return false;
        for (auto& t : p) add(out, t.first, -t.second);
        return true;
      }
      case Op::Add: case Op::Sub: {
        Poly p, q;
        if (!build(a->kids[0], p) || !build(a->kids[1], q)) return false;
        out = p;
        for (auto& t : q) add(out, t.first, a->op == Op::Add ? t.second : -t.second);
        return out.size() <= max_terms;
      }
      case Op::Mul: {
        Poly p, q;
        return build(a->kids[0], p) && build(a->kids[1], q) && mul(p, q, out);
      }
      case Op::Div: {
        double d;
        if (!is_constant(a->kids[1], &d) || d == 0.0) break;
        Poly p;
        if (!build(a->kids[0], p)) return false;
        for (auto& t : p) add(out, t.first, t.second / d);
        return true;
      }
      case Op::Pow: {
        double e;
        if (!is_constant(a->kids[1], &e) || e != std::floor(e) || e < 0 || e > (double)max_degree) break;
        Poly base, acc, tmp;
        if (!build(a->kids[0], base)) return false;
        add(acc, {}, 1.0);
        for (int i = 0; i < (int)e; ++i) {
          if (!mul(acc, base, tmp)) return false;
          acc = tmp;
        }
        out = acc;
        return true;
      }
      default: break;
    }
    add(out, {atom(a)}, 1.0);   // anything else is one opaque factor
    return true;
  }
};
}  // namespace

bool expand_polynomials(const std::vector<NodeP>& exprs, PolyForm& out, size_t max_terms, size_t max_degree) {
  out.atoms.clear();
  out.polys.clear();
  PolyBuilder b{out.atoms, {}, max_terms, max_degree};
  for (auto& e : exprs) {
    Poly p;
    if (!b.build(e, p)) return false;
    out.polys.push_back(std::move(p));
  }
  return true;
}

// ------------------------------------------------------------------------------------------------
// symbolic differentiation
namespace {
NodeP mk_call(const std::string& name, std::vector<NodeP> kids) {
  auto n = std::make_shared<Node>();
  n->op = Op::Call;
  n->name = name;
  n->kids = std::move(kids);
  return fold(n);
}
bool is_num(const NodeP& a, double v) { return a->op == Op::Num && a->num == v; }
// constructors with the algebraic identities that keep derived trees small
NodeP s_neg(const NodeP& a) {
  if (a->op == Op::Num) return mk_num(-a->num);
  if (a->op == Op::Neg) return a->kids[0];
  return mk(Op::Neg, {a});
}
NodeP s_add(const NodeP& a, const NodeP& b) {
  if (is_num(a, 0.0)) return b;
  if (is_num(b, 0.0)) return a;
  return fold(mk(Op::Add, {a, b}));
}
NodeP s_sub(const NodeP& a, const NodeP& b) {
  if (is_num(b, 0.0)) return a;
  if (is_num(a, 0.0)) return s_neg(b);
  return fold(mk(Op::Sub, {a, b}));
}
NodeP s_mul(const NodeP& a, const NodeP& b) {
  if (is_num(a, 0.0) || is_num(b, 0.0)) return mk_num(0.0);
  if (is_num(a, 1.0)) return b;
  if (is_num(b, 1.0)) return a;
  if (is_num(a, -1.0)) return s_neg(b);
  if (is_num(b, -1.0)) return s_neg(a);
  return fold(mk(Op::Mul, {a, b}));
}
NodeP s_div(const NodeP& a, const NodeP& b) {
  if (is_num(a, 0.0)) return mk_num(0.0);
  if (is_num(b, 1.0)) return a;
  return fold(mk(Op::Div, {a, b}));
}
NodeP s_pow(const NodeP& a, const NodeP& b) {
  if (is_num(b, 0.0)) return mk_num(1.0);
  if (is_num(b, 1.0)) return a;
  return fold(mk(Op::Pow, {a, b}));
}
NodeP s_sel(const NodeP& c, const NodeP& a, const NodeP& b) {
  if (is_num(a, 0.0) && is_num(b, 0.0)) return mk_num(0.0);
  return fold(mk(Op::Sel, {c, a, b}));
}

NodeP diff(const NodeP& a, const std::string& var) {
  switch (a->op) {
    case Op::Num: return mk_num(0.0);
    case Op::Var: return mk_num(a->name == var ? 1.0 : 0.0);
    case Op::Neg: return s_neg(diff(a->kids[0], var));
    case Op::Add: return s_add(diff(a->kids[0], var), diff(a->kids[1], var));
    case Op::Sub: return s_sub(diff(a->kids[0], var), diff(a->kids[1], var));
    case Op::Mul:
      return s_add(s_mul(diff(a->kids[0], var), a->kids[1]), s_mul(a->kids[0], diff(a->kids[1], var)));
    case Op::Div: {
      const NodeP &f = a->kids[0], &g = a->kids[1];
      NodeP df = diff(f, var), dg = diff(g, var);
      // f'/g - f g'/g^2
      return s_sub(s_div(df, g), s_div(s_mul(f, dg), s_mul(g, g)));
    }
    case Op::Pow: {
      const NodeP &f = a->kids[0], &g = a->kids[1];
      NodeP df = diff(f, var), dg = diff(g, var);
      double c;
      if (is_constant(g, &c)) return s_mul(s_mul(mk_num(c), s_pow(f, mk_num(c - 1.0))), df);
      // f^g (g' log f + g f'/f)
      return s_mul(a, s_add(s_mul(dg, mk_call("log", {f})), s_div(s_mul(g, df), f)));
    }
    case Op::Mod: return diff(a->kids[0], var);   // almost everywhere, for a constant modulus
    case Op::Not: case Op::Lt: case Op::Gt: case Op::Le: case Op::Ge: case Op::Eq: case Op::Ne:
    case Op::And: case Op::Or:
      return mk_num(0.0);
    case Op::Sel: return s_sel(a->kids[0], diff(a->kids[1], var), diff(a->kids[2], var));
    case Op::Call: {
      const std::string& f = a->name;
      if (a->kids.size() == 1) {
        const NodeP& x = a->kids[0];
        NodeP dx = diff(x, var);
        if (is_num(dx, 0.0)) return dx;
        if (f == "sqrt") return s_div(dx, s_mul(mk_num(2.0), a));
        if (f == "exp") return s_mul(a, dx);
        if (f == "log" || f == "ln") return s_div(dx, x);
        if (f == "log10") return s_div(dx, s_mul(x, mk_num(std::log(10.0))));
        if (f == "log2") return s_div(dx, s_mul(x, mk_num(std::log(2.0))));
        if (f == "exp2") return s_mul(s_mul(a, mk_num(std::log(2.0))), dx);
        if (f == "sin") return s_mul(mk_call("cos", {x}), dx);
        if (f == "cos") return s_neg(s_mul(mk_call("sin", {x}), dx));
        if (f == "tan") { NodeP c = mk_call("cos", {x}); return s_div(dx, s_mul(c, c)); }
        if (f == "tanh") return s_mul(s_sub(mk_num(1.0), s_mul(a, a)), dx);
        if (f == "sinh") return s_mul(mk_call("cosh", {x}), dx);
        if (f == "cosh") return s_mul(mk_call("sinh", {x}), dx);
        if (f == "asin") return s_div(dx, mk_call("sqrt", {s_sub(mk_num(1.0), s_mul(x, x))}));
        if (f == "acos") return s_neg(s_div(dx, mk_call("sqrt", {s_sub(mk_num(1.0), s_mul(x, x))})));
        if (f == "atan") return s_div(dx, s_add(mk_num(1.0), s_mul(x, x)));
        if (f == "abs") return s_mul(mk_call("sgn", {x}), dx);
        if (f == "floor" || f == "ceil" || f == "round" || f == "sgn" || f == "sign") return mk_num(0.0);
      } else if (a->kids.size() == 2) {
        const NodeP &x = a->kids[0], &y = a->kids[1];
        if (f == "pow") return diff(mk(Op::Pow, {x, y}), var);
        if (f == "min") return s_sel(mk(Op::Lt, {x, y}), diff(x, var), diff(y, var));
        if (f == "max") return s_sel(mk(Op::Gt, {x, y}), diff(x, var), diff(y, var));
        if (f == "atan2") {   // d atan2(x, y) = (y x' - x y') / (x^2 + y^2)
          NodeP den = s_add(s_mul(x, x), s_mul(y, y));
          return s_div(s_sub(s_mul(y, diff(x, var)), s_mul(x, diff(y, var))), den);
        }
      } else if (f == "min" || f == "max") {   // n-ary: fold pairwise like the evaluator does
        NodeP acc = a->kids[0];
        for (size_t i = 1; i < a->kids.size(); ++i) {
          auto n = std::make_shared<Node>();
          n->op = Op::Call; n->name = f; n->kids = {acc, a->kids[i]};
          acc = n;
        }
        return diff(acc, var);
      }
      if (f.rfind("dc_tab_", 0) == 0)
        fail("cannot differentiate a tabulated context function symbolically: give the jacobian entries in the ini");
      fail("cannot differentiate function '", f, "' with ", a->kids.size(), " argument(s)");
    }
  }
  fail("internal: differentiate");
}
}  // namespace

NodeP differentiate(const NodeP& ast, const std::string& var) { return diff(ast, var); }
bool is_zero(const NodeP& ast) { return ast->op == Op::Num && ast->num == 0.0; }

std::string to_text(const NodeP& ast) {
  return to_cuda(ast, [](const std::string& name) { return name; });
}

}  // namespace dcb
