#include "reduce.hpp"

#include <algorithm>
#include <cmath>
#include <sstream>

#include "comm.hpp"

namespace dcb {

extern const char* kKernelArgsSource;
extern const char* kAssemblySource;
extern const char* kReduceSource;

namespace {

// FloatCmp::ne(value, 0.) with dune-common's default relative epsilon (reduce.hh:238, 258)
bool nonzero(double v) { return std::fabs(v) > 1e-8 * std::max(1.0, std::fabs(v)); }

}  // namespace

std::vector<Reducer::Key> Reducer::parse_keys(const Model& m, const PTree& cfg) {
  std::vector<Key> keys_;
  const PTree& rc = cfg.sub("model.reduce");
  for (auto& name : rc.sub_keys()) {
    const PTree& s = rc.sub(name);
    Key k;
    k.name = name;
    k.initial = s.get("initial.value", 0.0);
    std::string ev = s.get("evaluation.expression", std::string());
    if (!expr_is_absent(ev)) k.evaluation = m.compile(ev);
    if (s.has_key("reduction.expression")) {
      auto fn = parse_function_expression(s.get("reduction.expression", std::string()), "model.reduce." + name + ".reduction");
      if (fn.args.size() != 2) fail("Reduction arguments must be exactly 2");   // reduce.hh:133-134
      k.has_reduction = true;
      k.ra = fn.args[0]; k.rb = fn.args[1];
      k.reduction = m.compile(fn.body);
    }
    auto fn1 = [&](const char* what, const char* msg, Fn1& out) {
      std::string key = std::string(what) + ".expression";
      if (!s.has_key(key)) return;
      auto fn = parse_function_expression(s.get(key, std::string()), "model.reduce." + name + "." + what);
      if (fn.args.size() != 1) fail(msg, " function must have exactly 1 argument");
      out.present = true;
      out.arg = fn.args[0];
      out.text = fn.body;
      out.ast = m.compile(fn.body);
    };
    fn1("transformation", "Warning", k.transformation);   // message as reduce.hh:223
    fn1("error", "Error", k.error);
    fn1("warn", "Warning", k.warn);
    keys_.push_back(std::move(k));
  }
  return keys_;
}

void Reducer::precompile(const Model& m, const PTree& cfg) {
  auto keys = parse_keys(m, cfg);
  if (!keys.empty()) jit_compile_cached(source(m, keys));
}

Reducer::Reducer(std::shared_ptr<DeviceOperator> op, const PTree& cfg, Communicator* comm)
    : op_(std::move(op)), comm_(comm) {
  const Model& m = *op_->model;
  const Grid& g = *op_->grid;
  keys_ = parse_keys(m, cfg);
  if (keys_.empty()) return;
  require_device();

  // ---- elements per compartment; cells outside every compartment are visited too (their species
  //      read 0).  On a partitioned grid an element counts on the rank that owns its vertex with
  //      the smallest global id.
  const int ncomp = m.ncomp(), nd = g.nd();
  std::vector<std::vector<int>> ids(ncomp + 1);
  for (int64_t e = 0; e < g.ne; ++e) {
    if (g.n_owned >= 0 && !g.global_vid.empty()) {
      int64_t best = -1, bv = 0;
      for (int a = 0; a < nd; ++a) {
        int64_t v = g.elems[e * nd + a];
        if (best < 0 || g.global_vid[v] < best) { best = g.global_vid[v]; bv = v; }
      }
      if (!g.owns(bv)) continue;
    }
    int c = g.elem_comp[e];
    ids[c >= 0 ? c : ncomp].push_back((int)e);
  }
  elem_ids_.resize(ncomp + 1);
  nelem_.assign(ncomp + 1, 0);
  for (int c = 0; c <= ncomp; ++c) {
    nelem_[c] = (int64_t)ids[c].size();
    if (!ids[c].empty()) elem_ids_[c].upload(ids[c], op_->stream);
  }
  max_blocks_ = 148 * 8;
  partials_.alloc((size_t)max_blocks_ * keys_.size());
  std::vector<double> init(keys_.size());
  for (size_t k = 0; k < keys_.size(); ++k) init[k] = keys_[k].initial;
  init_.upload(init, op_->stream);
  DCB_CUDA(cudaStreamSynchronize(op_->stream));
  jit_.load(jit_compile_cached(cuda_source()));
}

std::string Reducer::cuda_source() const { return source(*op_->model, keys_); }

std::string Reducer::source(const Model& m, const std::vector<Key>& keys_) {
  std::ostringstream o;
  o << m.cuda_source() << kKernelArgsSource << "\n" << kAssemblySource << "\n";
  o << "// ---- [model.reduce] -------------------------------------------------------------------\n";
  o << "#define DC_NRED " << keys_.size() << "\n";
  o << "template <int C> struct DcReduce;\n";
  for (int c = 0; c <= m.ncomp(); ++c) {
    int ns = c < m.ncomp() ? m.comp_nspec[c] : 0;
    o << "template <> struct DcReduce<" << c << "> {\n  static constexpr int NS = " << std::max(ns, 1) << ", NS_REAL = " << ns << ";\n";
    o << "  __device__ __forceinline__ static void eval(const DcCtx& c, const double* u, const double (*g)[DC_DIM], double* out) {\n";
    for (size_t k = 0; k < keys_.size(); ++k)
      o << "    out[" << k << "] = " << (keys_[k].evaluation ? m.lower_volume(keys_[k].evaluation, c) : std::string("0.0")) << ";\n";
    o << "    (void)c; (void)u; (void)g;\n  }\n};\n";
  }
  o << "// value = reduction(evaluation(), value)   (reduce.hh:187-188)\n";
  o << "__device__ __forceinline__ double dc_reduce_op(int k, double a, double b) {\n  switch (k) {\n";
  for (size_t k = 0; k < keys_.size(); ++k) {
    const Key& K = keys_[k];
    o << "    case " << k << ": return ";
    if (!K.has_reduction) o << "a + b";
    else
      o << to_cuda(K.reduction, [&](const std::string& n) -> std::string {
        if (n == K.ra) return "a";
        if (n == K.rb) return "b";
        return "";
      });
    o << ";\n";
  }
  o << "  }\n  return 0.0;\n}\n";
  o << kReduceSource << "\n";
  for (int c = 0; c <= m.ncomp(); ++c)
    o << "extern \"C\" __global__ void __launch_bounds__(DC_RED_THREADS) dc_k_reduce_" << c
      << "(DcReduceArgs a) { dc_reduce_kernel<" << c << ">(a); }\n"
      << "extern \"C\" __global__ void __launch_bounds__(DC_RED_THREADS) dc_k_reduce_q1_" << c
      << "(DcReduceArgs a) { dc_reduce_q1_kernel<" << c << ">(a); }\n";
  return o.str();
}

double Reducer::fold(const Key& k, double a, double b) const {
  if (!k.has_reduction) return a + b;
  return eval_expr(k.reduction, [&](const std::string& n) -> double {
    if (n == k.ra) return a;
    if (n == k.rb) return b;
    fail("unknown symbol '", n, "' in model.reduce.", k.name, ".reduction.expression");
  });
}

std::vector<ReduceEntry> Reducer::apply(double time, const double* x, bool throw_on_error) {
  std::vector<ReduceEntry> out;
  if (keys_.empty()) return out;
  const Model& m = *op_->model;
  const Grid& g = *op_->grid;
  cudaStream_t s = op_->stream;
  const int nk = (int)keys_.size();
  std::vector<double> values(nk), host;
  for (int k = 0; k < nk; ++k) values[k] = keys_[k].initial;
  for (int c = 0; c <= m.ncomp(); ++c) {
    if (nelem_[c] == 0) continue;
    DcReduceArgs a{};
    a.coords = op_->coords_.p; a.elems = op_->elems_.p; a.elem_ids = elem_ids_[c].p;
    a.vdof = c < m.ncomp() ? op_->comp_vdof_[c].p : nullptr;
    a.cell = op_->cell_.p; a.ne_total = g.ne; a.n = nelem_[c];
    a.dof_offset = c < m.ncomp() ? (int)g.comp_offset[c] : 0;
    a.time = time; a.x = x; a.init = init_.p; a.partials = partials_.p;
    const int blocks = (int)std::min<int64_t>(max_blocks_, (nelem_[c] + 127) / 128);
    {
      DeviceOperator::ProfScope ps(op_.get(), "reduce");
      jit_launch(jit_.kernel((g.elem_kind == 1 ? "dc_k_reduce_q1_" : "dc_k_reduce_") + std::to_string(c)), blocks, 128, 0, s, a);
      op_->stats.launches++;
    }
    host.resize((size_t)blocks * nk);
    DCB_CUDA(cudaMemcpyAsync(host.data(), partials_.p, sizeof(double) * host.size(), cudaMemcpyDeviceToHost, s));
    DCB_CUDA(cudaStreamSynchronize(s));
    for (int b = 0; b < blocks; ++b)
      for (int k = 0; k < nk; ++k) values[k] = fold(keys_[k], host[(size_t)b * nk + k], values[k]);
  }
  if (comm_ && comm_->size > 1) {
    // all-gather through a sum over disjoint slots, then fold in rank order
    const int np = comm_->size;
    std::vector<double> slots((size_t)np * nk, 0.0);
    for (int k = 0; k < nk; ++k) slots[(size_t)comm_->rank * nk + k] = values[k];
    gather_.upload(slots, s);
    comm_->allreduce_sum(gather_.p, np * nk, s);
    gather_.download(slots.data(), s);
    for (int k = 0; k < nk; ++k) {
      double v = keys_[k].initial;
      for (int r = 0; r < np; ++r) v = fold(keys_[k], slots[(size_t)r * nk + k], v);
      values[k] = v;
    }
  }
  std::string error_msg;
  for (int k = 0; k < nk; ++k) {
    const Key& K = keys_[k];
    ReduceEntry e;
    e.key = K.name;
    double v = values[k];
    auto call = [&](const Fn1& f, double arg) {
      return eval_expr(f.ast, [&](const std::string& n) -> double {
        if (n == f.arg) return arg;
        fail("unknown symbol '", n, "' in model.reduce.", K.name);
      });
    };
    if (K.transformation.present) v = call(K.transformation, v);
    e.value = v;
    if (K.error.present && nonzero(call(K.error, v))) {
      e.status = 2;
      if (!error_msg.empty()) error_msg += '\n';
      std::ostringstream msg;
      msg.precision(17);
      msg << "Reduction on the token '" << K.name << "' raised an error because the expression '" << K.error.text
          << "' with evaluates to false with '" << K.error.arg << " := " << v << "'";   // reduce.hh:241-247
      error_msg += msg.str();
    } else if (K.warn.present && nonzero(call(K.warn, v))) {
      e.status = 1;
    }
    out.push_back(e);
  }
  last_error = error_msg;
  if (throw_on_error && !error_msg.empty()) fail(error_msg);
  return out;
}

}  // namespace dcb
