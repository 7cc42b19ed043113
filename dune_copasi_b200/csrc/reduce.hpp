// [model.reduce]: user functionals over the grid, evaluated on the device.
//
// Mirrors dune/copasi/model/diffusion_reaction/reduce.hh:38-285 -- per key `evaluation.expression`
// (with every contextual symbol of the volume terms, incl. integration_factor), a two-argument
// `reduction.expression` (default: sum), `initial.value`, then on the result
// `transformation.expression`, `error.expression` (non-zero => the reduce fails) and
// `warn.expression`.  It is how the reference's system tests assert their known answers
// (test/*.ini) and what users monitor after every step (src/dune_copasi.cc:420-423).
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "operator.hpp"

namespace dcb {

struct Communicator;

struct ReduceEntry {
  std::string key;
  double value = 0.0;
  int status = 0;   // 0: fine, 1: warn.expression fired, 2: error.expression fired
};

class Reducer {
 public:
  // cfg = the whole ini; only model.reduce.* (and model.parser_type) is read
  Reducer(std::shared_ptr<DeviceOperator> op, const PTree& cfg, Communicator* comm = nullptr);

  int size() const { return (int)keys_.size(); }
  const std::string& key(int k) const { return keys_[k].name; }
  // evaluates every functional for the device-resident coefficients x at `time`; throws with the
  // reference's message when an error expression fires and `throw_on_error`
  std::vector<ReduceEntry> apply(double time, const double* x, bool throw_on_error = true);
  std::string cuda_source() const;
  std::string last_error;   // message of the error expressions that fired in the last apply ("" if none)
  // compile the reduce kernels of (model, cfg) into the on-disk JIT cache; needs no GPU
  static void precompile(const Model& model, const PTree& cfg);

 private:
  struct Fn1 { bool present = false; std::string arg, text; NodeP ast; };
  struct Key {
    std::string name;
    NodeP evaluation;          // null: evaluates to 0
    bool has_reduction = false;
    std::string ra, rb;        // argument names of the reduction
    NodeP reduction;
    double initial = 0.0;
    Fn1 transformation, error, warn;
  };
  static std::vector<Key> parse_keys(const Model& m, const PTree& cfg);
  static std::string source(const Model& m, const std::vector<Key>& keys);
  double fold(const Key& k, double a, double b) const;
  std::shared_ptr<DeviceOperator> op_;
  Communicator* comm_;
  std::vector<Key> keys_;
  JitModule jit_;
  std::vector<DeviceBuffer<int>> elem_ids_;   // per compartment, last = cells outside all compartments
  std::vector<int64_t> nelem_;
  DeviceBuffer<double> init_, partials_, gather_;
  int max_blocks_ = 0;
};

}  // namespace dcb
