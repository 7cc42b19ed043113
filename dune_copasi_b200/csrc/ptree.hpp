// Ordered parameter tree with the semantics of Dune::ParameterTree as the reference uses it:
// INI text + "--key=value" overrides (src/dune_copasi.cc:270-282); dotted keys address sub-trees;
// sub-key order is insertion order (compartment ids / species order depend on it,
// dune/copasi/grid/make_multi_domain_grid.hh:118-124); unknown keys are ignored.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "util.hpp"

namespace dcb {

class PTree {
 public:
  bool has_key(const std::string& key) const;
  bool has_sub(const std::string& key) const;
  const PTree& sub(const std::string& key) const;  // empty tree if absent
  PTree& sub_mut(const std::string& key);
  std::string get(const std::string& key, const std::string& def) const;
  double get(const std::string& key, double def) const;
  int get(const std::string& key, int def) const;
  bool get(const std::string& key, bool def) const;
  std::vector<double> get_vec(const std::string& key, const std::vector<double>& def) const;
  void set(const std::string& key, const std::string& val);
  const std::vector<std::string>& sub_keys() const { return sub_order_; }
  const std::vector<std::string>& value_keys() const { return val_order_; }
  void parse_ini(const std::string& text);
  std::string dump(const std::string& prefix = "") const;

 private:
  std::map<std::string, std::string> vals_;
  std::vector<std::string> val_order_;
  std::map<std::string, std::unique_ptr<PTree>> subs_;
  std::vector<std::string> sub_order_;
};

std::string trim(const std::string& s);

}  // namespace dcb
