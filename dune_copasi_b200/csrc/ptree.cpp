#include "ptree.hpp"

#include <algorithm>
#include <cstdlib>
#include <sstream>

namespace dcb {

std::string trim(const std::string& s) {
  size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
  return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}

static const PTree kEmpty;

bool PTree::has_key(const std::string& key) const {
  auto dot = key.find('.');
  if (dot == std::string::npos) return vals_.count(key) != 0;
  auto it = subs_.find(key.substr(0, dot));
  return it != subs_.end() && it->second->has_key(key.substr(dot + 1));
}

bool PTree::has_sub(const std::string& key) const {
  auto dot = key.find('.');
  auto it = subs_.find(key.substr(0, dot));
  if (it == subs_.end()) return false;
  return dot == std::string::npos ? true : it->second->has_sub(key.substr(dot + 1));
}

const PTree& PTree::sub(const std::string& key) const {
  auto dot = key.find('.');
  auto it = subs_.find(key.substr(0, dot));
  if (it == subs_.end()) return kEmpty;
  return dot == std::string::npos ? *it->second : it->second->sub(key.substr(dot + 1));
}

PTree& PTree::sub_mut(const std::string& key) {
  auto dot = key.find('.');
  std::string head = key.substr(0, dot);
  auto it = subs_.find(head);
  if (it == subs_.end()) {
    it = subs_.emplace(head, std::make_unique<PTree>()).first;
    sub_order_.push_back(head);
  }
  return dot == std::string::npos ? *it->second : it->second->sub_mut(key.substr(dot + 1));
}

std::string PTree::get(const std::string& key, const std::string& def) const {
  auto dot = key.find('.');
  if (dot == std::string::npos) {
    auto it = vals_.find(key);
    return it == vals_.end() ? def : it->second;
  }
  auto it = subs_.find(key.substr(0, dot));
  return it == subs_.end() ? def : it->second->get(key.substr(dot + 1), def);
}

double PTree::get(const std::string& key, double def) const {
  std::string s = get(key, std::string());
  if (s.empty()) return def;
  char* end = nullptr;
  double v = std::strtod(s.c_str(), &end);
  if (end == s.c_str()) fail("config key '", key, "': cannot parse number from '", s, "'");
  return v;
}

int PTree::get(const std::string& key, int def) const {
  std::string s = get(key, std::string());
  return s.empty() ? def : (int)std::strtol(s.c_str(), nullptr, 10);
}

bool PTree::get(const std::string& key, bool def) const {
  std::string s = get(key, std::string());
  if (s.empty()) return def;
  std::transform(s.begin(), s.end(), s.begin(), ::tolower);
  return s == "true" || s == "1" || s == "yes" || s == "on";
}

std::vector<double> PTree::get_vec(const std::string& key, const std::vector<double>& def) const {
  std::string s = get(key, std::string());
  if (s.empty()) return def;
  std::vector<double> out;
  std::istringstream is(s);
  double v;
  while (is >> v) out.push_back(v);
  return out;
}

void PTree::set(const std::string& key, const std::string& val) {
  auto dot = key.rfind('.');
  PTree& t = dot == std::string::npos ? *this : sub_mut(key.substr(0, dot));
  std::string leaf = dot == std::string::npos ? key : key.substr(dot + 1);
  if (!t.vals_.count(leaf)) t.val_order_.push_back(leaf);
  t.vals_[leaf] = val;
}

void PTree::parse_ini(const std::string& text) {
  std::istringstream is(text);
  std::string line, prefix;
  while (std::getline(is, line)) {
    auto hash = line.find('#');
    if (hash != std::string::npos) line = line.substr(0, hash);
    line = trim(line);
    if (line.empty()) continue;
    if (line.front() == '[' && line.back() == ']') {
      prefix = trim(line.substr(1, line.size() - 2));
      if (!prefix.empty()) sub_mut(prefix);
      continue;
    }
    auto eq = line.find('=');
    if (eq == std::string::npos) continue;
    std::string key = trim(line.substr(0, eq)), val = trim(line.substr(eq + 1));
    set(prefix.empty() ? key : prefix + "." + key, val);
  }
}

std::string PTree::dump(const std::string& prefix) const {
  std::string out;
  for (auto& k : val_order_) out += prefix + k + " = " + vals_.at(k) + "\n";
  for (auto& k : sub_order_) out += subs_.at(k)->dump(prefix + k + ".");
  return out;
}

}  // namespace dcb
