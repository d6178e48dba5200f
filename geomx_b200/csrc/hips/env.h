// Environment lookup with in-process override map (parity: ps-lite include/ps/internal/env.h:15-64; MXInitPSEnv → Environment::Init).
#pragma once
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>

namespace hips {

class Environment {
 public:
  static Environment* Get() { static Environment e; return &e; }
  void Set(const std::string& k, const std::string& v) { std::lock_guard<std::mutex> lk(mu_); kvs_[k] = v; }
  const char* find(const char* k) {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = kvs_.find(k);
    if (it != kvs_.end()) return it->second.c_str();
    return getenv(k);
  }
  int GetInt(const char* k, int dflt) { const char* v = find(k); return (v && *v) ? atoi(v) : dflt; }
  double GetFloat(const char* k, double dflt) { const char* v = find(k); return (v && *v) ? atof(v) : dflt; }
  std::string GetStr(const char* k, const std::string& dflt) { const char* v = find(k); return (v && *v) ? std::string(v) : dflt; }

 private:
  std::mutex mu_;
  std::map<std::string, std::string> kvs_;
};

}  // namespace hips
