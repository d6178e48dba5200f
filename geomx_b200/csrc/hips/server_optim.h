// Native server-side optimizers (host fp32): the declarative `Optimizer.spec()` executed without Python on the (global) server.
// Parity: the reference ships a pickled Python optimizer to the server and runs `Updater.__call__` on the server's main thread through
// `Executor` (python/mxnet/kvstore.py:452-499, kvstore_dist_server.h:109-168, optimizer ops src/operator/optimizer_op-inl.h).
// Optimizers without a native spec still use that host-callback path (see KVStoreDistServer::set_updater).
#pragma once
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <cmath>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

namespace hips {

struct OptSpec {
  std::string name;  // sgd | adam | dcasgd
  float lr = 0.01f, wd = 0.f, rescale = 1.f, clip = -1.f, momentum = 0.f, beta1 = 0.9f, beta2 = 0.999f, eps = 1e-8f, lamda = 0.04f;
  bool valid() const { return name == "sgd" || name == "adam" || name == "dcasgd"; }
  // "name=adam;lr=0.01;beta1=0.9;..."
  static OptSpec Parse(const std::string& s) {
    OptSpec o;
    size_t p = 0;
    while (p < s.size()) {
      size_t e = s.find(';', p); if (e == std::string::npos) e = s.size();
      const std::string kv = s.substr(p, e - p);
      const size_t eq = kv.find('=');
      if (eq != std::string::npos) {
        const std::string k = kv.substr(0, eq), v = kv.substr(eq + 1);
        if (k == "name") o.name = v;
        else {
          const float f = static_cast<float>(atof(v.c_str()));
          if (k == "lr") o.lr = f; else if (k == "wd") o.wd = f; else if (k == "rescale_grad") o.rescale = f;
          else if (k == "clip_gradient") o.clip = f; else if (k == "momentum") o.momentum = f; else if (k == "beta1") o.beta1 = f;
          else if (k == "beta2") o.beta2 = f; else if (k == "epsilon") o.eps = f; else if (k == "lamda") o.lamda = f;
        }
      }
      p = e + 1;
    }
    return o;
  }
};

// Element-wise work on big tensors (optimizer steps, aggregation) split over a few threads: GEOMX_SERVER_THREADS (default: half the cores,
// at most 4); ranges below `grain` elements per thread stay on the calling thread, so the small keys of a model never pay a thread start.
inline int ServerThreads() {
  static const int t = [] {
    const char* v = getenv("GEOMX_SERVER_THREADS");
    const unsigned hw = std::thread::hardware_concurrency();
    const int n = v && *v ? atoi(v) : static_cast<int>(std::min(4u, std::max(1u, hw / 2)));
    return std::max(1, n);
  }();
  return t;
}
template <typename F>
inline void ParallelFor(size_t n, size_t grain, F&& f) {
  const size_t T = std::min<size_t>(static_cast<size_t>(ServerThreads()), n / grain);
  if (T <= 1) { f(size_t(0), n); return; }
  const size_t chunk = ((n + T - 1) / T + 15) & ~size_t(15);
  std::vector<std::thread> th;
  for (size_t t = 1; t < T; ++t) {
    const size_t lo = t * chunk, hi = std::min(n, lo + chunk);
    if (lo < hi) th.emplace_back([&f, lo, hi] { f(lo, hi); });
  }
  f(size_t(0), std::min(n, chunk));
  for (auto& x : th) x.join();
}

class NativeOptimizer {
 public:
  explicit NativeOptimizer(const OptSpec& s) : s_(s) {}
  const OptSpec& spec() const { return s_; }
  // checkpointable server-side state of ONE key (the reference cannot save it: "Cannot save states for distributed training"); it lives with
  // the key on the server, so updates of different keys never share a container
  struct State { std::vector<float> a, b; int t = 0; };
  // weight (fp32 master) updated in place from grad; the state is initialised on first use
  void Update(State* state, float* w, const float* g, size_t n) const {
    State& st = *state;
    const int kind = s_.name == "adam" ? 0 : s_.name == "sgd" ? 1 : 2;      // resolved once per call, not per element
    if (st.a.size() != n) { st.a.assign(n, 0.f); st.b.assign(n, 0.f); st.t = 0; if (kind == 2) memcpy(st.b.data(), w, n * sizeof(float)); }
    ++st.t;
    const float lr = kind == 0 ? s_.lr * std::sqrt(1.f - std::pow(s_.beta2, (float)st.t)) / (1.f - std::pow(s_.beta1, (float)st.t)) : s_.lr;
    float* a = st.a.data(); float* b = st.b.data();
    const bool mom = s_.momentum != 0.f;
    // big tensors are updated by a few threads over disjoint ranges (element-wise arithmetic: the result does not depend on the split)
    ParallelFor(n, size_t(1) << 18, [=](size_t lo, size_t hi) {
      if (kind == 0) UpdateRange<0>(lo, hi, w, g, a, b, lr, mom);
      else if (kind == 1) UpdateRange<1>(lo, hi, w, g, a, b, lr, mom);
      else UpdateRange<2>(lo, hi, w, g, a, b, lr, mom);
    });
  }

 private:
  template <int KIND>
  void UpdateRange(size_t lo, size_t hi, float* __restrict__ w, const float* __restrict__ g, float* __restrict__ a, float* __restrict__ b, float lr, bool mom) const {
    const float rescale = s_.rescale, wd = s_.wd, clip = s_.clip, beta1 = s_.beta1, beta2 = s_.beta2, eps = s_.eps, momentum = s_.momentum, lamda = s_.lamda;
    for (size_t i = lo; i < hi; ++i) {
      float gi = g[i] * rescale;
      if (KIND == 0) gi += wd * w[i];      // adam_update clips grad + wd*w (optimizer_op-inl.h:840-873); sgd / dcasgd clip the raw gradient
      if (clip >= 0.f) gi = std::fmin(std::fmax(gi, -clip), clip);
      if (KIND == 0) {
        a[i] = beta1 * a[i] + (1.f - beta1) * gi;
        b[i] = beta2 * b[i] + (1.f - beta2) * gi * gi;
        w[i] -= lr * a[i] / (std::sqrt(b[i]) + eps);
      } else if (KIND == 1) {
        gi += wd * w[i];
        if (mom) { a[i] = momentum * a[i] - lr * gi; w[i] += a[i]; }
        else w[i] -= lr * gi;
      } else {  // dcasgd: one previous_weight per key (not per party), as in the reference
        const float upd = gi + wd * w[i] + lamda * gi * gi * (w[i] - b[i]);
        const float prev = w[i];
        if (mom) { a[i] = momentum * a[i] - lr * upd; w[i] += a[i]; }
        else w[i] -= lr * upd;
        b[i] = prev;
      }
    }
  }

  OptSpec s_;
};

}  // namespace hips
