// Native server-side optimizers (host fp32): the declarative `Optimizer.spec()` executed without Python on the (global) server.
// Parity: the reference ships a pickled Python optimizer to the server and runs `Updater.__call__` on the server's main thread through
// `Executor` (python/mxnet/kvstore.py:452-499, kvstore_dist_server.h:109-168, optimizer ops src/operator/optimizer_op-inl.h).
// Optimizers without a native spec still use that host-callback path (see KVStoreDistServer::set_updater).
#pragma once
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
    if (st.a.size() != n) { st.a.assign(n, 0.f); st.b.assign(n, 0.f); st.t = 0; if (s_.name == "dcasgd") for (size_t i = 0; i < n; ++i) st.b[i] = w[i]; }
    ++st.t;
    const float lr = s_.name == "adam" ? s_.lr * std::sqrt(1.f - std::pow(s_.beta2, (float)st.t)) / (1.f - std::pow(s_.beta1, (float)st.t)) : s_.lr;
    for (size_t i = 0; i < n; ++i) {
      float gi = g[i] * s_.rescale;
      if (s_.name == "adam") gi += s_.wd * w[i];      // adam_update clips grad + wd*w (optimizer_op-inl.h:840-873); sgd / dcasgd clip the raw gradient
      if (s_.clip >= 0.f) gi = std::fmin(std::fmax(gi, -s_.clip), s_.clip);
      if (s_.name == "adam") {
        st.a[i] = s_.beta1 * st.a[i] + (1.f - s_.beta1) * gi;
        st.b[i] = s_.beta2 * st.b[i] + (1.f - s_.beta2) * gi * gi;
        w[i] -= lr * st.a[i] / (std::sqrt(st.b[i]) + s_.eps);
      } else if (s_.name == "sgd") {
        gi += s_.wd * w[i];
        if (s_.momentum != 0.f) { st.a[i] = s_.momentum * st.a[i] - lr * gi; w[i] += st.a[i]; }
        else w[i] -= lr * gi;
      } else {  // dcasgd: one previous_weight per key (not per party), as in the reference
        const float upd = gi + s_.wd * w[i] + s_.lamda * gi * gi * (w[i] - st.b[i]);
        const float prev = w[i];
        if (s_.momentum != 0.f) { st.a[i] = s_.momentum * st.a[i] - lr * upd; w[i] += st.a[i]; }
        else w[i] -= lr * upd;
        st.b[i] = prev;
      }
    }
  }

 private:
  OptSpec s_;
};

}  // namespace hips
