// Host training executor over the native graph (graph.h): binds argument / gradient / auxiliary arrays, runs forward (training or inference
// mode) and backward in fp32 on the CPU.  It is what GXExecutor* and the GXAutograd* / GXImperativeInvoke C API run on (c_api_graph.cc), i.e.
// the path a non-Python front end uses to TRAIN through this framework without PyTorch in the process.
//
// Parity: include/mxnet/c_api.h:1530-1760 (MXExecutorBind / BindX / BindEX / SimpleBind, Forward, Backward(Ex), Outputs, Print, Free) over
// src/executor/graph_executor.cc (Init / InitArguments / Forward / Backward / RunOps) and the operator gradients registered with FGradient
// (src/operator/nn/*.cc, src/operator/softmax_output-inl.h, regression_output-inl.h).  The reference builds a separate backward graph with
// nnvm::pass::Gradient and plans memory for both; here the forward activations are kept per node and the backward pass is a reverse sweep
// that calls one gradient routine per operator — on the host the simplicity is worth more than the reuse (the device path's equivalents are
// the fused sm_100a kernels and the CUDA-graph executor of models/cnn.py).
#pragma once
#include <atomic>
#include <cstring>
#include <mutex>
#include <random>

#include "graph.h"

namespace gxrt {
namespace exec {

using graph::AttrView;
using graph::Entry;
using graph::Node;
using graph::Symbol;
using predict::Numel;
using predict::ParallelFor;
using predict::Shape;
using predict::ShapeStr;

enum GradReq { kNullOp = 0, kWriteTo = 1, kWriteInplace = 2, kAddTo = 3 };

// ------------------------------------------------------------------------------------------------ dense kernels
// serial C[M,N] (+)= op(A) . op(B); the j-inner loops vectorise.  A is [M,K] (ta = false) or [K,M]; B is [K,N] (tb = false) or [N,K].
inline void GemmSerial(bool ta, bool tb, int64_t M, int64_t N, int64_t K, const float* A, const float* B, float* C, bool accumulate) {
  if (!accumulate) std::fill(C, C + M * N, 0.f);
  if (!tb) {
    for (int64_t i = 0; i < M; ++i) {
      float* c = C + i * N;
      for (int64_t k = 0; k < K; ++k) {
        const float a = ta ? A[k * M + i] : A[i * K + k];
        if (a == 0.f) continue;
        const float* b = B + k * N;
        for (int64_t j = 0; j < N; ++j) c[j] += a * b[j];
      }
    }
  } else if (!ta) {
    for (int64_t i = 0; i < M; ++i) for (int64_t j = 0; j < N; ++j) {
      const float* a = A + i * K; const float* b = B + j * K;
      float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      int64_t k = 0;
      for (; k + 8 <= K; k += 8) for (int u = 0; u < 8; ++u) acc[u] += a[k + u] * b[k + u];
      float s = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
      for (; k < K; ++k) s += a[k] * b[k];
      C[i * N + j] += s;
    }
  } else {
    for (int64_t i = 0; i < M; ++i) for (int64_t j = 0; j < N; ++j) {
      float s = 0;
      for (int64_t k = 0; k < K; ++k) s += A[k * M + i] * B[j * K + k];
      C[i * N + j] += s;
    }
  }
}
// rows of C split over threads
inline void Gemm(bool ta, bool tb, int64_t M, int64_t N, int64_t K, const float* A, const float* B, float* C, bool accumulate) {
  ParallelFor(M, static_cast<double>(M) * N * K, [&](int64_t lo, int64_t hi) {
    if (!ta) GemmSerial(false, tb, hi - lo, N, K, A + lo * K, B, C + lo * N, accumulate);
    else {
      // A^T rows lo..hi are columns of A: walk them in place
      for (int64_t i = lo; i < hi; ++i) {
        float* c = C + i * N;
        if (!accumulate) std::fill(c, c + N, 0.f);
        for (int64_t k = 0; k < K; ++k) {
          const float a = A[k * M + i];
          if (a == 0.f) continue;
          if (!tb) { const float* b = B + k * N; for (int64_t j = 0; j < N; ++j) c[j] += a * b[j]; }
          else for (int64_t j = 0; j < N; ++j) c[j] += a * B[j * K + k];
        }
      }
    }
  });
}

struct Win { int64_t kh, kw, sh, sw, ph, pw, dh, dw; };

inline void Im2Col(const float* x, int64_t C, int64_t H, int64_t W, const Win& w, int64_t OH, int64_t OW, float* col) {
  for (int64_t c = 0; c < C; ++c) for (int64_t r = 0; r < w.kh; ++r) for (int64_t s = 0; s < w.kw; ++s) {
    float* row = col + ((c * w.kh + r) * w.kw + s) * OH * OW;
    for (int64_t oy = 0; oy < OH; ++oy) {
      const int64_t iy = oy * w.sh - w.ph + r * w.dh;
      if (iy < 0 || iy >= H) { std::fill(row + oy * OW, row + (oy + 1) * OW, 0.f); continue; }
      for (int64_t ox = 0; ox < OW; ++ox) {
        const int64_t ix = ox * w.sw - w.pw + s * w.dw;
        row[oy * OW + ox] = (ix >= 0 && ix < W) ? x[(c * H + iy) * W + ix] : 0.f;
      }
    }
  }
}
inline void Col2Im(const float* col, int64_t C, int64_t H, int64_t W, const Win& w, int64_t OH, int64_t OW, float* dx) {
  for (int64_t c = 0; c < C; ++c) for (int64_t r = 0; r < w.kh; ++r) for (int64_t s = 0; s < w.kw; ++s) {
    const float* row = col + ((c * w.kh + r) * w.kw + s) * OH * OW;
    for (int64_t oy = 0; oy < OH; ++oy) {
      const int64_t iy = oy * w.sh - w.ph + r * w.dh;
      if (iy < 0 || iy >= H) continue;
      for (int64_t ox = 0; ox < OW; ++ox) {
        const int64_t ix = ox * w.sw - w.pw + s * w.dw;
        if (ix >= 0 && ix < W) dx[(c * H + iy) * W + ix] += row[oy * OW + ox];
      }
    }
  }
}

inline float ActF(int kind, float v) {
  switch (kind) {
    case 0: return v > 0 ? v : 0;
    case 1: return 1.f / (1.f + std::exp(-v));
    case 2: return std::tanh(v);
    case 3: return v > 20.f ? v : std::log1p(std::exp(v));      // softrelu
    case 4: return v / (1.f + std::fabs(v));                    // softsign
    default: return v;
  }
}
// derivative from input x and output y
inline float ActG(int kind, float x, float y) {
  switch (kind) {
    case 0: return x > 0 ? 1.f : 0.f;
    case 1: return y * (1.f - y);
    case 2: return 1.f - y * y;
    case 3: return 1.f / (1.f + std::exp(-x));
    case 4: { const float d = 1.f + std::fabs(x); return 1.f / (d * d); }
    default: return 1.f;
  }
}
inline int ActKind(const std::string& t, const std::string& who) {
  static const char* names[] = {"relu", "sigmoid", "tanh", "softrelu", "softsign"};
  for (int i = 0; i < 5; ++i) if (t == names[i]) return i;
  throw std::runtime_error(who + ": activation " + t + " is not supported");
}

// index helper for broadcasting: maps a flat index of `out` to the flat index of an operand with (possibly) size-1 axes
struct Bcast {
  std::vector<int64_t> out_dims, stride;
  Bcast(const Shape& out, const Shape& in) {
    out_dims = out; stride.assign(out.size(), 0);
    int64_t st = 1;
    for (int i = static_cast<int>(in.size()) - 1, o = static_cast<int>(out.size()) - 1; o >= 0; --i, --o) {
      const int64_t d = i >= 0 ? in[i] : 1;
      stride[o] = d == 1 ? 0 : st;
      st *= d;
    }
  }
  int64_t At(int64_t flat) const {
    int64_t off = 0;
    for (int o = static_cast<int>(out_dims.size()) - 1; o >= 0; --o) { off += (flat % out_dims[o]) * stride[o]; flat /= out_dims[o]; }
    return off;
  }
};

// ------------------------------------------------------------------------------------------------ executor
struct Tensor {                         // a bound array: external float storage + shape
  float* data = nullptr;
  Shape shape;
};

class Executor {
 public:
  // args / grads / aux in ListArguments / ListAuxiliaryStates order; grads[i].data may be null when reqs[i] == kNullOp
  Executor(const Symbol& sym, const std::vector<Tensor>& args, const std::vector<Tensor>& grads, const std::vector<int>& reqs, const std::vector<Tensor>& aux)
      : sym_(sym) {
    order_ = graph::Topo(sym_);
    const auto aux_nodes = graph::AuxNodes(order_);
    std::map<std::string, Shape> known;
    size_t ai = 0, xi = 0;
    for (Node* n : order_) {
      index_[n] = static_cast<int>(slots_.size());
      slots_.emplace_back();
      Slot& s = slots_.back();
      s.node = n;
      if (n->op != "null") continue;
      if (aux_nodes.count(n)) {
        if (xi >= aux.size()) throw std::runtime_error("Bind: " + std::to_string(aux.size()) + " auxiliary states given, the symbol has more (missing " + n->name + ")");
        s.ext = aux[xi].data; s.shape = aux[xi].shape; s.is_aux = true; ++xi;
      } else {
        if (ai >= args.size()) throw std::runtime_error("Bind: " + std::to_string(args.size()) + " arguments given, the symbol has more (missing " + n->name + ")");
        s.ext = args[ai].data; s.shape = args[ai].shape;
        s.req = ai < reqs.size() ? reqs[ai] : kNullOp;
        if (s.req != kNullOp) {
          if (ai >= grads.size() || grads[ai].data == nullptr) throw std::runtime_error("Bind: argument " + n->name + " has grad_req != null but no gradient array");
          if (grads[ai].shape != s.shape) throw std::runtime_error("Bind: gradient of " + n->name + " has shape " + ShapeStr(grads[ai].shape) + ", the argument " + ShapeStr(s.shape));
          s.ext_grad = grads[ai].data;
        }
        arg_slots_.push_back(index_[n]);
        ++ai;
      }
      if (s.ext == nullptr) throw std::runtime_error("Bind: null array for " + n->name);
      known[n->name] = s.shape;
    }
    if (ai != args.size()) throw std::runtime_error("Bind: " + std::to_string(args.size()) + " arguments given, the symbol takes " + std::to_string(ai));
    if (xi != aux.size()) throw std::runtime_error("Bind: " + std::to_string(aux.size()) + " auxiliary states given, the symbol takes " + std::to_string(xi));
    const graph::ShapeResult sr = graph::InferShapes(sym_, known, false);
    for (auto& s : slots_) {
      s.shape = sr.shape.at(s.node);
      if (s.node->op != "null") { s.own.assign(static_cast<size_t>(Numel(s.shape)), 0.f); }
      for (auto& e : s.node->inputs) s.in.push_back(index_.at(e.node.get()));
    }
    // gradient flow: a node needs a gradient when any input does; BlockGrad cuts it
    for (auto& s : slots_) {
      if (s.node->op == "null") { s.need_grad = s.req != kNullOp; continue; }
      if (s.node->op == "BlockGrad") continue;
      for (int i : s.in) if (slots_[i].need_grad) s.need_grad = true;
    }
    for (auto& h : sym_.outputs) heads_.push_back(index_.at(h.node.get()));
    rng_.seed(GlobalSeed().fetch_add(1) * 2654435761u + 12345u);
  }

  static std::atomic<uint32_t>& GlobalSeed() { static std::atomic<uint32_t> s{0}; return s; }

  size_t NumOutputs() const { return heads_.size(); }
  const Shape& OutputShape(size_t i) const { return slots_[heads_.at(i)].shape; }
  const float* OutputData(size_t i) const { const Slot& s = slots_[heads_.at(i)]; return s.node->op == "null" ? s.ext : s.own.data(); }

  void Forward(bool is_train) {
    is_train_ = is_train;
    for (auto& s : slots_) if (s.node->op != "null") Run(s);
    forwarded_ = true;
  }

  // head_grads[i] may be null: loss heads (SoftmaxOutput, *RegressionOutput, MakeLoss) ignore it, other heads get ones (MXExecutorBackward with
  // no head gradient on a non-loss head is an error in the reference; autograd's default of ones is the useful convention for a C front end)
  void Backward(const std::vector<const float*>& head_grads) {
    if (!forwarded_) throw std::runtime_error("Backward: call Forward(is_train=1) first");
    if (!head_grads.empty() && head_grads.size() != heads_.size()) throw std::runtime_error("Backward: " + std::to_string(head_grads.size()) + " head gradients for " + std::to_string(heads_.size()) + " outputs");
    for (auto& s : slots_) if (s.need_grad) s.grad.assign(static_cast<size_t>(Numel(s.shape)), 0.f); else s.grad.clear();
    for (size_t i = 0; i < heads_.size(); ++i) {
      Slot& s = slots_[heads_[i]];
      if (!s.need_grad) continue;
      const float* g = head_grads.empty() ? nullptr : head_grads[i];
      if (g) for (size_t k = 0; k < s.grad.size(); ++k) s.grad[k] += g[k];
      else for (auto& v : s.grad) v += 1.f;
    }
    for (size_t k = slots_.size(); k-- > 0;) {
      Slot& s = slots_[k];
      if (s.node->op == "null" || !s.need_grad) continue;
      Grad(s);
      std::vector<float>().swap(s.grad);                   // activations' gradients are dead once propagated
    }
    for (int i : arg_slots_) {
      Slot& s = slots_[i];
      if (s.req == kNullOp) continue;
      if (s.req == kAddTo) for (size_t k = 0; k < s.grad.size(); ++k) s.ext_grad[k] += s.grad[k];
      else memcpy(s.ext_grad, s.grad.data(), s.grad.size() * sizeof(float));
    }
  }

  std::string Print() const {
    std::string o;
    int64_t act = 0;
    for (auto& s : slots_) {
      if (s.node->op == "null") { o += "Variable:" + s.node->name + " " + ShapeStr(s.shape) + (s.is_aux ? " aux" : s.req != kNullOp ? " grad" : "") + "\n"; continue; }
      o += "Op:" + s.node->op + ", Name=" + s.node->name + " -> " + ShapeStr(s.shape) + "\n";
      for (int i : s.in) o += "  arg: " + slots_[i].node->name + "\n";
      act += Numel(s.shape);
    }
    o += "Total " + std::to_string(act * 4 / 1024) + " KB allocated for activations\n";
    return o;
  }

 private:
  struct Slot {
    Node* node = nullptr;
    Shape shape;
    std::vector<int> in;
    float* ext = nullptr;              // variables: the bound array
    float* ext_grad = nullptr;
    int req = kNullOp;
    bool is_aux = false, need_grad = false;
    std::vector<float> own, grad;
    std::vector<int32_t> idx;          // Pooling(max): winning input offset per output
    std::vector<float> saved;          // Dropout mask / BatchNorm batch mean + inverse std
  };
  Symbol sym_;
  std::vector<Node*> order_;
  std::unordered_map<Node*, int> index_;
  std::vector<Slot> slots_;
  std::vector<int> arg_slots_, heads_;
  bool is_train_ = false, forwarded_ = false;
  std::mt19937 rng_;

  const float* Val(int i) const { const Slot& s = slots_[i]; return s.node->op == "null" ? s.ext : s.own.data(); }
  float* AuxPtr(int i) { return slots_[i].ext; }
  static Win WinOf(const graph::detail::Win& w) { return Win{w.kh, w.kw, w.sh, w.sw, w.ph, w.pw, w.dh, w.dw}; }

  // softmax over the middle axis of (outer, c, inner)
  static void SoftmaxFwd(const float* x, float* y, int64_t outer, int64_t c, int64_t inner, bool log) {
    ParallelFor(outer * inner, static_cast<double>(outer) * inner * c * 8, [&](int64_t lo, int64_t hi) {
      for (int64_t t = lo; t < hi; ++t) {
        const int64_t o = t / inner, i = t % inner;
        const float* xs = x + o * c * inner + i; float* ys = y + o * c * inner + i;
        float m = xs[0];
        for (int64_t k = 1; k < c; ++k) m = std::max(m, xs[k * inner]);
        float z = 0;
        for (int64_t k = 0; k < c; ++k) z += std::exp(xs[k * inner] - m);
        const float lz = std::log(z);
        for (int64_t k = 0; k < c; ++k) ys[k * inner] = log ? xs[k * inner] - m - lz : std::exp(xs[k * inner] - m) / z;
      }
    });
  }
  static void SplitAxis(const Shape& s, int64_t ax, int64_t* outer, int64_t* c, int64_t* inner) {
    *outer = 1; *inner = 1; *c = s[ax];
    for (int64_t i = 0; i < ax; ++i) *outer *= s[i];
    for (size_t i = ax + 1; i < s.size(); ++i) *inner *= s[i];
  }

  // ---- forward
  void Run(Slot& s) {
    const Node& n = *s.node;
    const std::string& op = n.op;
    AttrView a(n.attrs);
    float* y = s.own.data();
    const int64_t ny = Numel(s.shape);
    const float* x = Val(s.in[0]);
    const Shape& xs = slots_[s.in[0]].shape;
    const int64_t nx = Numel(xs);
    if (op == "FullyConnected") {
      const int64_t h = s.shape.back(), k = slots_[s.in[1]].shape[1], m = nx / k;
      const float* b = s.in.size() > 2 ? Val(s.in[2]) : nullptr;
      predict::GemmNT(x, Val(s.in[1]), y, m, h, k, b);
    } else if (op == "Convolution") {
      const Win w = WinOf(graph::detail::Window(n, false, xs));
      const int64_t N = xs[0], C = xs[1], H = xs[2], W = xs[3], F = s.shape[1], OH = s.shape[2], OW = s.shape[3], G = a.Int("num_group", 1);
      const int64_t Cg = C / G, Fg = F / G, K = Cg * w.kh * w.kw, P = OH * OW;
      const float* wt = Val(s.in[1]); const float* b = s.in.size() > 2 ? Val(s.in[2]) : nullptr;
      ParallelFor(N, static_cast<double>(N) * F * K * P, [&](int64_t lo, int64_t hi) {
        std::vector<float> col(static_cast<size_t>(K * P));
        for (int64_t i = lo; i < hi; ++i) for (int64_t g = 0; g < G; ++g) {
          Im2Col(x + (i * C + g * Cg) * H * W, Cg, H, W, w, OH, OW, col.data());
          float* out = y + (i * F + g * Fg) * P;
          GemmSerial(false, false, Fg, P, K, wt + g * Fg * K, col.data(), out, false);
          if (b) for (int64_t f = 0; f < Fg; ++f) { const float bv = b[g * Fg + f]; float* o = out + f * P; for (int64_t p = 0; p < P; ++p) o[p] += bv; }
        }
      });
    } else if (op == "Pooling") {
      const Win w = WinOf(graph::detail::Window(n, true, xs));
      const std::string t = a.Str("pool_type", "max");
      const int type = t == "max" ? 0 : t == "avg" ? 1 : t == "sum" ? 2 : -1;
      if (type < 0) throw std::runtime_error(n.name + ": pool_type " + t + " is not supported");
      const bool count_pad = a.Bool("count_include_pad", true);
      const int64_t NC = xs[0] * xs[1], H = xs[2], W = xs[3], OH = s.shape[2], OW = s.shape[3];
      if (type == 0) s.idx.assign(static_cast<size_t>(ny), -1);
      ParallelFor(NC, static_cast<double>(ny) * w.kh * w.kw * 4, [&](int64_t lo, int64_t hi) {
        for (int64_t c = lo; c < hi; ++c) for (int64_t oy = 0; oy < OH; ++oy) for (int64_t ox = 0; ox < OW; ++ox) {
          const int64_t y0 = oy * w.sh - w.ph, x0 = ox * w.sw - w.pw;
          const int64_t ya = std::max<int64_t>(y0, 0), yb = std::min(y0 + w.kh, H), xa = std::max<int64_t>(x0, 0), xb = std::min(x0 + w.kw, W);
          const float* src = x + c * H * W;
          const int64_t o = (c * OH + oy) * OW + ox;
          if (type == 0) {
            float best = -std::numeric_limits<float>::infinity(); int32_t bi = -1;
            for (int64_t iy = ya; iy < yb; ++iy) for (int64_t ix = xa; ix < xb; ++ix) if (src[iy * W + ix] > best) { best = src[iy * W + ix]; bi = static_cast<int32_t>(iy * W + ix); }
            y[o] = bi < 0 ? 0.f : best; s.idx[o] = bi;
          } else {
            float acc = 0;
            for (int64_t iy = ya; iy < yb; ++iy) for (int64_t ix = xa; ix < xb; ++ix) acc += src[iy * W + ix];
            if (type == 1) {
              const int64_t full = (std::min(y0 + w.kh, H + w.ph) - y0) * (std::min(x0 + w.kw, W + w.pw) - x0);
              acc /= static_cast<float>(count_pad ? full : std::max<int64_t>((yb - ya) * (xb - xa), 1));
            }
            y[o] = acc;
          }
        }
      });
    } else if (op == "Activation") {
      const int k = ActKind(a.Str("act_type", "relu"), n.name);
      for (int64_t i = 0; i < ny; ++i) y[i] = ActF(k, x[i]);
    } else if (op == "LeakyReLU") {
      const std::string t = a.Str("act_type", "leaky");
      const float slope = static_cast<float>(a.Float("slope", 0.25));
      if (t == "leaky") for (int64_t i = 0; i < ny; ++i) y[i] = x[i] > 0 ? x[i] : slope * x[i];
      else if (t == "elu") for (int64_t i = 0; i < ny; ++i) y[i] = x[i] > 0 ? x[i] : slope * (std::exp(x[i]) - 1.f);
      else throw std::runtime_error(n.name + ": LeakyReLU act_type " + t + " is not supported");
    } else if (op == "BatchNorm") {
      const int64_t ax = graph::detail::AxisOf(a.Int("axis", 1), xs.size(), n.name);
      int64_t outer, C, inner; SplitAxis(xs, ax, &outer, &C, &inner);
      const float eps = static_cast<float>(a.Float("eps", 1e-3)), mom = static_cast<float>(a.Float("momentum", 0.9));
      const bool fix_gamma = a.Bool("fix_gamma", true), global = a.Bool("use_global_stats", false) || !is_train_;
      const float* gamma = Val(s.in[1]); const float* beta = Val(s.in[2]);
      float* mm = AuxPtr(s.in[3]); float* mv = AuxPtr(s.in[4]);
      s.saved.assign(static_cast<size_t>(2 * C), 0.f);
      const int64_t cnt = outer * inner;
      ParallelFor(C, static_cast<double>(nx) * 6, [&](int64_t lo, int64_t hi) {
        for (int64_t c = lo; c < hi; ++c) {
          float mean, var;
          if (global) { mean = mm[c]; var = mv[c]; }
          else {
            double sm = 0, sq = 0;
            for (int64_t o = 0; o < outer; ++o) { const float* p = x + (o * C + c) * inner; for (int64_t i = 0; i < inner; ++i) sm += p[i]; }
            mean = static_cast<float>(sm / cnt);
            for (int64_t o = 0; o < outer; ++o) { const float* p = x + (o * C + c) * inner; for (int64_t i = 0; i < inner; ++i) { const double d = p[i] - mean; sq += d * d; } }
            var = static_cast<float>(sq / cnt);
            mm[c] = mm[c] * mom + mean * (1.f - mom);
            mv[c] = mv[c] * mom + var * (1.f - mom);
          }
          const float inv = 1.f / std::sqrt(var + eps), g = fix_gamma ? 1.f : gamma[c];
          s.saved[c] = mean; s.saved[C + c] = inv;
          for (int64_t o = 0; o < outer; ++o) {
            const float* p = x + (o * C + c) * inner; float* q = y + (o * C + c) * inner;
            for (int64_t i = 0; i < inner; ++i) q[i] = (p[i] - mean) * inv * g + beta[c];
          }
        }
      });
    } else if (op == "Dropout") {
      const float p = static_cast<float>(a.Float("p", 0.5));
      if (!is_train_ || p <= 0.f) { s.saved.clear(); memcpy(y, x, ny * sizeof(float)); }
      else {
        if (p >= 1.f) throw std::runtime_error(n.name + ": drop probability must be < 1");
        s.saved.resize(static_cast<size_t>(ny));
        std::bernoulli_distribution keep(1.0 - p);
        const float scale = 1.f / (1.f - p);
        for (int64_t i = 0; i < ny; ++i) { s.saved[i] = keep(rng_) ? scale : 0.f; y[i] = x[i] * s.saved[i]; }
      }
    } else if (op == "Flatten" || op == "Reshape" || op == "expand_dims" || op == "identity" || op == "BlockGrad" || op == "MakeLoss" ||
               op == "LinearRegressionOutput" || op == "MAERegressionOutput") {
      memcpy(y, x, ny * sizeof(float));
    } else if (op == "LogisticRegressionOutput") {
      for (int64_t i = 0; i < ny; ++i) y[i] = ActF(1, x[i]);
    } else if (op == "transpose") {
      auto axes = a.Tuple("axes", {});
      const size_t r = xs.size();
      if (axes.empty()) for (size_t i = 0; i < r; ++i) axes.push_back(static_cast<int64_t>(r - 1 - i));
      std::vector<int64_t> xstride(r, 1);
      for (int i = static_cast<int>(r) - 2; i >= 0; --i) xstride[i] = xstride[i + 1] * xs[i + 1];
      for (int64_t f = 0; f < ny; ++f) {
        int64_t rem = f, off = 0;
        for (int i = static_cast<int>(r) - 1; i >= 0; --i) { off += (rem % s.shape[i]) * xstride[graph::detail::AxisOf(axes[i], r, n.name)]; rem /= s.shape[i]; }
        y[f] = x[off];
      }
    } else if (op == "Concat") {
      const int64_t ax = graph::detail::AxisOf(a.Int("dim", 1), s.shape.size(), n.name);
      int64_t outer, C, inner; SplitAxis(s.shape, ax, &outer, &C, &inner);
      int64_t at = 0;
      for (int i : s.in) {
        const int64_t ci = slots_[i].shape[ax]; const float* src = Val(i);
        for (int64_t o = 0; o < outer; ++o) memcpy(y + (o * C + at) * inner, src + o * ci * inner, ci * inner * sizeof(float));
        at += ci;
      }
    } else if (op == "add_n") {
      memcpy(y, x, ny * sizeof(float));
      for (size_t k = 1; k < s.in.size(); ++k) { const float* v = Val(s.in[k]); for (int64_t i = 0; i < ny; ++i) y[i] += v[i]; }
    } else if (op == "Embedding") {
      const float* w = Val(s.in[1]);
      const int64_t V = slots_[s.in[1]].shape[0], D = slots_[s.in[1]].shape[1];
      for (int64_t i = 0; i < nx; ++i) {
        const int64_t r = std::min<int64_t>(std::max<int64_t>(static_cast<int64_t>(x[i]), 0), V - 1);
        memcpy(y + i * D, w + r * D, D * sizeof(float));
      }
    } else if (op == "SoftmaxOutput" || op == "SoftmaxActivation") {
      int64_t outer, C, inner; SplitAxis(xs, 1, &outer, &C, &inner);
      if (op == "SoftmaxOutput" && !a.Bool("multi_output", false) && xs.size() > 2) { C = nx / xs[0]; inner = 1; outer = xs[0]; }
      SoftmaxFwd(x, y, outer, C, inner, false);
    } else if (op == "softmax" || op == "log_softmax") {
      int64_t outer, C, inner; SplitAxis(xs, graph::detail::AxisOf(a.Int("axis", -1), xs.size(), n.name), &outer, &C, &inner);
      SoftmaxFwd(x, y, outer, C, inner, op == "log_softmax");
    } else if (op == "clip") {
      const float lo = static_cast<float>(a.Float("a_min", -std::numeric_limits<float>::infinity())), hi = static_cast<float>(a.Float("a_max", std::numeric_limits<float>::infinity()));
      for (int64_t i = 0; i < ny; ++i) y[i] = std::min(std::max(x[i], lo), hi);
    } else if (op == "sum" || op == "mean") {
      ReduceFwd(s, x, xs, y, op == "mean");
    } else if (op == "dot") {
      const bool ta = a.Bool("transpose_a", false), tb = a.Bool("transpose_b", false);
      const int64_t M = s.shape[0], N = s.shape[1], K = ta ? xs[0] : xs[1];
      Gemm(ta, tb, M, N, K, x, Val(s.in[1]), y, false);
    } else if (s.in.size() == 2) {
      const int kind = BinaryKind(op);
      const float* r = Val(s.in[1]);
      const Shape& rs = slots_[s.in[1]].shape;
      if (xs == s.shape && rs == s.shape) for (int64_t i = 0; i < ny; ++i) y[i] = Bin(kind, x[i], r[i]);
      else { const Bcast bl(s.shape, xs), br(s.shape, rs); for (int64_t i = 0; i < ny; ++i) y[i] = Bin(kind, x[bl.At(i)], r[br.At(i)]); }
    } else if (op[0] == '_') {
      const float c = static_cast<float>(a.Float("scalar", 0));
      const int k = ScalarKind(op);
      for (int64_t i = 0; i < ny; ++i) y[i] = Sc(k, x[i], c);
    } else {
      const int k = UnaryKind(op);
      if (k < 0) throw std::runtime_error("operator " + op + " has no host kernel");
      for (int64_t i = 0; i < ny; ++i) y[i] = Un(k, x[i]);
    }
  }

  static int BinaryKind(const std::string& op) {
    static const char* names[] = {"add", "sub", "mul", "div", "maximum", "minimum"};
    for (int i = 0; i < 6; ++i) if (op.find(names[i]) != std::string::npos) return i;
    throw std::runtime_error("operator " + op + " has no host kernel");
  }
  static float Bin(int k, float l, float r) { switch (k) { case 0: return l + r; case 1: return l - r; case 2: return l * r; case 3: return l / r; case 4: return std::max(l, r); default: return std::min(l, r); } }
  static int ScalarKind(const std::string& op) {
    static const char* names[] = {"_plus_scalar", "_minus_scalar", "_rminus_scalar", "_mul_scalar", "_div_scalar", "_rdiv_scalar", "_power_scalar"};
    for (int i = 0; i < 7; ++i) if (op == names[i]) return i;
    throw std::runtime_error("operator " + op + " has no host kernel");
  }
  static float Sc(int k, float x, float c) { switch (k) { case 0: return x + c; case 1: return x - c; case 2: return c - x; case 3: return x * c; case 4: return x / c; case 5: return c / x; default: return std::pow(x, c); } }
  static int UnaryKind(const std::string& op) {
    static const char* names[] = {"relu", "sigmoid", "tanh", "exp", "log", "sqrt", "abs", "negative", "square", "softsign"};
    for (int i = 0; i < 10; ++i) if (op == names[i]) return i;
    return -1;
  }
  static float Un(int k, float v) {
    switch (k) {
      case 0: return v > 0 ? v : 0; case 1: return 1.f / (1.f + std::exp(-v)); case 2: return std::tanh(v); case 3: return std::exp(v); case 4: return std::log(v);
      case 5: return std::sqrt(v); case 6: return std::fabs(v); case 7: return -v; case 8: return v * v; default: return v / (1.f + std::fabs(v));
    }
  }
  // d out / d in for unary kinds, from input x and output y
  static float UnG(int k, float x, float y) {
    switch (k) {
      case 0: return x > 0 ? 1.f : 0.f; case 1: return y * (1.f - y); case 2: return 1.f - y * y; case 3: return y; case 4: return 1.f / x;
      case 5: return 0.5f / y; case 6: return x > 0 ? 1.f : x < 0 ? -1.f : 0.f; case 7: return -1.f; case 8: return 2.f * x;
      default: { const float d = 1.f + std::fabs(x); return 1.f / (d * d); }
    }
  }

  std::vector<char> ReducedAxes(const Slot& s, const Shape& xs) const {
    auto axes = AttrView(s.node->attrs).Tuple("axis", {});
    std::vector<char> red(xs.size(), axes.empty());
    for (auto ax : axes) red[graph::detail::AxisOf(ax, xs.size(), s.node->name)] = 1;
    return red;
  }
  // flat index of the reduced output that input element `f` contributes to
  static int64_t ReducedIndex(int64_t f, const Shape& xs, const std::vector<char>& red) {
    int64_t off = 0, st = 1;
    for (int i = static_cast<int>(xs.size()) - 1; i >= 0; --i) { const int64_t c = f % xs[i]; f /= xs[i]; if (!red[i]) { off += c * st; st *= xs[i]; } }
    return off;
  }
  void ReduceFwd(Slot& s, const float* x, const Shape& xs, float* y, bool mean) const {
    const auto red = ReducedAxes(s, xs);
    const int64_t nx = Numel(xs), ny = Numel(s.shape);
    std::vector<double> acc(static_cast<size_t>(ny), 0.0);
    for (int64_t f = 0; f < nx; ++f) acc[ReducedIndex(f, xs, red)] += x[f];
    const double div = mean ? static_cast<double>(nx / ny) : 1.0;
    for (int64_t i = 0; i < ny; ++i) y[i] = static_cast<float>(acc[i] / div);
  }

  // ---- backward: s.grad holds d loss / d output; adds into the inputs' grad buffers (only where need_grad)
  float* GradOf(int i) { Slot& s = slots_[i]; return s.need_grad ? s.grad.data() : nullptr; }

  void Grad(Slot& s) {
    const Node& n = *s.node;
    const std::string& op = n.op;
    AttrView a(n.attrs);
    const float* dy = s.grad.data();
    const float* y = s.own.data();
    const int64_t ny = Numel(s.shape);
    const float* x = Val(s.in[0]);
    const Shape& xs = slots_[s.in[0]].shape;
    const int64_t nx = Numel(xs);
    float* dx = GradOf(s.in[0]);
    if (op == "FullyConnected") {
      const int64_t h = s.shape.back(), k = slots_[s.in[1]].shape[1], m = nx / k;
      if (dx) Gemm(false, false, m, k, h, dy, Val(s.in[1]), dx, true);                              // dX += dY . W
      if (float* dw = GradOf(s.in[1])) Gemm(true, false, h, k, m, dy, x, dw, true);                  // dW += dY^T . X
      if (s.in.size() > 2) if (float* db = GradOf(s.in[2])) for (int64_t i = 0; i < m; ++i) for (int64_t j = 0; j < h; ++j) db[j] += dy[i * h + j];
    } else if (op == "Convolution") {
      const Win w = WinOf(graph::detail::Window(n, false, xs));
      const int64_t N = xs[0], C = xs[1], H = xs[2], W = xs[3], F = s.shape[1], OH = s.shape[2], OW = s.shape[3], G = a.Int("num_group", 1);
      const int64_t Cg = C / G, Fg = F / G, K = Cg * w.kh * w.kw, P = OH * OW;
      const float* wt = Val(s.in[1]);
      float* dw = GradOf(s.in[1]);
      float* db = s.in.size() > 2 ? GradOf(s.in[2]) : nullptr;
      std::mutex mu;
      ParallelFor(N, static_cast<double>(N) * F * K * P * 2, [&](int64_t lo, int64_t hi) {
        std::vector<float> col(static_cast<size_t>(K * P)), dcol(dx ? static_cast<size_t>(K * P) : 0);
        std::vector<float> dw_local(dw ? static_cast<size_t>(F * K) : 0, 0.f), db_local(db ? static_cast<size_t>(F) : 0, 0.f);
        for (int64_t i = lo; i < hi; ++i) for (int64_t g = 0; g < G; ++g) {
          const float* dout = dy + (i * F + g * Fg) * P;
          if (dw) {
            Im2Col(x + (i * C + g * Cg) * H * W, Cg, H, W, w, OH, OW, col.data());
            GemmSerial(false, true, Fg, K, P, dout, col.data(), dw_local.data() + g * Fg * K, true);    // dW += dOut . col^T
          }
          if (db) for (int64_t f = 0; f < Fg; ++f) { float sm = 0; for (int64_t p = 0; p < P; ++p) sm += dout[f * P + p]; db_local[g * Fg + f] += sm; }
          if (dx) {
            GemmSerial(true, false, K, P, Fg, wt + g * Fg * K, dout, dcol.data(), false);               // dcol = W^T . dOut
            Col2Im(dcol.data(), Cg, H, W, w, OH, OW, dx + (i * C + g * Cg) * H * W);
          }
        }
        std::lock_guard<std::mutex> lk(mu);
        if (dw) for (size_t k = 0; k < dw_local.size(); ++k) dw[k] += dw_local[k];
        if (db) for (size_t k = 0; k < db_local.size(); ++k) db[k] += db_local[k];
      });
    } else if (op == "Pooling") {
      if (!dx) return;
      const Win w = WinOf(graph::detail::Window(n, true, xs));
      const std::string t = a.Str("pool_type", "max");
      const bool count_pad = a.Bool("count_include_pad", true);
      const int64_t NC = xs[0] * xs[1], H = xs[2], W = xs[3], OH = s.shape[2], OW = s.shape[3];
      for (int64_t c = 0; c < NC; ++c) for (int64_t oy = 0; oy < OH; ++oy) for (int64_t ox = 0; ox < OW; ++ox) {
        const int64_t o = (c * OH + oy) * OW + ox;
        float* dst = dx + c * H * W;
        if (t == "max") { if (s.idx[o] >= 0) dst[s.idx[o]] += dy[o]; continue; }
        const int64_t y0 = oy * w.sh - w.ph, x0 = ox * w.sw - w.pw;
        const int64_t ya = std::max<int64_t>(y0, 0), yb = std::min(y0 + w.kh, H), xa = std::max<int64_t>(x0, 0), xb = std::min(x0 + w.kw, W);
        float g = dy[o];
        if (t == "avg") {
          const int64_t full = (std::min(y0 + w.kh, H + w.ph) - y0) * (std::min(x0 + w.kw, W + w.pw) - x0);
          g /= static_cast<float>(count_pad ? full : std::max<int64_t>((yb - ya) * (xb - xa), 1));
        }
        for (int64_t iy = ya; iy < yb; ++iy) for (int64_t ix = xa; ix < xb; ++ix) dst[iy * W + ix] += g;
      }
    } else if (op == "Activation") {
      if (!dx) return;
      const int k = ActKind(a.Str("act_type", "relu"), n.name);
      for (int64_t i = 0; i < ny; ++i) dx[i] += dy[i] * ActG(k, x[i], y[i]);
    } else if (op == "LeakyReLU") {
      if (!dx) return;
      const float slope = static_cast<float>(a.Float("slope", 0.25));
      const bool elu = a.Str("act_type", "leaky") == "elu";
      for (int64_t i = 0; i < ny; ++i) dx[i] += dy[i] * (x[i] > 0 ? 1.f : elu ? y[i] + slope : slope);
    } else if (op == "BatchNorm") {
      const int64_t ax = graph::detail::AxisOf(a.Int("axis", 1), xs.size(), n.name);
      int64_t outer, C, inner; SplitAxis(xs, ax, &outer, &C, &inner);
      const bool fix_gamma = a.Bool("fix_gamma", true), global = a.Bool("use_global_stats", false) || !is_train_;
      const float* gamma = Val(s.in[1]);
      float* dg = GradOf(s.in[1]); float* dbeta = GradOf(s.in[2]);
      const int64_t cnt = outer * inner;
      for (int64_t c = 0; c < C; ++c) {
        const float mean = s.saved[c], inv = s.saved[C + c], g = fix_gamma ? 1.f : gamma[c];
        double sdy = 0, sdyx = 0;
        for (int64_t o = 0; o < outer; ++o) {
          const float* p = x + (o * C + c) * inner; const float* q = dy + (o * C + c) * inner;
          for (int64_t i = 0; i < inner; ++i) { sdy += q[i]; sdyx += q[i] * (p[i] - mean) * inv; }
        }
        if (dg && !fix_gamma) dg[c] += static_cast<float>(sdyx);
        if (dbeta) dbeta[c] += static_cast<float>(sdy);
        if (!dx) continue;
        const float msdy = static_cast<float>(sdy / cnt), msdyx = static_cast<float>(sdyx / cnt);
        for (int64_t o = 0; o < outer; ++o) {
          const float* p = x + (o * C + c) * inner; const float* q = dy + (o * C + c) * inner; float* d = dx + (o * C + c) * inner;
          if (global) for (int64_t i = 0; i < inner; ++i) d[i] += q[i] * g * inv;
          else for (int64_t i = 0; i < inner; ++i) d[i] += g * inv * (q[i] - msdy - (p[i] - mean) * inv * msdyx);
        }
      }
    } else if (op == "Dropout") {
      if (!dx) return;
      if (s.saved.empty()) for (int64_t i = 0; i < ny; ++i) dx[i] += dy[i];
      else for (int64_t i = 0; i < ny; ++i) dx[i] += dy[i] * s.saved[i];
    } else if (op == "Flatten" || op == "Reshape" || op == "expand_dims" || op == "identity") {
      if (dx) for (int64_t i = 0; i < ny; ++i) dx[i] += dy[i];
    } else if (op == "BlockGrad") {
    } else if (op == "MakeLoss") {
      if (dx) { const float gs = static_cast<float>(a.Float("grad_scale", 1.0)); for (int64_t i = 0; i < ny; ++i) dx[i] += gs; }
    } else if (op == "transpose") {
      if (!dx) return;
      auto axes = a.Tuple("axes", {});
      const size_t r = xs.size();
      if (axes.empty()) for (size_t i = 0; i < r; ++i) axes.push_back(static_cast<int64_t>(r - 1 - i));
      std::vector<int64_t> xstride(r, 1);
      for (int i = static_cast<int>(r) - 2; i >= 0; --i) xstride[i] = xstride[i + 1] * xs[i + 1];
      for (int64_t f = 0; f < ny; ++f) {
        int64_t rem = f, off = 0;
        for (int i = static_cast<int>(r) - 1; i >= 0; --i) { off += (rem % s.shape[i]) * xstride[graph::detail::AxisOf(axes[i], r, n.name)]; rem /= s.shape[i]; }
        dx[off] += dy[f];
      }
    } else if (op == "Concat") {
      const int64_t ax = graph::detail::AxisOf(a.Int("dim", 1), s.shape.size(), n.name);
      int64_t outer, C, inner; SplitAxis(s.shape, ax, &outer, &C, &inner);
      int64_t at = 0;
      for (int i : s.in) {
        const int64_t ci = slots_[i].shape[ax];
        if (float* d = GradOf(i)) for (int64_t o = 0; o < outer; ++o) { const float* q = dy + (o * C + at) * inner; float* dd = d + o * ci * inner; for (int64_t k = 0; k < ci * inner; ++k) dd[k] += q[k]; }
        at += ci;
      }
    } else if (op == "add_n") {
      for (int i : s.in) if (float* d = GradOf(i)) for (int64_t k = 0; k < ny; ++k) d[k] += dy[k];
    } else if (op == "Embedding") {
      if (float* dw = GradOf(s.in[1])) {
        const int64_t V = slots_[s.in[1]].shape[0], D = slots_[s.in[1]].shape[1];
        for (int64_t i = 0; i < nx; ++i) {
          const int64_t r = std::min<int64_t>(std::max<int64_t>(static_cast<int64_t>(x[i]), 0), V - 1);
          for (int64_t d = 0; d < D; ++d) dw[r * D + d] += dy[i * D + d];
        }
      }
    } else if (op == "SoftmaxOutput") {
      if (!dx) return;
      const float* label = Val(s.in[1]);
      const Shape& ls = slots_[s.in[1]].shape;
      const float gs = static_cast<float>(a.Float("grad_scale", 1.0));
      if (ls == xs) { for (int64_t i = 0; i < ny; ++i) dx[i] += (y[i] - label[i]) * gs; return; }      // probability labels
      int64_t outer, C, inner; SplitAxis(xs, 1, &outer, &C, &inner);
      if (!a.Bool("multi_output", false) && xs.size() > 2) { C = nx / xs[0]; inner = 1; outer = xs[0]; }
      if (Numel(ls) != outer * inner) throw std::runtime_error(n.name + ": label shape " + ShapeStr(ls) + " does not match the prediction " + ShapeStr(xs));
      const bool use_ignore = a.Bool("use_ignore", false);
      const float ignore = static_cast<float>(a.Float("ignore_label", -1));
      const std::string norm = a.Str("normalization", "null");
      int64_t valid = 0;
      for (int64_t t = 0; t < outer * inner; ++t) if (!(use_ignore && label[t] == ignore)) ++valid;
      const float scale = gs / (norm == "batch" ? static_cast<float>(outer) : norm == "valid" ? static_cast<float>(std::max<int64_t>(valid, 1)) : 1.f);
      for (int64_t o = 0; o < outer; ++o) for (int64_t i = 0; i < inner; ++i) {
        const float l = label[o * inner + i];
        if (use_ignore && l == ignore) continue;
        const int64_t cls = static_cast<int64_t>(l);
        for (int64_t k = 0; k < C; ++k) dx[(o * C + k) * inner + i] += (y[(o * C + k) * inner + i] - (k == cls ? 1.f : 0.f)) * scale;
      }
    } else if (op == "LinearRegressionOutput" || op == "LogisticRegressionOutput" || op == "MAERegressionOutput") {
      if (!dx) return;
      const float* label = Val(s.in[1]);
      if (Numel(slots_[s.in[1]].shape) != ny) throw std::runtime_error(n.name + ": label size does not match the prediction");
      const float scale = static_cast<float>(a.Float("grad_scale", 1.0)) / static_cast<float>(std::max<int64_t>(ny / std::max<int64_t>(xs[0], 1), 1));
      if (op == "MAERegressionOutput") for (int64_t i = 0; i < ny; ++i) dx[i] += (y[i] > label[i] ? 1.f : y[i] < label[i] ? -1.f : 0.f) * scale;
      else for (int64_t i = 0; i < ny; ++i) dx[i] += (y[i] - label[i]) * scale;
    } else if (op == "SoftmaxActivation" || op == "softmax" || op == "log_softmax") {
      if (!dx) return;
      int64_t outer, C, inner;
      SplitAxis(xs, op == "SoftmaxActivation" ? 1 : graph::detail::AxisOf(a.Int("axis", -1), xs.size(), n.name), &outer, &C, &inner);
      const bool lg = op == "log_softmax";
      for (int64_t o = 0; o < outer; ++o) for (int64_t i = 0; i < inner; ++i) {
        const float* ys = y + o * C * inner + i; const float* gs = dy + o * C * inner + i; float* ds = dx + o * C * inner + i;
        float dot = 0;
        for (int64_t k = 0; k < C; ++k) dot += lg ? gs[k * inner] : gs[k * inner] * ys[k * inner];
        for (int64_t k = 0; k < C; ++k) ds[k * inner] += lg ? gs[k * inner] - std::exp(ys[k * inner]) * dot : ys[k * inner] * (gs[k * inner] - dot);
      }
    } else if (op == "clip") {
      if (!dx) return;
      const float lo = static_cast<float>(a.Float("a_min", -std::numeric_limits<float>::infinity())), hi = static_cast<float>(a.Float("a_max", std::numeric_limits<float>::infinity()));
      for (int64_t i = 0; i < ny; ++i) if (x[i] >= lo && x[i] <= hi) dx[i] += dy[i];
    } else if (op == "sum" || op == "mean") {
      if (!dx) return;
      const auto red = ReducedAxes(s, xs);
      const float div = op == "mean" ? static_cast<float>(nx / ny) : 1.f;
      for (int64_t f = 0; f < nx; ++f) dx[f] += dy[ReducedIndex(f, xs, red)] / div;
    } else if (op == "dot") {
      const bool ta = a.Bool("transpose_a", false), tb = a.Bool("transpose_b", false);
      const float* r = Val(s.in[1]);
      const Shape& rs = slots_[s.in[1]].shape;
      const int64_t M = s.shape[0], N = s.shape[1], K = ta ? xs[0] : xs[1];
      // Y = op(A) op(B):  d op(A) = dY op(B)^T,  d op(B) = op(A)^T dY;  a transposed operand receives the transpose of that
      if (dx) { if (!ta) Gemm(false, !tb, M, K, N, dy, r, dx, true); else Gemm(tb, true, K, M, N, r, dy, dx, true); }
      if (float* dr = GradOf(s.in[1])) { if (!tb) Gemm(!ta, false, K, N, M, x, dy, dr, true); else Gemm(true, ta, N, K, M, dy, x, dr, true); }
      (void)rs;
    } else if (s.in.size() == 2) {
      const int kind = BinaryKind(op);
      const float* r = Val(s.in[1]);
      const Shape& rs = slots_[s.in[1]].shape;
      float* dr = GradOf(s.in[1]);
      const Bcast bl(s.shape, xs), br(s.shape, rs);
      const bool same = xs == s.shape && rs == s.shape;
      for (int64_t i = 0; i < ny; ++i) {
        const int64_t li = same ? i : bl.At(i), ri = same ? i : br.At(i);
        const float l = x[li], rv = r[ri], g = dy[i];
        float gl, gr;
        switch (kind) {
          case 0: gl = g; gr = g; break;
          case 1: gl = g; gr = -g; break;
          case 2: gl = g * rv; gr = g * l; break;
          case 3: gl = g / rv; gr = -g * l / (rv * rv); break;
          case 4: gl = l >= rv ? g : 0.f; gr = l >= rv ? 0.f : g; break;
          default: gl = l <= rv ? g : 0.f; gr = l <= rv ? 0.f : g; break;
        }
        if (dx) dx[li] += gl;
        if (dr) dr[ri] += gr;
      }
    } else if (op[0] == '_') {
      if (!dx) return;
      const float c = static_cast<float>(a.Float("scalar", 0));
      const int k = ScalarKind(op);
      for (int64_t i = 0; i < ny; ++i) {
        float g;
        switch (k) { case 0: case 1: g = 1.f; break; case 2: g = -1.f; break; case 3: g = c; break; case 4: g = 1.f / c; break; case 5: g = -c / (x[i] * x[i]); break; default: g = c * std::pow(x[i], c - 1.f); }
        dx[i] += dy[i] * g;
      }
    } else {
      if (!dx) return;
      const int k = UnaryKind(op);
      for (int64_t i = 0; i < ny; ++i) dx[i] += dy[i] * UnG(k, x[i], y[i]);
    }
  }
};

}  // namespace exec
}  // namespace gxrt
