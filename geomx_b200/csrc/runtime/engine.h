// Dependency engine: versioned variables + read/write dependency tracking + per-device prioritised worker pools.
// Capability parity: include/mxnet/engine.h:95-314 (NewVariable / DeleteVariable / PushAsync(exec_ctx, const_vars, mutable_vars, prop,
// priority) / WaitForVar / WaitForAll), src/engine/threaded_engine.h:66-553 (per-variable pending queue of readers / writers, an op is ready
// when all its dependencies are resolved; an exception thrown by an op is remembered on the variables it writes and re-thrown at the next
// wait point), src/engine/threaded_engine_perdevice.cc:48-312 (one worker pool per device for compute, a separate one for copies, a CPU
// pool and a high-priority CPU pool) and src/engine/naive_engine.cc (MXNET_ENGINE_TYPE=NaiveEngine runs everything inline).
//
// Design: dependency state is one table under one mutex (ops are host-side closures — kvstore sends / receives, staging copies,
// checkpoint IO, Python callbacks, per-device launch sequences — so the table is never the bottleneck); execution is sharded into POOLS keyed
// by (device, kind).  A pool owns a priority queue and its threads and is created on first use: the CPU pool (device -1, normal), the
// priority pool (FnProperty::kPriority ops: kvstore traffic must not queue behind checkpoint writes), and per device a compute and a copy pool
// (MXNET_GPU_WORKER_NTHREADS / MXNET_GPU_COPY_NTHREADS).  On B200 the DEVICE-side order of work is the CUDA streams' business; what a
// per-device pool buys is that the host threads that issue launches / copies for GPU d never wait behind host work for GPU e.
// Duplicate vars in const / mutable sets are rejected like threaded_engine.h:432.
#pragma once
#ifndef GEOMX_NO_PYTHON   // the Python-free C library (lib/libgeomx_capi.so) compiles the runtime without the pybind11 bindings
#include <pybind11/functional.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>
#endif

#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <deque>
#include <exception>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <queue>
#include <set>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace gxrt {
#ifndef GEOMX_NO_PYTHON   // the Python-free C library (lib/libgeomx_capi.so) compiles the runtime without the pybind11 bindings
namespace py = pybind11;
#endif

enum class FnProperty : int { kNormal = 0, kCopy = 1, kPriority = 2 };

class Engine {
 public:
  using Fn = std::function<void()>;
  explicit Engine(int num_threads, bool naive) : naive_(naive), cpu_threads_(std::max(1, num_threads)) {
    const char* g = getenv("MXNET_GPU_WORKER_NTHREADS"); gpu_threads_ = g ? std::max(1, atoi(g)) : 2;
    const char* c = getenv("MXNET_GPU_COPY_NTHREADS"); copy_threads_ = c ? std::max(1, atoi(c)) : 1;
    const char* p = getenv("MXNET_CPU_PRIORITY_NTHREADS"); prio_threads_ = p ? std::max(1, atoi(p)) : 2;
  }
  ~Engine() {
    std::vector<std::shared_ptr<Pool>> pools;
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
      for (auto& kv : pools_) pools.push_back(kv.second);
    }
    for (auto& p : pools) { { std::lock_guard<std::mutex> lk(p->mu); p->stop = true; } p->cv.notify_all(); }
    for (auto& p : pools) for (auto& t : p->threads) t.join();
  }

  int NewVariable() { std::lock_guard<std::mutex> lk(mu_); vars_.emplace(next_var_, Var()); return next_var_++; }

  // The variable disappears once every op pushed so far that touches it has finished (engine.h DeleteVariable).
  void DeleteVariable(int v) {
    Push([this, v] { std::lock_guard<std::mutex> lk(mu_); doomed_.push_back(v); }, {}, {v}, 0, "DeleteVariable", -1, FnProperty::kNormal);
  }

  void Push(Fn fn, const std::vector<int>& const_vars, const std::vector<int>& mutable_vars, int priority, const std::string& name,
            int device = -1, FnProperty prop = FnProperty::kNormal) {
    {
      std::set<int> seen;
      for (int v : mutable_vars) if (!seen.insert(v).second) throw std::runtime_error("duplicate variable in mutable_vars");
      for (int v : const_vars) if (seen.count(v)) throw std::runtime_error("variable appears in both const_vars and mutable_vars");
    }
    if (naive_) {
      try { fn(); } catch (...) { std::lock_guard<std::mutex> lk(mu_); for (int v : mutable_vars) vars_.at(v).error = std::current_exception(); global_error_ = std::current_exception(); }
      return;
    }
    auto op = std::make_shared<Op>();
    op->fn = std::move(fn); op->priority = priority; op->name = name; op->reads = const_vars; op->writes = mutable_vars;
    std::unique_lock<std::mutex> lk(mu_);
    op->pool = PoolFor(device, prop);
    ++pending_;
    int wait = 0;
    for (int v : const_vars) { Var& var = vars_.at(v); if (var.writer_active || !var.queue.empty()) { var.queue.push_back({op, false}); ++wait; } else ++var.readers; }
    for (int v : mutable_vars) { Var& var = vars_.at(v); if (var.writer_active || var.readers > 0 || !var.queue.empty()) { var.queue.push_back({op, true}); ++wait; } else var.writer_active = true; }
    op->wait = wait;
    if (wait == 0) { op->seq = seq_++; lk.unlock(); Enqueue(op); }
  }

  // Blocks until everything pushed so far that WRITES `v` has run; re-throws the exception of a failed writer (once).
  void WaitForVar(int v) {
    std::mutex m; std::condition_variable c; bool done = false;
    Push([&] { std::lock_guard<std::mutex> lk(m); done = true; c.notify_all(); }, {v}, {}, 1 << 20, "WaitForVar", -1, FnProperty::kPriority);
    { std::unique_lock<std::mutex> lk(m); c.wait(lk, [&] { return done; }); }
    std::exception_ptr e;
    { std::lock_guard<std::mutex> lk(mu_); auto it = vars_.find(v); if (it != vars_.end()) { e = it->second.error; it->second.error = nullptr; } }
    if (e) std::rethrow_exception(e);
  }
  void WaitForAll() {
    std::exception_ptr e;
    {
      std::unique_lock<std::mutex> lk(mu_);
      done_cv_.wait(lk, [this] { return pending_ == 0; });
      e = global_error_; global_error_ = nullptr;
      for (auto& kv : vars_) kv.second.error = nullptr;
      for (int v : doomed_) vars_.erase(v);
      doomed_.clear();
    }
    if (e) std::rethrow_exception(e);
  }

  // {"cpu": n, "priority": n, "gpu0": n, "gpu0/copy": n, ...}: ops executed per pool (tests, profiler)
  std::map<std::string, long> Stats() {
    std::lock_guard<std::mutex> lk(mu_);
    std::map<std::string, long> out;
    for (auto& kv : pools_) out[kv.second->label] = kv.second->executed.load();
    return out;
  }
  int NumVariables() { std::lock_guard<std::mutex> lk(mu_); return static_cast<int>(vars_.size()); }

 private:
  struct Pool;
  struct Op { Fn fn; int priority = 0; std::string name; std::vector<int> reads, writes; int wait = 0; uint64_t seq = 0; std::shared_ptr<Pool> pool; };
  struct Pending { std::shared_ptr<Op> op; bool write; };
  struct Var { int readers = 0; bool writer_active = false; std::deque<Pending> queue; std::exception_ptr error; };
  struct Cmp { bool operator()(const std::shared_ptr<Op>& a, const std::shared_ptr<Op>& b) const { return a->priority != b->priority ? a->priority < b->priority : a->seq > b->seq; } };
  struct Pool {
    std::string label;
    std::mutex mu;
    std::condition_variable cv;
    std::priority_queue<std::shared_ptr<Op>, std::vector<std::shared_ptr<Op>>, Cmp> ready;
    std::vector<std::thread> threads;
    std::atomic<long> executed{0};
    bool stop = false;
  };

  // mu_ held
  std::shared_ptr<Pool> PoolFor(int device, FnProperty prop) {
    const int kind = device < 0 ? (prop == FnProperty::kPriority ? 2 : 0) : (prop == FnProperty::kCopy ? 1 : 0);
    const long key = (static_cast<long>(device < 0 ? -1 : device) << 2) | kind;
    auto it = pools_.find(key);
    if (it != pools_.end()) return it->second;
    auto p = std::make_shared<Pool>();
    int n;
    if (device < 0) { p->label = kind == 2 ? "priority" : "cpu"; n = kind == 2 ? prio_threads_ : cpu_threads_; }
    else { p->label = "gpu" + std::to_string(device) + (kind == 1 ? "/copy" : ""); n = kind == 1 ? copy_threads_ : gpu_threads_; }
    for (int i = 0; i < n; ++i) p->threads.emplace_back([this, p] { Work(p.get()); });
    pools_[key] = p;
    return p;
  }
  void Enqueue(const std::shared_ptr<Op>& op) {
    Pool* p = op->pool.get();
    { std::lock_guard<std::mutex> lk(p->mu); p->ready.push(op); }
    p->cv.notify_one();
  }
  void Work(Pool* pool) {
    while (true) {
      std::shared_ptr<Op> op;
      {
        std::unique_lock<std::mutex> lk(pool->mu);
        pool->cv.wait(lk, [pool] { return pool->stop || !pool->ready.empty(); });
        if (pool->ready.empty()) return;
        op = pool->ready.top(); pool->ready.pop();
      }
      std::exception_ptr err;
      try { op->fn(); } catch (...) { err = std::current_exception(); }
      ++pool->executed;
      Complete(op, err);
    }
  }
  void Grant(Var& var, std::vector<std::shared_ptr<Op>>* runnable) {
    while (!var.queue.empty()) {
      Pending p = var.queue.front();
      if (p.write) {
        if (var.readers > 0 || var.writer_active) break;
        var.writer_active = true; var.queue.pop_front();
        if (--p.op->wait == 0) runnable->push_back(p.op);
        break;
      }
      if (var.writer_active) break;
      ++var.readers; var.queue.pop_front();
      if (--p.op->wait == 0) runnable->push_back(p.op);
    }
  }
  void Complete(const std::shared_ptr<Op>& op, std::exception_ptr err) {
    std::vector<std::shared_ptr<Op>> runnable;
    {
      std::lock_guard<std::mutex> lk(mu_);
      if (err) { for (int v : op->writes) { auto it = vars_.find(v); if (it != vars_.end()) it->second.error = err; } global_error_ = err; }
      for (int v : op->reads) { Var& var = vars_.at(v); --var.readers; Grant(var, &runnable); }
      for (int v : op->writes) { Var& var = vars_.at(v); var.writer_active = false; Grant(var, &runnable); }
      for (auto& r : runnable) r->seq = seq_++;
      if (--pending_ == 0) done_cv_.notify_all();
    }
    for (auto& r : runnable) Enqueue(r);
  }

  bool naive_;
  int cpu_threads_, gpu_threads_ = 2, copy_threads_ = 1, prio_threads_ = 2;
  std::mutex mu_;
  std::condition_variable done_cv_;
  std::map<long, std::shared_ptr<Pool>> pools_;
  std::unordered_map<int, Var> vars_;
  std::vector<int> doomed_;
  std::exception_ptr global_error_;
  int next_var_ = 0;
  long pending_ = 0;
  uint64_t seq_ = 0;
  bool stop_ = false;
};

#ifndef GEOMX_NO_PYTHON   // the Python-free C library (lib/libgeomx_capi.so) compiles the runtime without the pybind11 bindings
inline void BindEngine(py::module_& m) {
  py::class_<Engine>(m, "Engine")
      .def(py::init<int, bool>(), py::arg("num_threads") = 2, py::arg("naive") = false)
      .def("new_variable", &Engine::NewVariable)
      .def("delete_variable", &Engine::DeleteVariable, py::call_guard<py::gil_scoped_release>())
      .def("push", [](Engine& e, py::object fn, std::vector<int> cv, std::vector<int> mv, int priority, const std::string& name, int device, int prop) {
        // the callable is released under the GIL wherever the last reference dies (a worker thread, usually)
        std::shared_ptr<py::object> holder(new py::object(std::move(fn)), [](py::object* p) { py::gil_scoped_acquire g; delete p; });
        py::gil_scoped_release nogil;
        e.Push([holder] {
          py::gil_scoped_acquire g;
          (*holder)();          // a Python exception propagates as py::error_already_set: remembered on the written variables, re-raised at the wait
        }, cv, mv, priority, name, device, static_cast<FnProperty>(prop));
      }, py::arg("fn"), py::arg("const_vars") = std::vector<int>(), py::arg("mutable_vars") = std::vector<int>(), py::arg("priority") = 0,
           py::arg("name") = "", py::arg("device") = -1, py::arg("prop") = 0)
      .def("wait_for_var", &Engine::WaitForVar, py::call_guard<py::gil_scoped_release>())
      .def("wait_for_all", &Engine::WaitForAll, py::call_guard<py::gil_scoped_release>())
      .def("stats", &Engine::Stats)
      .def("num_variables", &Engine::NumVariables);
}
#endif

}  // namespace gxrt
