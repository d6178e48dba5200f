// Dependency engine: versioned variables + read/write dependency tracking + prioritised worker pool.
// Parity: include/mxnet/engine.h:95-314 (NewVariable / PushAsync(const_vars, mutable_vars, priority) / WaitForVar / WaitForAll),
// src/engine/threaded_engine.h:66-553 (per-variable pending queue of readers/writers, ready when all deps resolved) and
// src/engine/naive_engine.cc (MXNET_ENGINE_TYPE=NaiveEngine runs everything inline for debugging races).
// On B200 device work is ordered by CUDA streams; this engine schedules the HOST side of the framework (kvstore sends/receives,
// staging copies, checkpoint IO, Python callbacks) with the same var semantics, and honours `priority` in its ready queue (the
// reference's normal CPU pool is FIFO).  Duplicate vars in const/mutable sets are rejected like threaded_engine.h:432.
#pragma once
#include <pybind11/functional.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <queue>
#include <set>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace gxrt {
namespace py = pybind11;

class Engine {
 public:
  using Fn = std::function<void()>;
  explicit Engine(int num_threads, bool naive) : naive_(naive) {
    if (!naive_) for (int i = 0; i < std::max(1, num_threads); ++i) workers_.emplace_back([this] { Work(); });
  }
  ~Engine() {
    { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  int NewVariable() { std::lock_guard<std::mutex> lk(mu_); vars_.emplace(next_var_, Var()); return next_var_++; }

  void Push(Fn fn, const std::vector<int>& const_vars, const std::vector<int>& mutable_vars, int priority, const std::string& name) {
    {
      std::set<int> seen;
      for (int v : mutable_vars) if (!seen.insert(v).second) throw std::runtime_error("duplicate variable in mutable_vars");
      for (int v : const_vars) if (seen.count(v)) throw std::runtime_error("variable appears in both const_vars and mutable_vars");
    }
    if (naive_) { fn(); return; }
    auto op = std::make_shared<Op>();
    op->fn = std::move(fn); op->priority = priority; op->name = name; op->reads = const_vars; op->writes = mutable_vars;
    std::unique_lock<std::mutex> lk(mu_);
    ++pending_;
    int wait = 0;
    for (int v : const_vars) { Var& var = vars_.at(v); if (var.writer_active || !var.queue.empty()) { var.queue.push_back({op, false}); ++wait; } else ++var.readers; }
    for (int v : mutable_vars) { Var& var = vars_.at(v); if (var.writer_active || var.readers > 0 || !var.queue.empty()) { var.queue.push_back({op, true}); ++wait; } else var.writer_active = true; }
    op->wait = wait;
    if (wait == 0) { ready_.push(op); lk.unlock(); cv_.notify_one(); }
  }
  void WaitForVar(int v) {
    std::mutex m; std::condition_variable c; bool done = false;
    Push([&] { std::lock_guard<std::mutex> lk(m); done = true; c.notify_all(); }, {v}, {}, 1 << 20, "WaitForVar");
    std::unique_lock<std::mutex> lk(m);
    c.wait(lk, [&] { return done; });
  }
  void WaitForAll() {
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [this] { return pending_ == 0; });
  }

 private:
  struct Op { Fn fn; int priority = 0; std::string name; std::vector<int> reads, writes; int wait = 0; uint64_t seq = 0; };
  struct Pending { std::shared_ptr<Op> op; bool write; };
  struct Var { int readers = 0; bool writer_active = false; std::deque<Pending> queue; };
  struct Cmp { bool operator()(const std::shared_ptr<Op>& a, const std::shared_ptr<Op>& b) const { return a->priority != b->priority ? a->priority < b->priority : a->seq > b->seq; } };

  void Work() {
    while (true) {
      std::shared_ptr<Op> op;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [this] { return stop_ || !ready_.empty(); });
        if (stop_ && ready_.empty()) return;
        op = ready_.top(); ready_.pop();
      }
      op->fn();
      Complete(op);
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
  void Complete(const std::shared_ptr<Op>& op) {
    std::vector<std::shared_ptr<Op>> runnable;
    {
      std::lock_guard<std::mutex> lk(mu_);
      for (int v : op->reads) { Var& var = vars_.at(v); --var.readers; Grant(var, &runnable); }
      for (int v : op->writes) { Var& var = vars_.at(v); var.writer_active = false; Grant(var, &runnable); }
      for (auto& r : runnable) { r->seq = seq_++; ready_.push(r); }
      if (--pending_ == 0) done_cv_.notify_all();
    }
    for (size_t i = 0; i < runnable.size(); ++i) cv_.notify_one();
  }
  bool naive_;
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_, done_cv_;
  std::priority_queue<std::shared_ptr<Op>, std::vector<std::shared_ptr<Op>>, Cmp> ready_;
  std::unordered_map<int, Var> vars_;
  int next_var_ = 0;
  long pending_ = 0;
  uint64_t seq_ = 0;
  bool stop_ = false;
};

inline void BindEngine(py::module_& m) {
  py::class_<Engine>(m, "Engine")
      .def(py::init<int, bool>(), py::arg("num_threads") = 2, py::arg("naive") = false)
      .def("new_variable", &Engine::NewVariable)
      .def("push", [](Engine& e, py::object fn, std::vector<int> cv, std::vector<int> mv, int priority, const std::string& name) {
        // the callable is released under the GIL wherever the last reference dies (a worker thread, usually)
        std::shared_ptr<py::object> holder(new py::object(std::move(fn)), [](py::object* p) { py::gil_scoped_acquire g; delete p; });
        py::gil_scoped_release nogil;
        e.Push([holder] {
          py::gil_scoped_acquire g;
          try { (*holder)(); }
          catch (py::error_already_set& err) { err.restore(); PyErr_WriteUnraisable(holder->ptr()); }   // a failing op must not kill the worker
        }, cv, mv, priority, name);
      }, py::arg("fn"), py::arg("const_vars") = std::vector<int>(), py::arg("mutable_vars") = std::vector<int>(), py::arg("priority") = 0,
           py::arg("name") = "")
      .def("wait_for_var", &Engine::WaitForVar, py::call_guard<py::gil_scoped_release>())
      .def("wait_for_all", &Engine::WaitForAll, py::call_guard<py::gil_scoped_release>());
}

}  // namespace gxrt
