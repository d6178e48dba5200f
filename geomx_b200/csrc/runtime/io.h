// Native data IO: idx (MNIST) file parsing and a threaded batch prefetcher that assembles float batches.
// Parity: src/io/iter_mnist.cc:80-260 (MNISTIter: idx magic parsing, scaling to [0,1], shuffle / partition), src/io/iter_prefetcher.h
// (background producer with a bounded queue).  Batches are written straight into caller-provided (pinned) host buffers.
#pragma once
#ifndef GEOMX_NO_PYTHON   // the Python-free C library (lib/libgeomx_capi.so) compiles the runtime without the pybind11 bindings
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>
#endif

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <fstream>
#include <mutex>
#include <queue>
#include <random>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace gxrt {
#ifndef GEOMX_NO_PYTHON   // the Python-free C library (lib/libgeomx_capi.so) compiles the runtime without the pybind11 bindings
namespace py = pybind11;
#endif

inline uint32_t BE32(const unsigned char* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }

inline std::pair<std::vector<int64_t>, std::string> ReadIdx(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f.good()) throw std::runtime_error("cannot open " + path);
  std::string buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  if (buf.size() < 4) throw std::runtime_error("bad idx file " + path);
  const unsigned char* p = reinterpret_cast<const unsigned char*>(buf.data());
  const int ndim = p[3];
  std::vector<int64_t> dims(ndim);
  for (int i = 0; i < ndim; ++i) dims[i] = BE32(p + 4 + 4 * i);
  return {dims, buf.substr(4 + 4 * ndim)};
}

#ifndef GEOMX_NO_PYTHON   // the Python-free C library (lib/libgeomx_capi.so) compiles the runtime without the pybind11 bindings
// Prefetcher over an in-memory uint8 image set: worker threads build (float image batch scaled by 1/255, float label batch)
class BatchPrefetcher {
 public:
  BatchPrefetcher(py::array_t<uint8_t, py::array::c_style> images, py::array_t<int32_t, py::array::c_style> labels, int batch, bool shuffle,
                  int seed, int part_index, int num_parts, int depth)
      : batch_(batch), shuffle_(shuffle), rng_(seed), depth_(std::max(1, depth)) {
    const auto n_all = images.shape(0);
    item_ = images.size() / std::max<py::ssize_t>(1, n_all);
    const py::ssize_t part = n_all / std::max(1, num_parts);
    begin_ = part * part_index; n_ = part;
    img_.assign(images.data() + begin_ * item_, images.data() + (begin_ + n_) * item_);
    lab_.assign(labels.data() + begin_, labels.data() + begin_ + n_);
    Reset();
  }
  ~BatchPrefetcher() { Stop(); }
  void Reset() {
    Stop();
    order_.resize(n_);
    for (py::ssize_t i = 0; i < n_; ++i) order_[i] = i;
    if (shuffle_) std::shuffle(order_.begin(), order_.end(), rng_);
    cursor_ = 0; stop_ = false; done_ = false;
    worker_ = std::thread([this] { Produce(); });
  }
  py::ssize_t num_batches() const { return n_ / batch_; }
  // copies the next batch into x (B*item floats) / y (B floats); returns false at end of epoch
  bool Next(uintptr_t x_ptr, uintptr_t y_ptr) {
    Item it;
    {
      std::unique_lock<std::mutex> lk(mu_);
      cv_.wait(lk, [this] { return !q_.empty() || done_; });
      if (q_.empty()) return false;
      it = std::move(q_.front()); q_.pop();
    }
    cv_space_.notify_one();
    memcpy(reinterpret_cast<float*>(x_ptr), it.x.data(), it.x.size() * sizeof(float));
    memcpy(reinterpret_cast<float*>(y_ptr), it.y.data(), it.y.size() * sizeof(float));
    return true;
  }

 private:
  struct Item { std::vector<float> x, y; };
  void Stop() {
    stop_ = true;
    cv_space_.notify_all();
    if (worker_.joinable()) worker_.join();
    std::queue<Item>().swap(q_);
  }
  void Produce() {
    while (!stop_ && cursor_ + batch_ <= n_) {
      Item it; it.x.resize(batch_ * item_); it.y.resize(batch_);
      for (int b = 0; b < batch_; ++b) {
        const py::ssize_t idx = order_[cursor_ + b];
        const uint8_t* src = img_.data() + idx * item_;
        float* dst = it.x.data() + b * item_;
        for (py::ssize_t j = 0; j < item_; ++j) dst[j] = src[j] * (1.f / 255.f);
        it.y[b] = static_cast<float>(lab_[idx]);
      }
      cursor_ += batch_;
      std::unique_lock<std::mutex> lk(mu_);
      cv_space_.wait(lk, [this] { return static_cast<int>(q_.size()) < depth_ || stop_; });
      if (stop_) break;
      q_.push(std::move(it));
      cv_.notify_one();
    }
    { std::lock_guard<std::mutex> lk(mu_); done_ = true; }
    cv_.notify_all();
  }
  int batch_; bool shuffle_; std::mt19937 rng_; int depth_;
  py::ssize_t item_ = 0, begin_ = 0, n_ = 0, cursor_ = 0;
  std::vector<uint8_t> img_; std::vector<int32_t> lab_; std::vector<py::ssize_t> order_;
  std::thread worker_;
  std::mutex mu_; std::condition_variable cv_, cv_space_;
  std::queue<Item> q_;
  std::atomic<bool> stop_{false};
  bool done_ = false;
};

inline void BindIO(py::module_& m) {
  m.def("read_idx", [](const std::string& path) {
    auto r = ReadIdx(path);
    return py::make_tuple(r.first, py::bytes(r.second));
  }, "parse an (uncompressed) idx file -> (dims, raw bytes)");
  py::class_<BatchPrefetcher>(m, "BatchPrefetcher")
      .def(py::init<py::array_t<uint8_t, py::array::c_style>, py::array_t<int32_t, py::array::c_style>, int, bool, int, int, int, int>(),
           py::arg("images"), py::arg("labels"), py::arg("batch"), py::arg("shuffle") = false, py::arg("seed") = 0, py::arg("part_index") = 0,
           py::arg("num_parts") = 1, py::arg("depth") = 4)
      .def("reset", &BatchPrefetcher::Reset)
      .def("num_batches", &BatchPrefetcher::num_batches)
      .def("next", &BatchPrefetcher::Next, py::call_guard<py::gil_scoped_release>());
}
#endif

}  // namespace gxrt
