// SArray<T>: reference-counted, zero-copy, sliceable array (parity: 3rdparty/ps-lite/include/ps/sarray.h).
#pragma once
#include <cstring>
#include <memory>
#include <vector>

#include <type_traits>

#include "base.h"
#include "block_pool.h"

namespace hips {

template <typename T>
class SArray {
 public:
  SArray() {}
  explicit SArray(size_t n, T val = T()) { resize(n, val); }
  // wrap external memory; `deletable=false` -> caller keeps ownership (zero copy)
  SArray(T* data, size_t size, bool deletable = false) { reset(data, size, deletable); }
  explicit SArray(const std::vector<T>& v) { CopyFrom(v.data(), v.size()); }
  template <typename W>
  explicit SArray(const SArray<W>& o) { *this = o; }
  // reinterpret (zero copy), like ps-lite's templated assignment
  template <typename W>
  SArray<T>& operator=(const SArray<W>& o) {
    size_ = o.size() * sizeof(W) / sizeof(T);
    HIPS_CHECK(size_ * sizeof(T) == o.size() * sizeof(W));
    capacity_ = size_;
    ptr_ = std::shared_ptr<T>(o.ptr(), reinterpret_cast<T*>(o.data()));
    return *this;
  }
  void reset(T* data, size_t size, bool deletable) {
    size_ = capacity_ = size;
    if (deletable) ptr_.reset(data, [](T* p) { delete[] p; });
    else ptr_.reset(data, [](T*) {});
  }
  // n uninitialised elements.  Large arrays (message payloads) come from the process-wide BlockPool and go back to it when the last
  // reference dies, so that steady-state rounds reuse warm pages instead of faulting in a fresh mapping per message (block_pool.h).
  void Allocate(size_t n) {
    static_assert(std::is_trivially_copyable<T>::value, "SArray holds plain data");
    const size_t bytes = n * sizeof(T);
    if (bytes >= BlockPool::kMinPooled) {
      size_t cap = 0;
      char* b = BlockPool::Get()->Acquire(bytes, &cap);
      HIPS_CHECK_MSG(b != nullptr, "out of memory: " + std::to_string(bytes) + " bytes");
      size_ = capacity_ = n;
      ptr_.reset(reinterpret_cast<T*>(b), [cap](T* p) { BlockPool::Get()->Release(reinterpret_cast<char*>(p), cap); });
    } else {
      reset(new T[n ? n : 1], n, true);
    }
  }
  void resize(size_t n, T val = T()) {
    if (n <= capacity_) { size_ = n; return; }
    SArray<T> nd;
    nd.Allocate(n);
    if (size_) memcpy(nd.data(), data(), size_ * sizeof(T));
    for (size_t i = size_; i < n; ++i) nd.data()[i] = val;
    ptr_ = nd.ptr_; size_ = capacity_ = n;
  }
  void CopyFrom(const T* src, size_t n) {
    SArray<T> nd;
    nd.Allocate(n);
    if (n) memcpy(nd.data(), src, n * sizeof(T));
    ptr_ = nd.ptr_; size_ = capacity_ = n;
  }
  void CopyFrom(const SArray<T>& o) { if (this != &o) CopyFrom(o.data(), o.size()); }
  SArray<T> segment(size_t begin, size_t end) const {
    HIPS_CHECK(end >= begin && end <= size_);
    SArray<T> r;
    r.ptr_ = std::shared_ptr<T>(ptr_, data() + begin);
    r.size_ = r.capacity_ = end - begin;
    return r;
  }
  void push_back(const T& v) {
    if (size_ == capacity_) {
      size_t nc = capacity_ ? capacity_ * 2 : 4, os = size_;
      resize(nc); size_ = os;
    }
    data()[size_++] = v;
  }
  void clear() { ptr_.reset(); size_ = capacity_ = 0; }
  size_t size() const { return size_; }
  bool empty() const { return size_ == 0; }
  T* data() const { return ptr_.get(); }
  T* begin() const { return data(); }
  T* end() const { return data() + size_; }
  const std::shared_ptr<T>& ptr() const { return ptr_; }
  T& operator[](size_t i) const { return data()[i]; }
  T& front() const { return data()[0]; }
  T& back() const { return data()[size_ - 1]; }

 private:
  std::shared_ptr<T> ptr_;
  size_t size_ = 0, capacity_ = 0;
};

struct Range {
  Range() : b(0), e(0) {}
  Range(uint64_t begin, uint64_t end) : b(begin), e(end) {}
  uint64_t begin() const { return b; }
  uint64_t end() const { return e; }
  uint64_t size() const { return e - b; }
  uint64_t b, e;
};

}  // namespace hips
