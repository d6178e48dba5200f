// BlockPool: process-wide cache of large message buffers (payload parts of received frames, pull responses, resized SArrays).
//
// Why: a parameter server moves the same few tensor sizes every round.  Allocating each payload with new[] hands every message a fresh
// anonymous mapping (glibc serves blocks above M_MMAP_THRESHOLD with mmap and returns them with munmap), so every byte of every message is
// first touched through a page fault: ~0.25 ms per MB on bare metal, ~4 ms per MB inside a micro-VM without huge pages — more than the
// loop-back transfer itself (profiles/tcp_plane_bench.txt, "buffer pool" section).  The pool keeps released blocks mapped and hands them
// out again by size class, so steady-state rounds touch only warm pages.  (The reference gets the same effect from ZeroMQ's message pool
// plus MXNet's pooled CPU storage, src/storage/cpu_device_storage.h; here it is one explicit component.)
//
//   * blocks below kMinPooled (64 KiB) are not pooled (the allocator's own bins are fine for them);
//   * size classes: powers of two up to 1 MiB, quarter steps of the enclosing power of two above (at most 25 % slack);
//   * the cache is bounded: PS_BUFFER_POOL_MB (default 1024; 0 disables pooling) — a release that would exceed it frees the block;
//   * the singleton is never destroyed: SArrays released by detached threads during interpreter shutdown stay valid.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>

namespace hips {

class BlockPool {
 public:
  static constexpr size_t kMinPooled = 64 << 10;
  static BlockPool* Get() {
    static BlockPool* p = new BlockPool();
    return p;
  }
  static size_t ClassOf(size_t n) {
    size_t p = kMinPooled;
    while (p < n) p <<= 1;
    if (p <= (size_t(1) << 20)) return p;
    const size_t step = p >> 3;                 // p/2 < n <= p: quarter steps of the lower power of two = eighths of the upper one
    return (n + step - 1) / step * step;
  }
  // a block of at least n bytes (64-byte aligned); *cap is what Release must be told.  nullptr when the system is out of memory.
  char* Acquire(size_t n, size_t* cap) {
    const size_t c = ClassOf(n);
    *cap = c;
    if (limit_) {
      std::lock_guard<std::mutex> lk(mu_);
      auto it = free_.find(c);
      if (it != free_.end() && !it->second.empty()) {
        char* b = it->second.back(); it->second.pop_back();
        cached_ -= c; ++hits_;
        return b;
      }
      ++misses_;
    }
    void* b = nullptr;
    if (posix_memalign(&b, 64, c) != 0) return nullptr;
    return static_cast<char*>(b);
  }
  void Release(char* b, size_t cap) {
    if (b == nullptr) return;
    if (limit_) {
      std::lock_guard<std::mutex> lk(mu_);
      if (cached_ + cap <= limit_) { free_[cap].push_back(b); cached_ += cap; return; }
    }
    free(b);
  }
  // give every cached block back to the system (tests, memory pressure)
  void Trim() {
    std::lock_guard<std::mutex> lk(mu_);
    for (auto& kv : free_) for (char* b : kv.second) free(b);
    free_.clear(); cached_ = 0;
  }
  void SetLimit(size_t bytes) { { std::lock_guard<std::mutex> lk(mu_); limit_ = bytes; } if (bytes == 0) Trim(); }
  // cached bytes, hits, misses, limit
  void Stats(uint64_t out[4]) { std::lock_guard<std::mutex> lk(mu_); out[0] = cached_; out[1] = hits_; out[2] = misses_; out[3] = limit_; }

 private:
  BlockPool() {
    const char* v = getenv("PS_BUFFER_POOL_MB");
    const long mb = v && *v ? atol(v) : 1024;
    limit_ = mb <= 0 ? 0 : static_cast<size_t>(mb) << 20;
  }
  std::mutex mu_;
  std::map<size_t, std::vector<char*>> free_;
  size_t cached_ = 0, limit_ = 0;
  uint64_t hits_ = 0, misses_ = 0;
};

}  // namespace hips
