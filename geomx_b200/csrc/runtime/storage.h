// Host storage pools + per-device resources (temp workspace, RNG seeds).
//
// Parity: src/storage/storage.cc:36-229 + pooled_storage_manager.h:52-172 (size-bucketed free lists, page rounding, MXNET_*_MEM_POOL_RESERVE
// style limits, ReleaseAll, DirectFree) for the HOST side — device memory is owned by PyTorch's caching allocator in this design — and
// src/resource.cc (ResourceManager: kTempSpace growing workspaces, kRandom per-device seeds derived from one global seed).
// Pinning is done by the caller (cudaHostRegister through PyTorch) so that this header carries no CUDA link dependency.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <map>
#include <mutex>
#include <random>
#include <unordered_map>
#include <vector>

namespace gx_rt {

class PooledHostStorage {
 public:
  struct Stats { size_t used_bytes = 0, pooled_bytes = 0, num_alloc = 0, num_pool_hits = 0, num_system_alloc = 0; };

  explicit PooledHostStorage(size_t page = 4096, size_t max_pooled = size_t(4) << 30) : page_(page), max_pooled_(max_pooled) {}
  ~PooledHostStorage() { ReleaseAll(); }

  // bucket: page multiples below 1 MiB (exact fit, little waste), powers of two above (few distinct large sizes, high reuse)
  size_t RoundSize(size_t n) const {
    if (n == 0) n = 1;
    if (n <= (size_t(1) << 20)) return (n + page_ - 1) / page_ * page_;
    size_t p = size_t(1) << 20;
    while (p < n) p <<= 1;
    return p;
  }

  void* Alloc(size_t nbytes, bool* from_pool = nullptr) {
    const size_t sz = RoundSize(nbytes);
    std::lock_guard<std::mutex> lk(mu_);
    ++stats_.num_alloc;
    auto it = free_.find(sz);
    void* p = nullptr;
    if (it != free_.end() && !it->second.empty()) {
      p = it->second.back(); it->second.pop_back();
      stats_.pooled_bytes -= sz; ++stats_.num_pool_hits;
      if (from_pool) *from_pool = true;
    } else {
      if (posix_memalign(&p, page_, sz) != 0) return nullptr;
      ++stats_.num_system_alloc;
      if (from_pool) *from_pool = false;
    }
    live_[p] = sz;
    stats_.used_bytes += sz;
    return p;
  }

  // returns true if the block went back to the pool, false if it was released to the system (pool limit) — the caller un-pins it then
  bool Free(void* p) {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = live_.find(p);
    if (it == live_.end()) return false;
    const size_t sz = it->second;
    live_.erase(it);
    stats_.used_bytes -= sz;
    if (stats_.pooled_bytes + sz <= max_pooled_) { free_[sz].push_back(p); stats_.pooled_bytes += sz; return true; }
    ::free(p);
    return false;
  }
  size_t SizeOf(void* p) { std::lock_guard<std::mutex> lk(mu_); auto it = live_.find(p); return it == live_.end() ? 0 : it->second; }

  std::vector<std::pair<void*, size_t>> ReleaseAll() {   // returns what was released so that the caller can un-pin
    std::lock_guard<std::mutex> lk(mu_);
    std::vector<std::pair<void*, size_t>> out;
    for (auto& kv : free_) for (void* p : kv.second) { out.emplace_back(p, kv.first); ::free(p); }
    free_.clear();
    stats_.pooled_bytes = 0;
    return out;
  }
  Stats stats() { std::lock_guard<std::mutex> lk(mu_); return stats_; }

 private:
  size_t page_, max_pooled_;
  std::mutex mu_;
  std::map<size_t, std::vector<void*>> free_;
  std::unordered_map<void*, size_t> live_;
  Stats stats_;
};

// kTempSpace: one growing workspace per (device, slot); kRandom: reproducible per-device seed streams from one global seed
class ResourceManager {
 public:
  void* TempSpace(int device, int slot, size_t nbytes, PooledHostStorage* pool) {
    std::lock_guard<std::mutex> lk(mu_);
    auto& ws = temp_[{device, slot}];
    if (ws.second < nbytes) {
      if (ws.first) pool->Free(ws.first);
      ws.first = pool->Alloc(nbytes);
      ws.second = pool->SizeOf(ws.first);
    }
    return ws.first;
  }
  void SeedAll(uint64_t seed) { std::lock_guard<std::mutex> lk(mu_); seed_ = seed; counters_.clear(); }
  uint64_t NextSeed(int device) {
    std::lock_guard<std::mutex> lk(mu_);
    const uint64_t c = counters_[device]++;
    uint64_t z = seed_ + 0x9e3779b97f4a7c15ull * (static_cast<uint64_t>(device) * 0x10001ull + c + 1);   // splitmix64
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
  }

 private:
  std::mutex mu_;
  std::map<std::pair<int, int>, std::pair<void*, size_t>> temp_;
  std::map<int, uint64_t> counters_;
  uint64_t seed_ = 0;
};

}  // namespace gx_rt
