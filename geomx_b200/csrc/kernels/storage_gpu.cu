// Pooled device-memory manager: size-bucketed free lists per (device, stream) over cudaMalloc, the native counterpart of the host pool in
// csrc/runtime/storage.h.
//
// Parity: src/storage/pooled_storage_manager.h — GPUPooledStorageManager (:52-172: exact page-rounded buckets, MXNET_GPU_MEM_POOL_PAGE_SIZE,
// ReleaseAll when an allocation would eat into the MXNET_GPU_MEM_POOL_RESERVE percent of the device) and GPUPooledRoundedStorageManager
// (:175-330: power-of-two buckets above MXNET_GPU_MEM_POOL_ROUND_LINEAR_CUTOFF), selected by MXNET_GPU_MEM_POOL_TYPE = Naive | Round |
// Unpooled (src/storage/storage.cc:106-150).  Differences, because this pool can also serve PyTorch as a pluggable allocator
// (gx_torch_alloc / gx_torch_free, storage.py::use_native_gpu_pool):
//   * free lists are keyed by (stream, size): a block is reused only by work queued on the stream that last used it, so no event
//     bookkeeping is needed for stream safety (the reference serialises through the engine's per-device worker instead);
//   * 180 GB of HBM3e: sizes are 64-bit throughout and buckets above 1 GiB round to 1/8 of the next power of two instead of the full
//     power (a 65 GiB tensor must not reserve 128 GiB);
//   * a simulated backend (device < 0: host malloc with a configurable capacity) runs the same policy on machines without a GPU — that is
//     what the CPU unit tests drive.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#define GX_API extern "C" __attribute__((visibility("default")))

namespace {

struct Backend {
  virtual ~Backend() {}
  virtual void* Raw(size_t n) = 0;
  virtual void Release(void* p, size_t n) = 0;
  virtual void MemInfo(size_t* free_b, size_t* total_b) = 0;
};

struct CudaBackend : Backend {
  int dev;
  explicit CudaBackend(int d) : dev(d) {}
  void* Raw(size_t n) override {
    int cur = 0; cudaGetDevice(&cur);
    if (cur != dev) cudaSetDevice(dev);
    void* p = nullptr;
    const cudaError_t e = cudaMalloc(&p, n);
    if (cur != dev) cudaSetDevice(cur);
    if (e != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
  }
  void Release(void* p, size_t) override { cudaFree(p); }
  void MemInfo(size_t* f, size_t* t) override {
    int cur = 0; cudaGetDevice(&cur);
    if (cur != dev) cudaSetDevice(dev);
    if (cudaMemGetInfo(f, t) != cudaSuccess) { cudaGetLastError(); *f = *t = 0; }
    if (cur != dev) cudaSetDevice(cur);
  }
};

// host malloc behind a fixed capacity: the policy under test is the same, only the raw allocator differs
struct SimBackend : Backend {
  size_t capacity, used = 0;
  explicit SimBackend(size_t cap) : capacity(cap) {}
  void* Raw(size_t n) override { if (used + n > capacity) return nullptr; void* p = malloc(64); if (p) used += n; return p; }   // 64-byte stubs: the tests only look at addresses
  void Release(void* p, size_t n) override { free(p); used -= n; }
  void MemInfo(size_t* f, size_t* t) override { *t = capacity; *f = capacity - used; }
};

enum PoolType { kNaive = 0, kRound = 1, kUnpooled = 2 };

class DevicePool {
 public:
  DevicePool(std::unique_ptr<Backend> b, PoolType type, size_t page, int reserve_pct, int cutoff_log2)
      : backend_(std::move(b)), type_(type), page_(page < 32 ? 32 : page), reserve_(reserve_pct), cutoff_(cutoff_log2) {}
  ~DevicePool() { ReleaseAll(); }

  size_t RoundSize(size_t n) const {
    if (n == 0) n = 1;
    if (type_ == kUnpooled) return (n + 255) / 256 * 256;
    if (type_ == kNaive || n <= (size_t(1) << cutoff_)) return (n + page_ - 1) / page_ * page_;
    size_t p = size_t(1) << cutoff_;
    while (p < n) p <<= 1;
    if (p > (size_t(1) << 30)) { const size_t step = p >> 4; return (n + step - 1) / step * step; }     // above 1 GiB: sixteenths of the enclosing power of two
    return p;
  }

  void* Alloc(size_t nbytes, uintptr_t stream) {
    const size_t sz = RoundSize(nbytes);
    std::lock_guard<std::mutex> lk(mu_);
    ++stats_[2];
    if (type_ != kUnpooled) {
      auto it = free_.find({stream, sz});
      if (it != free_.end() && !it->second.empty()) {
        void* p = it->second.back(); it->second.pop_back();
        stats_[1] -= sz; ++stats_[3];
        live_[p] = {sz, stream}; stats_[0] += sz;
        return p;
      }
    }
    // would this allocation eat into the reserve?  then give the cached blocks back first (pooled_storage_manager.h:129-140)
    size_t free_b = 0, total_b = 0;
    backend_->MemInfo(&free_b, &total_b);
    if (total_b && free_b < sz + total_b / 100 * static_cast<size_t>(reserve_)) ReleaseAllLocked();
    void* p = backend_->Raw(sz);
    if (!p) { ReleaseAllLocked(); p = backend_->Raw(sz); }
    if (!p) return nullptr;
    ++stats_[4];
    live_[p] = {sz, stream}; stats_[0] += sz;
    return p;
  }

  // 0 = pooled, 1 = released to the driver, -1 = not a block of this pool
  int Free(void* p, uintptr_t stream, bool have_stream) {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = live_.find(p);
    if (it == live_.end()) return -1;
    const size_t sz = it->second.first;
    const uintptr_t st = have_stream ? stream : it->second.second;
    live_.erase(it);
    stats_[0] -= sz;
    if (type_ == kUnpooled) { backend_->Release(p, sz); return 1; }
    free_[{st, sz}].push_back(p); stats_[1] += sz;
    return 0;
  }
  void ReleaseAll() { std::lock_guard<std::mutex> lk(mu_); ReleaseAllLocked(); }
  void Stats(uint64_t out[5]) { std::lock_guard<std::mutex> lk(mu_); for (int i = 0; i < 5; ++i) out[i] = stats_[i]; }

 private:
  void ReleaseAllLocked() {
    for (auto& kv : free_) for (void* p : kv.second) backend_->Release(p, kv.first.second);
    free_.clear(); stats_[1] = 0;
  }
  std::unique_ptr<Backend> backend_;
  PoolType type_;
  size_t page_;
  int reserve_, cutoff_;
  std::mutex mu_;
  std::map<std::pair<uintptr_t, size_t>, std::vector<void*>> free_;
  std::unordered_map<void*, std::pair<size_t, uintptr_t>> live_;
  uint64_t stats_[5] = {0, 0, 0, 0, 0};      // used bytes, cached bytes, allocations, pool hits, driver allocations
};

std::mutex reg_mu;
std::map<int, std::unique_ptr<DevicePool>> pools;       // device id; negative ids are simulated devices

long EnvLong(const char* k, long def) { const char* v = getenv(k); return v && *v ? atol(v) : def; }
PoolType EnvType() {
  const char* v = getenv("MXNET_GPU_MEM_POOL_TYPE");
  const std::string s = v ? v : "Naive";
  return s == "Round" ? kRound : s == "Unpooled" ? kUnpooled : kNaive;
}
DevicePool* Pool(int dev) {
  std::lock_guard<std::mutex> lk(reg_mu);
  auto it = pools.find(dev);
  if (it != pools.end()) return it->second.get();
  if (dev < 0) return nullptr;                          // simulated devices are created explicitly
  pools[dev].reset(new DevicePool(std::unique_ptr<Backend>(new CudaBackend(dev)), EnvType(), static_cast<size_t>(EnvLong("MXNET_GPU_MEM_POOL_PAGE_SIZE", 4096)),
                                  static_cast<int>(EnvLong("MXNET_GPU_MEM_POOL_RESERVE", 5)), static_cast<int>(EnvLong("MXNET_GPU_MEM_POOL_ROUND_LINEAR_CUTOFF", 24))));
  return pools[dev].get();
}
}  // namespace

// a simulated device (dev < 0) of `capacity` bytes with an explicit policy: type 0 Naive / 1 Round / 2 Unpooled
GX_API int gx_gpu_pool_create_sim(int dev, uint64_t capacity, int type, uint64_t page, int reserve_pct, int cutoff_log2) {
  if (dev >= 0 || type < 0 || type > 2 || reserve_pct < 0 || reserve_pct > 100 || cutoff_log2 < 5 || cutoff_log2 > 40) return -1;
  std::lock_guard<std::mutex> lk(reg_mu);
  pools[dev].reset(new DevicePool(std::unique_ptr<Backend>(new SimBackend(capacity)), static_cast<PoolType>(type), page, reserve_pct, cutoff_log2));
  return 0;
}
GX_API int gx_gpu_pool_destroy(int dev) { std::lock_guard<std::mutex> lk(reg_mu); return pools.erase(dev) ? 0 : -1; }
GX_API void* gx_gpu_pool_alloc(int dev, uint64_t nbytes, void* stream) { DevicePool* p = Pool(dev); return p ? p->Alloc(nbytes, reinterpret_cast<uintptr_t>(stream)) : nullptr; }
GX_API int gx_gpu_pool_free(int dev, void* ptr, void* stream) { DevicePool* p = Pool(dev); return p ? p->Free(ptr, reinterpret_cast<uintptr_t>(stream), true) : -1; }
GX_API int gx_gpu_pool_release_all(int dev) { DevicePool* p = Pool(dev); if (!p) return -1; p->ReleaseAll(); return 0; }
GX_API uint64_t gx_gpu_pool_round_size(int dev, uint64_t nbytes) { DevicePool* p = Pool(dev); return p ? p->RoundSize(nbytes) : 0; }
// out: used bytes, cached bytes, allocations, pool hits, driver allocations
GX_API int gx_gpu_pool_stats(int dev, uint64_t* out) { DevicePool* p = Pool(dev); if (!p) return -1; p->Stats(out); return 0; }

// torch.cuda.memory.CUDAPluggableAllocator entry points
GX_API void* gx_torch_alloc(ssize_t size, int device, cudaStream_t stream) { return gx_gpu_pool_alloc(device, static_cast<uint64_t>(size), stream); }
GX_API void gx_torch_free(void* ptr, ssize_t, int device, cudaStream_t stream) { gx_gpu_pool_free(device, ptr, stream); }
