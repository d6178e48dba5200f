// Thread-safe PRIORITY queue on Message.meta.priority (parity: ps-lite internal/threadsafe_queue.h:19-63 — GeoMX turned the FIFO
// into a priority queue for P3).  Equal priorities keep FIFO order (a monotonically increasing sequence breaks ties), which the
// reference does not guarantee.
#pragma once
#include <condition_variable>
#include <mutex>
#include <queue>
#include <vector>

namespace hips {

template <typename T, typename PriorityOf>
class ThreadsafeQueue {
 public:
  void Push(T v) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      q_.push(Item{PriorityOf()(v), seq_++, std::move(v)});
    }
    cv_.notify_one();
  }
  void WaitAndPop(T* out) {
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [this] { return !q_.empty(); });
    *out = std::move(const_cast<Item&>(q_.top()).v);
    q_.pop();
  }
  bool TryPop(T* out) {
    std::lock_guard<std::mutex> lk(mu_);
    if (q_.empty()) return false;
    *out = std::move(const_cast<Item&>(q_.top()).v);
    q_.pop();
    return true;
  }
  size_t Size() {
    std::lock_guard<std::mutex> lk(mu_);
    return q_.size();
  }

 private:
  struct Item {
    int pri; uint64_t seq; T v;
    bool operator<(const Item& o) const { return pri != o.pri ? pri < o.pri : seq > o.seq; }
  };
  std::mutex mu_;
  std::condition_variable cv_;
  std::priority_queue<Item> q_;
  uint64_t seq_ = 0;
};

}  // namespace hips
