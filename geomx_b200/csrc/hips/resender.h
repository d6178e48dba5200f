// Reliable delivery on top of the Van's datagram-like sends: acknowledgements, deadline-driven retransmission with exponential back-off and
// bounded duplicate suppression.  Enabled by PS_RESEND=1; PS_RESEND_TIMEOUT (ms) is the first retransmission delay.
//
// Capability parity with the reference's optional resender (3rdparty/ps-lite/src/resender.h:15-139: every non-ACK message is acknowledged,
// unacknowledged ones are sent again, duplicates are dropped), designed differently:
//   * retransmissions are scheduled, not polled: pending messages sit in a min-heap keyed by their next deadline and ONE timer thread sleeps
//     on a condition variable until the earliest deadline (or until a new message / shutdown wakes it) — no fixed-period scan of every
//     in-flight message, and an ACK is O(1);
//   * the delay doubles per attempt (timeout, 2x, 4x, ... capped at 32x), so a slow inter-party link is not flooded;
//   * the duplicate filter is a per-sender sliding window of the most recent signatures (kWindow each) instead of a set that grows for the
//     life time of the process: memory is bounded by the number of peers, and a signature can only be replayed within the retransmission
//     horizon anyway;
//   * `WaitDrained` lets a node flush its last messages (e.g. the final barrier release) before it tears the transport down.
// A message signature covers everything that distinguishes two logical messages between the same pair of nodes (app, customer, timestamp,
// request/push/simple flags, control command, sequence number and command head).
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <queue>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "message.h"
#include "van.h"

namespace hips {

class Resender {
 public:
  Resender(int first_delay_ms, int max_attempts, Van* van)
      : first_delay_(std::max(1, first_delay_ms)), max_attempts_(max_attempts), van_(van), timer_([this] { TimerLoop(); }) {}

  ~Resender() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stopping_ = true;
    }
    wake_.notify_all();
    timer_.join();
  }

  // Block (bounded) until every tracked message has been acknowledged.
  void WaitDrained(int max_ms) {
    std::unique_lock<std::mutex> lk(mu_);
    drained_.wait_for(lk, std::chrono::milliseconds(max_ms), [this] { return inflight_.empty(); });
  }

  // Track an outgoing message until its acknowledgement arrives.
  void AddOutgoing(const Message& msg) {
    if (msg.meta.control.cmd == Control::ACK) return;
    const uint64_t sig = Signature(msg.meta);
    const Clock::time_point due = Clock::now() + std::chrono::milliseconds(first_delay_);
    bool earliest;
    {
      std::lock_guard<std::mutex> lk(mu_);
      if (!inflight_.emplace(sig, Pending{msg, 0}).second) return;      // a retransmission of something already tracked
      earliest = schedule_.empty() || due < schedule_.top().due;
      schedule_.push(Deadline{due, sig});
    }
    if (earliest) wake_.notify_one();
  }

  // Returns true when `msg` must NOT be delivered to the application: it is an acknowledgement (consumed here) or a duplicate.
  bool AddIncoming(const Message& msg) {
    const Meta& m = msg.meta;
    if (m.control.cmd == Control::TERMINATE) return false;
    if (m.control.cmd == Control::ACK) {
      std::lock_guard<std::mutex> lk(mu_);
      inflight_.erase(m.control.msg_sig);          // its heap entry becomes stale and is skipped by the timer
      if (inflight_.empty()) drained_.notify_all();
      return true;
    }
    const uint64_t sig = Signature(m);
    bool seen_before;
    {
      std::lock_guard<std::mutex> lk(mu_);
      seen_before = !windows_[m.sender].Insert(sig);
    }
    Message ack;                                    // always acknowledge: the sender may have missed the first ACK
    ack.meta.sender = m.recver;
    ack.meta.recver = m.sender;
    ack.meta.control.cmd = Control::ACK;
    ack.meta.control.msg_sig = sig;
    van_->SendNow(ack);
    return seen_before;
  }

 private:
  using Clock = std::chrono::steady_clock;
  struct Pending { Message msg; int attempts; };
  struct Deadline {
    Clock::time_point due; uint64_t sig;
    bool operator>(const Deadline& o) const { return due > o.due; }
  };
  // most recent signatures of one peer: FIFO of kWindow entries + hash index
  struct Window {
    static constexpr size_t kWindow = 4096;
    std::deque<uint64_t> order;
    std::unordered_set<uint64_t> index;
    bool Insert(uint64_t sig) {                    // false if already present
      if (!index.insert(sig).second) return false;
      order.push_back(sig);
      if (order.size() > kWindow) { index.erase(order.front()); order.pop_front(); }
      return true;
    }
  };

  static uint64_t Signature(const Meta& m) {
    // 64-bit FNV-1a over the identifying fields, one field per round
    const uint64_t fields[] = {static_cast<uint64_t>(m.app_id), static_cast<uint64_t>(m.customer_id), static_cast<uint64_t>(m.timestamp),
                               static_cast<uint64_t>(m.sender), static_cast<uint64_t>(m.recver),
                               (m.request ? 1ull : 0ull) | (m.push ? 2ull : 0ull) | (m.simple_app ? 4ull : 0ull),
                               static_cast<uint64_t>(m.control.cmd), static_cast<uint64_t>(m.seq), static_cast<uint64_t>(m.head)};
    uint64_t h = 0xcbf29ce484222325ull;
    for (uint64_t f : fields)
      for (int b = 0; b < 8; ++b) { h ^= (f >> (8 * b)) & 0xffu; h *= 0x100000001b3ull; }
    return h;
  }

  void TimerLoop() {
    std::unique_lock<std::mutex> lk(mu_);
    while (!stopping_) {
      if (schedule_.empty()) { wake_.wait(lk); continue; }
      const Deadline next = schedule_.top();
      if (Clock::now() < next.due) { wake_.wait_until(lk, next.due); continue; }
      schedule_.pop();
      auto it = inflight_.find(next.sig);
      if (it == inflight_.end()) continue;          // acknowledged in the meantime
      Pending& p = it->second;
      ++p.attempts;
      HIPS_CHECK_MSG(p.attempts < max_attempts_, "message retransmitted too many times: peer " + std::to_string(p.msg.meta.recver) + " is unreachable");
      const int backoff = 1 << std::min(p.attempts, 5);
      schedule_.push(Deadline{Clock::now() + std::chrono::milliseconds(static_cast<long long>(first_delay_) * backoff), next.sig});
      Message copy = p.msg;
      lk.unlock();                                  // never hold the lock across a socket write
      van_->SendNow(copy);
      lk.lock();
    }
  }

  const int first_delay_, max_attempts_;
  Van* const van_;
  std::mutex mu_;
  std::condition_variable wake_, drained_;
  bool stopping_ = false;
  std::unordered_map<uint64_t, Pending> inflight_;
  std::priority_queue<Deadline, std::vector<Deadline>, std::greater<Deadline>> schedule_;
  std::unordered_map<int, Window> windows_;
  std::thread timer_;          // last member: starts after everything above is constructed
};

}  // namespace hips
