// Resender: ACK + timeout resend + duplicate suppression (parity: ps-lite src/resender.h:15-139; enabled by PS_RESEND=1,
// PS_RESEND_TIMEOUT ms).  Signature = hash of (app, customer, timestamp, sender, recver, request, push/simple flags, seq).
#pragma once
#include <atomic>
#include <chrono>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "message.h"
#include "van.h"

namespace hips {

class Resender {
 public:
  Resender(int timeout_ms, int max_retry, Van* van) : timeout_(timeout_ms), max_retry_(max_retry), van_(van) {
    monitor_ = new std::thread(&Resender::Monitoring, this);
  }
  ~Resender() {
    exit_ = true;
    monitor_->join();
    delete monitor_;
  }
  // block (bounded) until every outgoing message has been acknowledged — called before a node tears its Van down so that peers
  // still waiting for a retransmission (e.g. of the final barrier release) are served
  void WaitDrained(int max_ms) {
    for (int waited = 0; waited < max_ms; waited += 10) {
      { std::lock_guard<std::mutex> lk(mu_); if (send_buff_.empty()) return; }
      std::this_thread::sleep_for(Time(10));
    }
  }
  // remember an outgoing message until its ACK arrives
  void AddOutgoing(const Message& msg) {
    if (msg.meta.control.cmd == Control::ACK) return;
    uint64_t key = GetKey(msg);
    std::lock_guard<std::mutex> lk(mu_);
    if (send_buff_.find(key) != send_buff_.end()) return;
    Entry e; e.msg = msg; e.send = Now(); e.num_retry = 0;
    send_buff_[key] = e;
  }
  // returns true if the message is a duplicate or an ACK (i.e. must not be delivered)
  bool AddIncoming(const Message& msg) {
    if (msg.meta.control.cmd == Control::TERMINATE) return false;
    if (msg.meta.control.cmd == Control::ACK) {
      std::lock_guard<std::mutex> lk(mu_);
      send_buff_.erase(msg.meta.control.msg_sig);
      return true;
    }
    uint64_t key = GetKey(msg);
    bool duplicated;
    {
      std::lock_guard<std::mutex> lk(mu_);
      duplicated = !acked_.insert(key).second;
    }
    Message ack;
    ack.meta.recver = msg.meta.sender;
    ack.meta.sender = msg.meta.recver;
    ack.meta.control.cmd = Control::ACK;
    ack.meta.control.msg_sig = key;
    van_->SendNow(ack);
    return duplicated;
  }

 private:
  using Time = std::chrono::milliseconds;
  struct Entry { Message msg; Time send; int num_retry; };
  static Time Now() { return std::chrono::duration_cast<Time>(std::chrono::steady_clock::now().time_since_epoch()); }
  static uint64_t GetKey(const Message& msg) {
    const Meta& m = msg.meta;
    uint64_t h = 1469598103934665603ull;
    auto mix = [&h](uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(static_cast<uint64_t>(m.app_id)); mix(static_cast<uint64_t>(m.customer_id)); mix(static_cast<uint64_t>(m.timestamp));
    mix(static_cast<uint64_t>(m.sender)); mix(static_cast<uint64_t>(m.recver)); mix(m.request); mix(m.push); mix(m.simple_app);
    mix(static_cast<uint64_t>(m.control.cmd)); mix(static_cast<uint64_t>(m.seq)); mix(static_cast<uint64_t>(m.head + 7));
    return h;
  }
  void Monitoring() {
    while (!exit_) {
      std::this_thread::sleep_for(Time(timeout_ / 2 > 0 ? timeout_ / 2 : 1));
      std::vector<Message> resend;
      Time now = Now();
      {
        std::lock_guard<std::mutex> lk(mu_);
        for (auto& it : send_buff_) {
          if (it.second.send + Time(timeout_) * (1 + it.second.num_retry) < now) {
            resend.push_back(it.second.msg);
            ++it.second.num_retry;
            HIPS_CHECK_MSG(it.second.num_retry < max_retry_, "message resent too many times: peer is unreachable");
          }
        }
      }
      for (const auto& m : resend) van_->SendNow(m);
    }
  }
  int timeout_, max_retry_;
  Van* van_;
  std::thread* monitor_;
  std::atomic<bool> exit_{false};
  std::mutex mu_;
  std::unordered_map<uint64_t, Entry> send_buff_;
  std::unordered_set<uint64_t> acked_;
};

}  // namespace hips
