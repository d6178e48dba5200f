// Customer: per-application request tracker + receive threads.
// Parity: ps-lite include/ps/internal/customer.h + src/customer.cc:14-87 (NewRequest/WaitRequest/NumResponse/AddResponse, recv
// thread).  GeoMX gave servers a SECOND queue+thread for pull requests so pulls never queue behind pushes (customer.h:91-101); here
// the split is by message kind (pull request vs. everything else) instead of a hard-coded list of cmd heads.
#pragma once
#include <atomic>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <utility>
#include <vector>

#include "message.h"
#include "threadsafe_queue.h"

namespace hips {

class Postoffice;

class Customer {
 public:
  using RecvHandle = std::function<void(const Message&)>;
  Customer(int app_id, int customer_id, const RecvHandle& handle, bool dual_queue);
  ~Customer();
  int app_id() const { return app_id_; }
  int customer_id() const { return customer_id_; }
  // expects one response from every node of `recver` group in plane p
  int NewRequest(int recver, Plane p);
  int NewRequestCount(int num);
  void WaitRequest(int timestamp);
  int NumResponse(int timestamp);
  void AddResponse(int timestamp, int num = 1);
  void Accept(const Message& recved);
  // responses are handled on the caller's (the Van receive) thread instead of being queued for the customer thread: saves one thread
  // wake-up per response on the worker side, where response handlers only copy data and signal the waiter (never block, never wait for
  // other messages).  Requests still go through the queue.
  void set_inline_responses(bool on) { inline_responses_ = on; }
  // data requests (push / pull, not commands, not control) are handled on the Van receive thread as well: for servers whose request
  // handler only hands the message to a worker lane and never blocks
  void set_inline_requests(bool on) { inline_requests_ = on; }

 private:
  void CountResponse(const Message& recv);
  bool inline_responses_ = false;
  std::atomic<bool> inline_requests_{false};
  void Receiving(ThreadsafeQueue<Message, MessagePriority>* q);
  int app_id_, customer_id_;
  RecvHandle recv_handle_;
  ThreadsafeQueue<Message, MessagePriority> recv_queue_, pull_queue_;
  std::unique_ptr<std::thread> recv_thread_, pull_thread_;
  std::mutex tracker_mu_;
  std::condition_variable tracker_cond_;
  std::vector<std::pair<int, int>> tracker_;
};

}  // namespace hips
