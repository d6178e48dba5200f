#include "customer.h"

#include "postoffice.h"

namespace hips {

Customer::Customer(int app_id, int customer_id, const RecvHandle& handle, bool dual_queue)
    : app_id_(app_id), customer_id_(customer_id), recv_handle_(handle) {
  Postoffice::Get()->AddCustomer(this);
  recv_thread_.reset(new std::thread(&Customer::Receiving, this, &recv_queue_));
  if (dual_queue) pull_thread_.reset(new std::thread(&Customer::Receiving, this, &pull_queue_));
}

Customer::~Customer() {
  Postoffice::Get()->RemoveCustomer(this);
  Message exit;
  exit.meta.control.cmd = Control::TERMINATE;
  exit.meta.priority = -(1 << 30);
  recv_queue_.Push(exit);
  recv_thread_->join();
  if (pull_thread_) { pull_queue_.Push(exit); pull_thread_->join(); }
}

int Customer::NewRequest(int recver, Plane p) {
  const int num = static_cast<int>(Postoffice::Get()->GetNodeIDs(recver, p).size());
  return NewRequestCount(num);
}

int Customer::NewRequestCount(int num) {
  std::lock_guard<std::mutex> lk(tracker_mu_);
  tracker_.push_back(std::make_pair(num, 0));
  return static_cast<int>(tracker_.size()) - 1;
}

void Customer::WaitRequest(int timestamp) {
  std::unique_lock<std::mutex> lk(tracker_mu_);
  tracker_cond_.wait(lk, [this, timestamp] { return tracker_[timestamp].first == tracker_[timestamp].second; });
}

int Customer::NumResponse(int timestamp) {
  std::lock_guard<std::mutex> lk(tracker_mu_);
  return tracker_[timestamp].second;
}

void Customer::AddResponse(int timestamp, int num) {
  std::lock_guard<std::mutex> lk(tracker_mu_);
  tracker_[timestamp].second += num;
  tracker_cond_.notify_all();
}

void Customer::CountResponse(const Message& recv) {
  if (!recv.meta.request && recv.meta.control.empty()) {
    std::lock_guard<std::mutex> lk(tracker_mu_);
    if (recv.meta.timestamp >= 0 && static_cast<size_t>(recv.meta.timestamp) < tracker_.size()) {
      tracker_[recv.meta.timestamp].second++;
      tracker_cond_.notify_all();
    }
  }
}

void Customer::Accept(const Message& recved) {
  if (inline_responses_ && !recved.meta.request && recved.meta.control.empty()) {
    recv_handle_(recved);
    CountResponse(recved);
    return;
  }
  if (inline_requests_.load() && recved.meta.request && !recved.meta.simple_app && recved.meta.control.empty()) {
    recv_handle_(recved);
    return;
  }
  // pull requests get their own queue/thread on servers so that they never wait behind pushes (reference customer.h:91-101)
  const bool is_pull_request = recved.meta.request && !recved.meta.push && !recved.meta.simple_app && recved.meta.control.empty();
  if (pull_thread_ && is_pull_request) pull_queue_.Push(recved);
  else recv_queue_.Push(recved);
}

void Customer::Receiving(ThreadsafeQueue<Message, MessagePriority>* q) {
  while (true) {
    Message recv;
    q->WaitAndPop(&recv);
    if (!recv.meta.control.empty() && recv.meta.control.cmd == Control::TERMINATE) break;
    recv_handle_(recv);
    CountResponse(recv);
  }
}

}  // namespace hips
