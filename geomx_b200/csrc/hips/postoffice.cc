#include "postoffice.h"

#include <chrono>
#include <thread>

#include "customer.h"

namespace hips {

void Postoffice::InitEnvironment() {
  Environment* e = Environment::Get();
  const std::string role = e->GetStr("DMLC_ROLE", "");
  const std::string grole = e->GetStr("DMLC_ROLE_GLOBAL", "");
  is_worker_ = role == "worker";
  is_server_ = role == "server";
  is_scheduler_ = role == "scheduler";
  is_global_server_ = grole == "global_server";
  is_global_scheduler_ = grole == "global_scheduler";
  is_master_worker_ = e->GetInt("DMLC_ROLE_MASTER_WORKER", 0) != 0;
  enable_central_worker_ = e->GetInt("DMLC_ENABLE_CENTRAL_WORKER", 0) != 0;
  num_workers_ = e->GetInt("DMLC_NUM_WORKER", 0);
  num_servers_ = e->GetInt("DMLC_NUM_SERVER", 0);
  num_global_workers_ = e->GetInt("DMLC_NUM_GLOBAL_WORKER", 0);
  num_global_servers_ = e->GetInt("DMLC_NUM_GLOBAL_SERVER", 0);
  num_all_workers_ = e->GetInt("DMLC_NUM_ALL_WORKER", num_workers_);
  has_local_ = is_worker_ || is_server_ || is_scheduler_;
  // a server takes part in the global plane when a global scheduler address is configured (local server = "global worker")
  const bool has_global_addr = e->find("DMLC_PS_GLOBAL_ROOT_URI") != nullptr && num_global_servers_ > 0;
  has_global_ = is_global_scheduler_ || is_global_server_ || (is_server_ && has_global_addr);
  // reference postoffice.cc:55-57: a non-central party has exactly one local server
  if (is_server_ && !is_global_server_ && has_global_) HIPS_CHECK_MSG(num_servers_ <= 1, "only one local server per party is supported");
}

int Postoffice::role_in(Plane p) const {
  if (p == kLocal) return is_scheduler_ ? Node::SCHEDULER : (is_server_ ? Node::SERVER : Node::WORKER);
  return is_global_scheduler_ ? Node::SCHEDULER : (is_global_server_ ? Node::SERVER : Node::WORKER);
}

void Postoffice::BuildGroups(Plane p) {
  auto& ids = node_ids_[p];
  ids.clear();
  const int nw = num_workers_in(p), ns = num_servers_in(p);
  for (int i = 0; i < nw; ++i) {
    const int id = WorkerRankToID(i, p);
    for (int g : {id, kWorkerGroup, kWorkerGroup + kServerGroup, kWorkerGroup + kScheduler, kWorkerGroup + kServerGroup + kScheduler}) ids[g].push_back(id);
  }
  for (int i = 0; i < ns; ++i) {
    const int id = ServerRankToID(i, p);
    for (int g : {id, kServerGroup, kWorkerGroup + kServerGroup, kServerGroup + kScheduler, kWorkerGroup + kServerGroup + kScheduler}) ids[g].push_back(id);
  }
  for (int g : {kScheduler, kScheduler + kServerGroup + kWorkerGroup, kScheduler + kWorkerGroup, kScheduler + kServerGroup}) ids[g].push_back(kScheduler);
  key_ranges_[p].clear();
  for (int i = 0; i < ns; ++i) key_ranges_[p].push_back(Range(kMaxKey / ns * i, kMaxKey / ns * (i + 1)));
}

void Postoffice::Start(int customer_id, bool do_barrier) {
  std::lock_guard<std::mutex> lk(start_mu_);
  if (init_stage_ == 0) {
    InitEnvironment();
    if (has_local_) BuildGroups(kLocal);
    if (has_global_) BuildGroups(kGlobal);
    if (has_local_) van_local_.reset(new Van(this, kLocal));
    if (has_global_) van_global_.reset(new Van(this, kGlobal));
    start_time_ = time(nullptr);
    init_stage_ = 1;
  }
  if (init_stage_ == 1) {
    if (is_global_server_ && has_local_) {
      // a global server joins the global plane first and asks the central party's scheduler for the SAME server rank, so that the
      // key -> server hash used by the master worker (local plane) and by the local servers (global plane) hits the same process
      van_global_->Start(customer_id);
      van_local_->set_rank_hint(IDtoRank(van_global_->my_node().id, kGlobal));
      van_local_->Start(customer_id);
    } else {
      if (has_local_) van_local_->Start(customer_id);
      if (has_global_) van_global_->Start(customer_id);
    }
    init_stage_ = 2;
    started_ = true;
  }
  if (do_barrier && !is_recovery()) {   // a recovered node joins a running job (reference: kvstore_dist.h:63 `if (!ps::Postoffice::Get()->is_recovery())`)
    if (has_local_) Barrier(customer_id, kWorkerGroup + kServerGroup + kScheduler, kLocal);
    if (has_global_) Barrier(customer_id, kWorkerGroup + kServerGroup + kScheduler, kGlobal);
  }
}

void Postoffice::Finalize(int customer_id, bool do_barrier) {
  if (!started_) return;
  if (do_barrier) {
    if (has_local_) Barrier(customer_id, kWorkerGroup + kServerGroup + kScheduler, kLocal);
    if (has_global_) Barrier(customer_id, kWorkerGroup + kServerGroup + kScheduler, kGlobal);
  }
  if (has_global_) van_global_->Stop();
  if (has_local_) van_local_->Stop();
  started_ = false;
  init_stage_ = 0;
  van_local_.reset(); van_global_.reset();
  { std::lock_guard<std::mutex> lk(barrier_mu_); barrier_done_[0].clear(); barrier_done_[1].clear(); }
}

void Postoffice::AddCustomer(Customer* c) {
  std::lock_guard<std::mutex> lk(mu_);
  customers_[c->app_id()][c->customer_id()] = c;
  std::lock_guard<std::mutex> bl(barrier_mu_);
  barrier_done_[0][c->app_id()][c->customer_id()] = false;
  barrier_done_[1][c->app_id()][c->customer_id()] = false;
}

void Postoffice::RemoveCustomer(Customer* c) {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = customers_.find(c->app_id());
  if (it != customers_.end()) it->second.erase(c->customer_id());
}

Customer* Postoffice::GetCustomer(int app_id, int customer_id, int timeout_sec) {
  for (int i = 0; i <= timeout_sec * 1000; ++i) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      auto it = customers_.find(app_id);
      if (it != customers_.end()) {
        auto jt = it->second.find(customer_id);
        if (jt != it->second.end()) return jt->second;
        if (!it->second.empty() && customer_id == Meta::kEmpty) return it->second.begin()->second;
      }
    }
    std::this_thread::sleep_for(std::chrono::milliseconds(1));
  }
  return nullptr;
}

bool Postoffice::DeliverTo(int app_id, int customer_id, const Message& msg, int timeout_sec) {
  for (int i = 0; i <= timeout_sec * 1000; ++i) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      auto it = customers_.find(app_id);
      if (it != customers_.end()) {
        auto jt = it->second.find(customer_id);
        Customer* c = jt != it->second.end() ? jt->second : ((!it->second.empty() && customer_id == Meta::kEmpty) ? it->second.begin()->second : nullptr);
        if (c != nullptr) { c->Accept(msg); return true; }     // Accept only enqueues
      }
    }
    if (finalizing_.load()) return false;
    std::this_thread::sleep_for(std::chrono::milliseconds(1));
  }
  return false;
}

void Postoffice::Barrier(int customer_id, int node_group, Plane p) {
  if (GetNodeIDs(node_group, p).size() <= 1) return;
  const int role = role_in(p);
  if (role == Node::SCHEDULER) HIPS_CHECK(node_group & kScheduler);
  else if (role == Node::WORKER) HIPS_CHECK(node_group & kWorkerGroup);
  else HIPS_CHECK(node_group & kServerGroup);
  std::unique_lock<std::mutex> ulk(barrier_mu_);
  barrier_done_[p][0][customer_id] = false;
  ulk.unlock();
  Message req;
  req.meta.recver = kScheduler;
  req.meta.request = true;
  req.meta.control.cmd = Control::BARRIER;
  req.meta.app_id = 0;
  req.meta.customer_id = customer_id;
  req.meta.control.barrier_group = node_group;
  req.meta.timestamp = van(p)->GetTimestamp();
  if (role == Node::SCHEDULER) {
    // the scheduler counts itself directly (no self-connection needed)
    Message self = req;
    self.meta.sender = kScheduler;
    // deliver through the normal path by sending to our own listening socket
    van(p)->SendNow(req);
  } else {
    van(p)->SendNow(req);
  }
  ulk.lock();
  barrier_cv_.wait(ulk, [this, p, customer_id] { return barrier_done_[p][0][customer_id]; });
}

void Postoffice::Manage(const Message& recv, Plane p) {
  const auto& ctrl = recv.meta.control;
  if (ctrl.cmd == Control::BARRIER && !recv.meta.request) {
    std::lock_guard<std::mutex> lk(barrier_mu_);
    for (auto& kv : barrier_done_[p][recv.meta.app_id]) kv.second = true;
    barrier_done_[p][0][recv.meta.customer_id] = true;
    barrier_cv_.notify_all();
  }
}

const std::vector<Range>& Postoffice::GetServerKeyRanges(Plane p) { return key_ranges_[p]; }

std::vector<int> Postoffice::GetDeadNodes(int t, Plane p) {
  std::vector<int> dead;
  if (!started_ && init_stage_ < 1) return dead;
  if (t <= 0) return dead;
  const time_t now = time(nullptr);
  std::lock_guard<std::mutex> lk(hb_mu_);
  const bool sched = role_in(p) == Node::SCHEDULER;
  std::vector<int> ids = sched ? GetNodeIDs(kWorkerGroup + kServerGroup, p) : GetNodeIDs(kScheduler, p);
  for (int r : ids) {
    auto it = heartbeats_[p].find(r);
    if ((it == heartbeats_[p].end() || it->second + t < now) && start_time_ + t < now) dead.push_back(r);
  }
  return dead;
}

}  // namespace hips
