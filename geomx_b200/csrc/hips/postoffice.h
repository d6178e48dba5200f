// Postoffice: process singleton — roles, counts, id groups, both Vans, customers, barriers, key ranges, liveness.
// Parity: ps-lite include/ps/internal/postoffice.h:18-233 + src/postoffice.cc (InitEnvironment :18-58, Start :60-111, StartGlobal
// :113-148, Barrier :202-244, GetServerKeyRanges :246-259, GetDeadNodes :284-303).
#pragma once
#include <atomic>
#include <condition_variable>
#include <ctime>
#include <map>
#include <memory>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "env.h"
#include "message.h"
#include "van.h"

namespace hips {

class Customer;

class Postoffice {
 public:
  static Postoffice* Get() { static Postoffice p; return &p; }
  // (re)reads the environment; safe to call once per process
  void InitEnvironment();
  void Start(int customer_id, bool do_barrier);   // local plane (if this process has a local role) + global plane (if it has a global role)
  void Finalize(int customer_id, bool do_barrier);
  Van* van(Plane p) { return p == kLocal ? van_local_.get() : van_global_.get(); }
  bool has_plane(Plane p) const { return p == kLocal ? has_local_ : has_global_; }

  // ---- role predicates (reference include/mxnet/kvstore.h:282-358) -------------------------------------------------
  bool is_worker() const { return is_worker_; }
  bool is_server() const { return is_server_; }
  bool is_scheduler() const { return is_scheduler_; }
  bool is_global_server() const { return is_global_server_; }
  bool is_global_scheduler() const { return is_global_scheduler_; }
  bool is_master_worker() const { return is_master_worker_; }
  // set once the applications of this process are being torn down: late (re)transmissions for them are dropped instead of waited for
  void set_finalizing() { finalizing_ = true; }
  bool is_finalizing() const { return finalizing_.load(); }
  // this process took over the id of a dead node (van.cc recovery path): it must not wait at start-up barriers the others passed long ago
  bool is_recovery() { return (has_local_ && van_local_ && van_local_->my_node().is_recovery) || (has_global_ && van_global_ && van_global_->my_node().is_recovery); }
  bool enable_central_workers() const { return enable_central_worker_; }
  int num_workers() const { return num_workers_; }
  int num_servers() const { return num_servers_; }
  int num_global_workers() const { return num_global_workers_; }
  int num_global_servers() const { return num_global_servers_; }
  int num_all_workers() const { return num_all_workers_; }
  // role of this process inside plane p
  int role_in(Plane p) const;
  int num_workers_in(Plane p) const { return p == kLocal ? num_workers_ : num_global_workers_; }
  int num_servers_in(Plane p) const { return p == kLocal ? num_servers_ : num_global_servers_; }
  int my_rank(Plane p) { return IDtoRank(van(p)->my_node().id, p); }
  const std::vector<int>& GetNodeIDs(int group, Plane p) { return node_ids_[p].at(group); }

  // ---- customers ---------------------------------------------------------------------------------------------------
  void AddCustomer(Customer* c);
  void RemoveCustomer(Customer* c);
  Customer* GetCustomer(int app_id, int customer_id, int timeout_sec = 0);
  // look the customer up and hand it `msg` while the registry lock is held, so that a customer being destroyed on another thread cannot be
  // dereferenced after its removal; false = no such customer (yet / any more)
  bool DeliverTo(int app_id, int customer_id, const Message& msg, int timeout_sec);

  // ---- barrier / manage ----------------------------------------------------------------------------------------------
  void Barrier(int customer_id, int node_group, Plane p);
  void Manage(const Message& recv, Plane p);   // called by the Van for BARRIER responses
  const std::vector<Range>& GetServerKeyRanges(Plane p);

  // ---- liveness ------------------------------------------------------------------------------------------------------
  void UpdateHeartbeat(int node_id, time_t t, Plane p) { std::lock_guard<std::mutex> lk(hb_mu_); heartbeats_[p][node_id] = t; }
  std::vector<int> GetDeadNodes(int timeout_sec, Plane p);
  time_t start_time() const { return start_time_; }
  bool started() const { return started_; }

 private:
  Postoffice() {}
  void BuildGroups(Plane p);
  std::unique_ptr<Van> van_local_, van_global_;
  bool has_local_ = false, has_global_ = false, started_ = false;
  bool is_worker_ = false, is_server_ = false, is_scheduler_ = false, is_global_server_ = false, is_global_scheduler_ = false;
  bool is_master_worker_ = false, enable_central_worker_ = false;
  std::atomic<bool> finalizing_{false};
  int num_workers_ = 0, num_servers_ = 0, num_global_workers_ = 0, num_global_servers_ = 0, num_all_workers_ = 0;
  std::unordered_map<int, std::vector<int>> node_ids_[2];
  std::vector<Range> key_ranges_[2];
  std::mutex mu_, barrier_mu_, hb_mu_, start_mu_;
  std::condition_variable barrier_cv_;
  std::unordered_map<int, std::unordered_map<int, bool>> barrier_done_[2];  // [plane][app][customer]
  std::unordered_map<int, std::unordered_map<int, Customer*>> customers_;
  std::unordered_map<int, time_t> heartbeats_[2];
  time_t start_time_ = 0;
  int init_stage_ = 0;
};

}  // namespace hips
