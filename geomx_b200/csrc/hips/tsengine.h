// TSEngine — scheduler-driven peer-merge (push) and relay-broadcast (pull) overlay.
//
// Parity: 3rdparty/ps-lite/src/van.cc:1174-1504.  ENABLE_INTRA_TS runs it among the workers of a party (local plane), ENABLE_INTER_TS
// among the local servers (global plane).
//   push : every sender merges its own tensor locally, then ASKPUSHes the scheduler.  The scheduler keeps a FIFO of askers and pairs
//          the two oldest: if one of them is the server it tells the other to send to the server; otherwise node a sends to b when the
//          recorded throughput A[a][b] > A[b][a], else b to a (ProcessAskPushCommand :1197-1252).  The receiver merges
//          (num_merge accumulates) and asks again, until the server holds NumWorkers merges.
//   pull : the holder of fresh parameters ASKPULLs (reporting the throughput it measured to its previous receiver); the scheduler
//          marks busy nodes in B and picks an idle receiver epsilon-greedily from the throughput matrix A — greedy with probability
//          min(known/(known+unknown), MAX_GREED_RATE_TS), random otherwise (:1312-1386) — or answers -1 once every worker has the
//          version.  Receivers relay onward themselves.
// This class is the scheduler side; the node side (merge buffers, relays) lives in kv_app.h (KVWorker / KVServer).
#pragma once
#include <algorithm>
#include <deque>
#include <map>
#include <mutex>
#include <random>
#include <set>
#include <vector>

#include "env.h"
#include "message.h"
#include "van.h"

namespace hips {

class TSScheduler {
 public:
  TSScheduler(Van* van, int num_workers, Plane plane)
      : van_(van), plane_(plane), num_workers_(num_workers), rng_(20240601u) {
    max_greed_ = static_cast<float>(Environment::Get()->GetFloat("MAX_GREED_RATE_TS", 0.9));
    server_id_ = ServerRankToID(0, plane);
  }

  void Process(const Message& msg) {
    if (msg.meta.control.cmd == Control::ASKPUSH) AskPush(msg);
    else if (msg.meta.control.cmd == Control::ASKPULL) AskPull(msg);
  }

  // exposed for unit tests
  int PickReceiver(int requester, const std::vector<int>& idle) {
    int known = 0, unknown = 0;
    for (int r : idle) (A_[requester].count(r) ? known : unknown)++;
    if (idle.empty()) return -1;
    float greed = std::min(known / static_cast<float>(known + unknown), max_greed_);
    std::uniform_real_distribution<float> u(0.f, 1.f);
    if (known > 0 && u(rng_) <= greed) {
      int best = -1; long bt = -1;
      for (int r : idle) { auto it = A_[requester].find(r); if (it != A_[requester].end() && it->second > bt) { bt = it->second; best = r; } }
      return best;
    }
    std::uniform_int_distribution<int> pick(0, static_cast<int>(idle.size()) - 1);
    return idle[pick(rng_)];
  }
  void Record(int from, int to, long throughput) { A_[from][to] = throughput; }

 private:
  void AskPush(const Message& msg) {
    std::lock_guard<std::mutex> lk(mu_);
    const int key = msg.meta.key;
    auto& q = ask_q_[key];
    const int sender = msg.meta.sender;
    if (q.size() == 1 && q.front().sender == sender) return;  // duplicate ask (van.cc:1198)
    q.push_back({sender, msg.meta.app_id, msg.meta.customer_id, msg.meta.timestamp});
    while (q.size() > 1) {
      Asker a = q.front(); q.pop_front();
      Asker b = q.front(); q.pop_front();
      Asker from, to;
      if (a.sender == server_id_) { from = b; to = a; }
      else if (b.sender == server_id_) { from = a; to = b; }
      else if (Throughput(a.sender, b.sender) > Throughput(b.sender, a.sender)) { from = a; to = b; }
      else { from = b; to = a; }
      Message rpl;
      rpl.meta.recver = from.sender;
      rpl.meta.app_id = from.app; rpl.meta.customer_id = from.customer; rpl.meta.timestamp = from.ts;
      rpl.meta.control.cmd = Control::REPLY;
      rpl.meta.push = true; rpl.meta.request = true;
      rpl.meta.key = key;
      rpl.meta.iters = to.sender;  // destination id travels in `iters`, as in the reference
      van_->SendNow(rpl);
    }
  }

  void AskPull(const Message& msg) {
    std::unique_lock<std::mutex> lk(mu_);
    const int req = msg.meta.sender, key = msg.meta.key, version = msg.meta.version;
    if (msg.meta.app_id != -1 && msg.meta.customer_id >= 0) A_[req][msg.meta.customer_id] = msg.meta.app_id;  // report of the last transfer
    auto& st = pull_[key];
    int recv = -1;
    if (version >= st.version) {   // a straggler still relaying an older version is simply told that everybody is served
      if (version > st.version) { st.version = version; st.served.clear(); }
      st.served.insert(req);  // a node holding the version never needs it again
      std::vector<int> idle;
      for (int r = 0; r < num_workers_; ++r) {
        const int id = WorkerRankToID(r, plane_);
        if (!st.served.count(id)) idle.push_back(id);
      }
      recv = PickReceiver(req, idle);
      if (recv >= 0) st.served.insert(recv);
    }
    lk.unlock();
    Message reply;
    reply.meta.recver = req;
    reply.meta.control.cmd = Control::REPLY;
    reply.meta.push = false; reply.meta.request = true;
    reply.meta.key = key; reply.meta.version = version;
    reply.meta.app_id = msg.meta.head;          // echo: owning app
    reply.meta.customer_id = msg.meta.body.empty() ? 0 : atoi(msg.meta.body.c_str());
    reply.meta.iters = recv;                     // -1 = everybody has this version
    reply.meta.timestamp = van_->GetTimestamp();
    van_->SendNow(reply);
  }

  long Throughput(int a, int b) { auto it = A_[a].find(b); return it == A_[a].end() ? -1 : it->second; }

  struct Asker { int sender, app, customer, ts; };
  struct PullState { int version = -1; std::set<int> served; };
  Van* van_;
  Plane plane_;
  int num_workers_, server_id_;
  float max_greed_;
  std::mt19937 rng_;
  std::mutex mu_;
  std::map<int, std::deque<Asker>> ask_q_;
  std::map<int, std::map<int, long>> A_;  // throughput matrix (bytes / ms), sparse
  std::map<int, PullState> pull_;
};

}  // namespace hips
