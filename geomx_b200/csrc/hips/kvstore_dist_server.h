// KVStoreDistServer — one class, three roles: *standalone* server (single tier), *local* server of a party (global-plane client),
// *global* server (central party).
//
// Parity: src/kvstore/kvstore_dist_server.h — command channel (enums :49-56, CommandHandle :312-367), Cantor-paired data cmd (:82-104),
// DataHandleEx dispatch (:458-524), sync aggregation with deferred worker ack (DataHandleSyncDefault :1213-1366), async apply
// (DataHandleAsyncDefault :1519-1611), local->global push (DataPushToGlobalServers{Default,Compressed,BSCompressed} :758-897), push-ack ->
// pull (:941-957, :899-936), pull response + HFA milestone algebra (:959-972, :974-1169), storage responses (:1171-1211, :1705-1763),
// multi-precision master copies (:374-407, :526-553), key sharding across global servers (MultiGPS :1765-1906), P3 push-response-with-
// params (:1154-1164, :1257-1267), stop protocol (:315-328), initialized_ gate (:1719-1724), server profiler commands (:409-456).
//
// Concurrency: there is no server-wide lock and no handler ever blocks.  Every key owns its state and a mutex (KeyState); a small
// reader/writer registry maps key -> state; data requests (pushes AND pulls) are handed by the transport's receive thread straight to one of
// GEOMX_SERVER_LANES worker threads chosen by key (per-key FIFO order; different keys aggregate, run their optimizer and talk to the global
// tier in parallel; GEOMX_INLINE_REQUESTS=0 or P3 / TSEngine route them through the customer's priority queue first).  A pull that arrives
// before its key has a value is parked with the key and answered by whoever initialises it.  Responses from the global tier are handled on
// the receive thread of the connection they arrived on.  Replies are built under the key's lock and sent after it is released.
//
// Design differences: continuation-style state machine on raw byte buffers (no NDArray/engine on the server), native optimizers run
// in the receiving thread (server_optim.h) and only *foreign* (pickled Python) optimizers hop to the main thread through `Executor`,
// server-side state (optimizer moments, HFA milestones, BSC residuals) can be checkpointed (kSaveStates/kLoadStates) — the reference
// cannot ("Cannot save states for distributed training", python/mxnet/kvstore.py:578).
#pragma once
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <future>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <queue>
#include <shared_mutex>
#include <deque>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "gradient_compression.h"
#include "half.h"
#include "kv_app.h"
#include "../runtime/profiler.h"
#include "server_optim.h"

namespace hips {

enum class CommandType : int {
  kController = 0, kSetMultiPrecision = 1, kStopServer = 2, kSyncMode = 3, kSyncGlobalMode = 4, kSetGradientCompression = 5,
  kSetProfilerParams = 6, kSetOptimizerSpec = 7, kSaveStates = 8, kLoadStates = 9
};
enum class RequestType : int { kDefaultPushPull = 0, kRowSparsePushPull = 1, kCompressedPushPull = 2, kBSCompressedPushPull = 3 };

// mshadow dtype flags (3rdparty/mshadow/mshadow/base.h:302-310) + bf16 extension
enum DType : int { kFloat32 = 0, kFloat64 = 1, kFloat16 = 2, kUint8 = 3, kInt32 = 4, kInt8 = 5, kInt64 = 6, kBfloat16 = 12 };
inline int DTypeSize(int dt) {
  switch (dt) { case kFloat32: case kInt32: return 4; case kFloat64: case kInt64: return 8; case kFloat16: case kBfloat16: return 2; default: return 1; }
}

struct DataHandleType { RequestType requestType; int dtype; };
// Cantor pairing (reference :82-104)
inline int GetCommandType(RequestType t, int dtype) { const int m = static_cast<int>(t); return (((m + dtype) * (m + dtype + 1)) / 2) + dtype; }
inline DataHandleType DepairDataHandleType(int cmd) {
  const int w = static_cast<int>(std::floor((std::sqrt(8.0 * cmd + 1) - 1) / 2));
  const int t = ((w * w) + w) / 2;
  const int y = cmd - t, x = w - y;
  return DataHandleType{static_cast<RequestType>(x), y};
}

// main-thread executor for foreign (Python) updaters/controllers (reference :109-168)
class Executor {
 public:
  using Func = std::function<void()>;
  void Start() {
    std::unique_lock<std::mutex> lk(mu_);
    while (true) {
      cond_.wait(lk, [this] { return !queue_.empty(); });
      Block blk = std::move(queue_.front());
      queue_.pop();
      lk.unlock();
      if (blk.f) { blk.f(); blk.p->set_value(); }
      else { blk.p->set_value(); break; }
      lk.lock();
    }
  }
  void Exec(const Func& func) {
    Block blk(func);
    auto fut = blk.p->get_future();
    {
      std::lock_guard<std::mutex> lk(mu_);
      if (stopped_) return;          // the loop has ended (or is about to): nothing will ever run this block
      if (!func) stopped_ = true;
      queue_.push(std::move(blk)); cond_.notify_one();
    }
    fut.wait();
  }
  void Stop() { Exec(Func()); }      // idempotent

 private:
  struct Block {
    explicit Block(const Func& func) : f(func), p(std::make_shared<std::promise<void>>()) {}
    Func f;
    std::shared_ptr<std::promise<void>> p;
  };
  std::queue<Block> queue_;
  std::mutex mu_;
  std::condition_variable cond_;
  bool stopped_ = false;
};

class KVStoreDistServer {
 public:
  // foreign updater: (key, grad fp32*, weight fp32*, n) executed on the main thread; controller: (head, body)
  using Updater = std::function<void(int, const float*, float*, size_t)>;
  using Controller = std::function<void(int, const std::string&)>;

  KVStoreDistServer();
  ~KVStoreDistServer();
  void set_controller(const Controller& c) { controller_ = c; }
  void set_updater(const Updater& u) { updater_ = u; }
  void Run() { exec_.Start(); }            // blocks the (Python) main thread until kStopServer
  int rank_local();
  // introspection for tests / checkpoints
  std::vector<float> GetStored(int key);
  long num_pushes() const { return num_pushes_.load(); }

 private:
  struct Entry {
    std::vector<char> data;       // payload in the key's dtype (what pulls return)
    std::vector<float> master;    // fp32 master copy (multi-precision) or fp32 working copy for non-fp32 keys
    int dtype = kFloat32;
    size_t elems = 0;
    bool has_master = false;
  };
  struct UpdateBuf {
    std::vector<KVMeta> request;
    std::vector<float> merged;    // fp32 accumulation
    int count = 0;
  };
  struct GlobalRound {           // local server: one in-flight push/pull round per key
    std::vector<KVMeta> waiting; // worker push requests to ack when the round completes
    int push_ts = -1, pull_ts = -1;
    int parts_expected = 0;
    bool via_ts = false;         // pushed through the inter-party TSEngine overlay: the fresh value arrives by relay, not by pull
    std::vector<std::pair<Key, std::vector<char>>> parts;
    int cmd = 0;
  };

  struct ParkedPull { DataHandleType type; KVMeta req; KVPairs data; };
  // everything the server knows about one key; `mu` guards all of it except `version`
  struct KeyState {
    std::mutex mu;
    std::vector<ParkedPull> parked_pulls;   // pulls that arrived before the key had a value: answered by whoever initialises it
    Entry entry;
    UpdateBuf ub;
    GlobalRound round;
    std::vector<float> milestone, bsc_u, bsc_v, residual_2bit;
    NativeOptimizer::State opt;
    bool initialized = false;
    long local_rounds = 0;                  // completed local aggregation rounds (HFA: every K2-th one goes to the global tier)
    bool skip_init_push = false;            // resumed key: the next init push of the (re)started job must not overwrite it
    std::atomic<int> version{0};            // completed synchronisation rounds (TSEngine relay version, checkpoint cadence)
  };
  struct Reply { KVMeta to; KVPairs data; };
  // push lanes: requests of one key always run on the same worker thread, in arrival order
  class Lane {
   public:
    Lane() : th_([this] { Loop(); }) {}
    ~Lane() {
      { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
      cv_.notify_all();
      th_.join();
    }
    void Post(std::function<void()> f) {
      { std::lock_guard<std::mutex> lk(mu_); q_.push_back(std::move(f)); }
      cv_.notify_one();
    }
   private:
    void Loop() {
      std::unique_lock<std::mutex> lk(mu_);
      while (true) {
        cv_.wait(lk, [this] { return stop_ || !q_.empty(); });
        if (q_.empty()) return;             // stop requested and everything queued has run
        std::function<void()> f = std::move(q_.front());
        q_.pop_front();
        lk.unlock();
        f();
        lk.lock();
      }
    }
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<std::function<void()>> q_;
    bool stop_ = false;
    std::thread th_;                        // last member: starts after the queue exists
  };

  KeyState& Slot(int key);                                   // find or create
  KeyState* Find(int key);
  std::vector<std::pair<int, KeyState*>> Slots();            // snapshot of the registry, ordered by key
  // handlers
  void CommandHandle(const SimpleData& recved, SimpleApp* app);
  void DataHandleEx(const KVMeta& req, const KVPairs& data, KVServer* server);
  void ResponseHandle(const KVMeta& res, const KVPairs& data, KVServer* server);
  void HandlePush(const DataHandleType& type, const KVMeta& req, const KVPairs& data);
  void HandlePull(const DataHandleType& type, const KVMeta& req, const KVPairs& data);
  // the functions below taking a KeyState expect its mutex to be held by the caller
  bool FinishLocalAggregation(int key, KeyState* ks, const DataHandleType& type, std::vector<Reply>* out);   // true: the round ended locally (HFA)
  void PushToGlobal(int key, KeyState* ks, const DataHandleType& type);
  void PullFromGlobal(int key, const DataHandleType& type);                                                  // takes the key's lock itself
  void ApplyUpdate(int key, KeyState* ks, const float* grad, size_t n);
  void ApplyFreshFromGlobal(int key, KeyState* ks, std::vector<float>* recved, std::vector<Reply>* out);
  bool BumpRound(KeyState* ks);                               // true: a periodic checkpoint is due (write it after releasing the key)
  Reply StoredReply(const KVMeta& to, int key, const Entry& e) const;
  Reply PullReply(KeyState* ks, const DataHandleType& type, const KVMeta& req, const KVPairs& data);
  void FlushParkedPulls(KeyState* ks, std::vector<Reply>* out);
  void Send(std::vector<Reply>* out);
  void StoreFromFloat(Entry* e, const float* src, size_t n);
  void ToFloat(const char* src, int dtype, size_t n, float* dst);
  // TSEngine (ts_node.h): a merged push stands for several requests; servers take part in the pairing and start the relay
  std::vector<KVMeta> ExpandOrigins(const KVMeta& req);
  void AskTS(int key);
  void RoundCompleted(int key, bool bumped = false);
  void OnRelayedFromGlobal(int key, int version, int cmd, const std::vector<char>& bytes);
  void SaveStates(const std::string& path);
  void LoadStates(const std::string& path);

  std::unique_ptr<KVServer> ps_server_;
  Executor exec_;
  Controller controller_;
  Updater updater_;
  std::shared_ptr<const NativeOptimizer> native_opt_;     // replaced as a whole by kSetOptimizerSpec (atomic_load / atomic_store)
  GradientCompression gc_;
  std::shared_mutex reg_mu_;                              // the registry only: never held while a key's mutex is taken
  std::map<int, std::unique_ptr<KeyState>> keys_;
  std::vector<std::unique_ptr<Lane>> lanes_;
  std::mutex ctl_mu_;                                     // stop votes
  std::mutex round_mu_;                                   // checkpoint cadence (ckpt_key_)
  std::atomic<bool> sync_mode_{false}, sync_global_mode_{false}, multi_precision_{false}, any_skip_init_{false};
  bool is_global_ = false, has_global_ = false, standalone_ = false;
  bool use_hfa_ = false;
  int hfa_k2_ = 1;
  std::atomic<long> local_iters_{0};
  size_t bigarray_bound_ = 1000000, size_lower_bound_ = 200000;
  int stop_votes_ = 0;
  bool stop_requested_ = false;
  std::atomic<long> num_pushes_{0};
  // periodic server-state checkpoints + resume (GEOMX_SERVER_CKPT_PREFIX / _EVERY / GEOMX_SERVER_RESUME), see BumpRound / TryResume
  std::string ckpt_prefix_;
  int ckpt_every_ = 0, ckpt_key_ = 0;                  // ckpt_key_: round number of the last snapshot
  bool fused_tier_pull_ = true;                          // global server: dense push responses to local servers carry the fresh value
  bool resume_wanted_ = false;
  std::once_flag resume_once_;
  void TryResume();
  std::string StatePath(const std::string& prefix) const;
};

}  // namespace hips
