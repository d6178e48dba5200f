// Application layer: SimpleApp (head/body RPC), KVWorker (push/pull client), KVServer (request handler + global-tier client).
// Parity: ps-lite include/ps/simple_app.h:131-167 and include/ps/kv_app.h (KVWorker::ZPush :171-202, P3_ZPush :204-259, ZPull
// :303-309, Send/DefaultSlicer :679-839, Process :1087-1109; KVServer::Push/Pull to global servers :480-512, Response :657-676,
// Process with dual handles :1227-1307).  Values are opaque bytes (the dtype travels in `cmd`, as in the reference's KVWorker<char>).
#pragma once
#include <algorithm>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "customer.h"
#include "dgt.h"
#include "postoffice.h"
#include "ts_node.h"

namespace hips {

struct KVPairs {
  SArray<Key> keys;
  SArray<char> vals;
  SArray<int> lens;
};

struct KVMeta {
  int cmd = 0;
  bool push = false;
  int sender = 0;
  int timestamp = 0;
  int customer_id = 0;
  int priority = 0;
  int key = 0, version = 0, num_merge = 1;
  int plane = kLocal;
  int app_id = 0;
  std::string body;   // TSEngine: origins of a merged push (ts_node.h EncodeOrigins)
};

struct SimpleData {
  int head = 0;
  std::string body;
  int sender = 0, timestamp = 0, customer_id = 0, plane = kLocal;
};

// ------------------------------------------------------------------------------------------------ SimpleApp
class SimpleApp {
 public:
  using Handle = std::function<void(const SimpleData&, SimpleApp*)>;
  SimpleApp(int app_id, int customer_id) {
    request_handle_ = [](const SimpleData& d, SimpleApp* app) { app->Response(d); };
    response_handle_ = [](const SimpleData&, SimpleApp*) {};
    obj_.reset(new Customer(app_id, customer_id, [this](const Message& m) { Process(m); }, false));
  }
  virtual ~SimpleApp() {}
  // send (head, body) to every node of `recv_group` in plane p; returns the timestamp to Wait on
  int Request(int head, const std::string& body, int recv_group, Plane p = kLocal) {
    Message msg;
    msg.meta.head = head;
    msg.meta.body = body;
    msg.meta.timestamp = obj_->NewRequest(recv_group, p);
    msg.meta.request = true;
    msg.meta.simple_app = true;
    msg.meta.app_id = obj_->app_id();
    msg.meta.customer_id = obj_->customer_id();
    for (int r : Postoffice::Get()->GetNodeIDs(recv_group, p)) {
      msg.meta.recver = r;
      Postoffice::Get()->van(p)->Send(msg);
    }
    return msg.meta.timestamp;
  }
  void Wait(int timestamp) { obj_->WaitRequest(timestamp); }
  void Response(const SimpleData& req, const std::string& res_body = "") {
    Message msg;
    msg.meta.head = req.head;
    msg.meta.body = res_body;
    msg.meta.timestamp = req.timestamp;
    msg.meta.request = false;
    msg.meta.simple_app = true;
    msg.meta.app_id = obj_->app_id();
    msg.meta.customer_id = req.customer_id;
    msg.meta.recver = req.sender;
    Postoffice::Get()->van(static_cast<Plane>(req.plane))->Send(msg);
  }
  void set_request_handle(const Handle& h) { request_handle_ = h; }
  void set_response_handle(const Handle& h) { response_handle_ = h; }
  Customer* get_customer() { return obj_.get(); }

 protected:
  SimpleApp() {
    request_handle_ = [](const SimpleData& d, SimpleApp* app) { app->Response(d); };
    response_handle_ = [](const SimpleData&, SimpleApp*) {};
  }
  void ProcessSimple(const Message& msg) {
    SimpleData d;
    d.sender = msg.meta.sender; d.head = msg.meta.head; d.body = msg.meta.body; d.timestamp = msg.meta.timestamp;
    d.customer_id = msg.meta.customer_id; d.plane = msg.meta.plane;
    if (msg.meta.request) request_handle_(d, this);
    else response_handle_(d, this);
  }
  virtual void Process(const Message& msg) { ProcessSimple(msg); }
  std::unique_ptr<Customer> obj_;
  Handle request_handle_, response_handle_;
};

// slice (keys sorted ascending) by server key ranges
inline void DefaultSlicer(const KVPairs& send, const std::vector<Range>& ranges, std::vector<std::pair<bool, KVPairs>>* sliced) {
  sliced->resize(ranges.size());
  const size_t n = ranges.size();
  std::vector<size_t> pos(n + 1);
  const Key* begin = send.keys.begin();
  const Key* end = send.keys.end();
  for (size_t i = 0; i < n; ++i) {
    if (i == 0) pos[0] = std::lower_bound(begin, end, ranges[0].begin()) - begin;
    size_t len = std::lower_bound(begin + pos[i], end, ranges[i].end()) - begin - pos[i];
    if (i == n - 1) len = (end - begin) - pos[i];  // kMaxKey/n truncation: the last range takes the remainder
    pos[i + 1] = pos[i] + len;
    sliced->at(i).first = len != 0;
  }
  if (send.keys.empty()) return;
  const size_t k = send.vals.size() / std::max<size_t>(1, send.keys.size());
  size_t val_begin = 0;
  for (size_t i = 0; i < n; ++i) {
    if (pos[i + 1] == pos[i]) { sliced->at(i).first = false; continue; }
    auto& kv = sliced->at(i).second;
    kv.keys = send.keys.segment(pos[i], pos[i + 1]);
    if (send.lens.size()) {
      kv.lens = send.lens.segment(pos[i], pos[i + 1]);
      size_t val_end = val_begin;
      for (int l : kv.lens) val_end += l;
      kv.vals = send.vals.segment(val_begin, val_end);
      val_begin = val_end;
    } else {
      kv.vals = send.vals.segment(pos[i] * k, pos[i + 1] * k);
    }
  }
}

inline void FillDataMessage(Message* msg, const KVPairs& kvs) {
  msg->data.clear();
  if (kvs.keys.size()) {
    msg->AddData(kvs.keys);
    msg->AddData(kvs.vals);
    if (kvs.lens.size()) msg->AddData(kvs.lens);
  }
}
inline void ExtractKVPairs(const Message& msg, KVPairs* kvs) {
  if (msg.data.size() >= 2) {
    kvs->keys = msg.data[0];
    kvs->vals = msg.data[1];
    if (msg.data.size() > 2 && msg.data[2].size()) kvs->lens = msg.data[2];
  }
}

// ------------------------------------------------------------------------------------------------ KVWorker
class KVWorker : public SimpleApp {
 public:
  using Callback = std::function<void()>;
  KVWorker(int app_id, int customer_id) : SimpleApp() {
    // intra-party TSEngine: pushes are merged peer-to-peer and fresh parameters arrive through the relay (ts_node.h)
    if (Environment::Get()->GetInt("ENABLE_INTRA_TS", 0) != 0 && Postoffice::Get()->num_workers() > 1 &&
        Environment::Get()->GetInt("ENABLE_P3", 0) == 0)
      ts_.reset(new TSNode(kLocal, app_id, customer_id));
    obj_.reset(new Customer(app_id, customer_id, [this](const Message& m) { Process(m); }, false));
    // without the TSEngine overlay a worker's response handlers are copy-and-signal only: run them on the receive thread
    if (!ts_ && Environment::Get()->GetInt("GEOMX_INLINE_RESPONSES", 1) != 0) obj_->set_inline_responses(true);
  }
  TSNode* ts() { return ts_.get(); }
  ~KVWorker() override { obj_.reset(); }

  // zero-copy push: the caller keeps keys/vals/lens alive until the callback / Wait returns
  int ZPush(const SArray<Key>& keys, const SArray<char>& vals, const SArray<int>& lens, int cmd = 0, const Callback& cb = nullptr,
            int priority = 0, int int_key = 0, bool allow_ts = false) {
    const int ts = obj_->NewRequest(kServerGroup, kLocal);
    AddCallback(ts, cb);
    if (ts_ && allow_ts && keys.size() == 1 && Postoffice::Get()->num_servers() == 1) {
      // the server acknowledges THIS request (origin) once the round completes, wherever the gradient was merged on its way
      ts_->Offer(int_key, cmd, keys[0], vals.data(), vals.size(),
                 TSOrigin{Postoffice::Get()->van(kLocal)->my_node().id, ts, obj_->customer_id()});
      return ts;
    }
    KVPairs kvs; kvs.keys = keys; kvs.vals = vals; kvs.lens = lens;
    Send(ts, true, cmd, kvs, priority, int_key, nullptr);
    return ts;
  }
  // P3: the push response carries the updated values, copied into `vals` (kv_app.h:204-259)
  int P3_ZPush(const SArray<Key>& keys, const SArray<char>& vals, const SArray<int>& lens, int cmd, const Callback& cb, int priority, int int_key) {
    const int ts = obj_->NewRequest(kServerGroup, kLocal);
    { std::lock_guard<std::mutex> lk(mu_); p3_targets_[ts] = vals; }
    AddCallback(ts, cb);
    KVPairs kvs; kvs.keys = keys; kvs.vals = vals; kvs.lens = lens;
    Send(ts, true, cmd, kvs, priority, int_key, nullptr);
    return ts;
  }
  // `request_payload` (optional) travels with the pull request — row_sparse pulls send the wanted row ids this way
  int ZPull(const SArray<Key>& keys, SArray<char>* vals, SArray<int>* lens = nullptr, int cmd = 0, const Callback& cb = nullptr, int priority = 0,
            int int_key = 0, const SArray<char>* request_payload = nullptr) {
    const int ts = obj_->NewRequest(kServerGroup, kLocal);
    { std::lock_guard<std::mutex> lk(mu_); pull_targets_[ts] = PullTarget{keys, vals, lens}; }
    AddCallback(ts, cb);
    KVPairs kvs; kvs.keys = keys;
    if (request_payload != nullptr && keys.size() == 1) { kvs.vals = *request_payload; kvs.lens.push_back(static_cast<int>(request_payload->size())); }
    Send(ts, false, cmd, kvs, priority, int_key, nullptr);
    return ts;
  }
  void Wait(int timestamp) { obj_->WaitRequest(timestamp); }

 private:
  struct PullTarget { SArray<Key> keys; SArray<char>* vals; SArray<int>* lens; };
  void AddCallback(int ts, const Callback& cb) {
    if (!cb) return;
    std::lock_guard<std::mutex> lk(mu_);
    callbacks_[ts] = cb;
  }
  void RunCallback(int ts) {
    Callback cb;
    {
      std::lock_guard<std::mutex> lk(mu_);
      auto it = callbacks_.find(ts);
      if (it == callbacks_.end()) return;
      cb = it->second;
      callbacks_.erase(it);
    }
    cb();
  }
  void Send(int ts, bool push, int cmd, const KVPairs& kvs, int priority, int int_key, const int* only_server) {
    std::vector<std::pair<bool, KVPairs>> sliced;
    DefaultSlicer(kvs, Postoffice::Get()->GetServerKeyRanges(kLocal), &sliced);
    int skipped = 0;
    for (auto& s : sliced) if (!s.first) ++skipped;
    obj_->AddResponse(ts, skipped);
    if (static_cast<size_t>(skipped) == sliced.size()) RunCallback(ts);
    for (size_t i = 0; i < sliced.size(); ++i) {
      if (!sliced[i].first) continue;
      Message msg;
      msg.meta.app_id = obj_->app_id();
      msg.meta.customer_id = obj_->customer_id();
      msg.meta.request = true;
      msg.meta.push = push;
      msg.meta.head = cmd;
      msg.meta.timestamp = ts;
      msg.meta.recver = ServerRankToID(static_cast<int>(i), kLocal);
      msg.meta.priority = priority;
      msg.meta.key = int_key;
      msg.meta.iters = 1;
      FillDataMessage(&msg, sliced[i].second);
      Postoffice::Get()->van(kLocal)->Send(msg);
    }
  }
  void Process(const Message& msg) override {
    if (ts_ && ts_->Handle(msg)) return;
    if (msg.meta.simple_app) { ProcessSimple(msg); return; }
    if (msg.meta.request) return;  // workers do not serve requests
    const int ts = msg.meta.timestamp;
    KVPairs kvs;
    ExtractKVPairs(msg, &kvs);
    if (!msg.meta.push && kvs.keys.size()) {
      std::lock_guard<std::mutex> lk(mu_);
      recv_kvs_[ts].push_back(kvs);
    }
    if (msg.meta.push && kvs.vals.size()) {  // P3 response with parameters
      std::lock_guard<std::mutex> lk(mu_);
      recv_kvs_[ts].push_back(kvs);
    }
    if (obj_->NumResponse(ts) == static_cast<int>(Postoffice::Get()->num_servers()) - 1) Finish(ts, msg.meta.push);
  }
  void Finish(int ts, bool push) {
    std::vector<KVPairs> parts;
    PullTarget tgt{SArray<Key>(), nullptr, nullptr};
    SArray<char> p3;
    {
      std::lock_guard<std::mutex> lk(mu_);
      auto it = recv_kvs_.find(ts);
      if (it != recv_kvs_.end()) { parts = it->second; recv_kvs_.erase(it); }
      if (!push) { auto jt = pull_targets_.find(ts); if (jt != pull_targets_.end()) { tgt = jt->second; pull_targets_.erase(jt); } }
      else { auto jt = p3_targets_.find(ts); if (jt != p3_targets_.end()) { p3 = jt->second; p3_targets_.erase(jt); } }
    }
    if (!parts.empty()) {
      std::sort(parts.begin(), parts.end(), [](const KVPairs& a, const KVPairs& b) { return a.keys.front() < b.keys.front(); });
      size_t total = 0;
      for (auto& s : parts) total += s.vals.size();
      char* dst = nullptr;
      if (!push && tgt.vals) {
        if (tgt.vals->empty()) tgt.vals->resize(total);
        HIPS_CHECK_MSG(tgt.vals->size() == total, "pull size mismatch");
        dst = tgt.vals->data();
      } else if (push && p3.size()) {
        HIPS_CHECK_MSG(p3.size() == total, "P3 response size mismatch");
        dst = p3.data();
      }
      if (dst) {
        std::vector<int> lens_all;
        for (auto& s : parts) { memcpy(dst, s.vals.data(), s.vals.size()); dst += s.vals.size(); for (int l : s.lens) lens_all.push_back(l); }
        if (!push && tgt.lens) { tgt.lens->resize(lens_all.size()); if (!lens_all.empty()) memcpy(tgt.lens->data(), lens_all.data(), lens_all.size() * sizeof(int)); }
      }
    }
    RunCallback(ts);
  }
  std::mutex mu_;
  std::unordered_map<int, std::vector<KVPairs>> recv_kvs_;
  std::unordered_map<int, Callback> callbacks_;
  std::unordered_map<int, PullTarget> pull_targets_;
  std::unordered_map<int, SArray<char>> p3_targets_;
  std::unique_ptr<TSNode> ts_;
};

// ------------------------------------------------------------------------------------------------ KVServer
class KVServer : public SimpleApp {
 public:
  using ReqHandle = std::function<void(const KVMeta&, const KVPairs&, KVServer*)>;
  explicit KVServer(int app_id) : SimpleApp() {
    obj_.reset(new Customer(app_id, app_id, [this](const Message& m) { Process(m); }, true));
    Environment* e = Environment::Get();
    enable_p3 = e->GetInt("ENABLE_P3", 0) != 0;
    enable_inter_ts = e->GetInt("ENABLE_INTER_TS", 0) != 0;
    enable_intra_ts = e->GetInt("ENABLE_INTRA_TS", 0) != 0;
    enable_dgt = e->GetInt("ENABLE_DGT", 0);
    Postoffice* po = Postoffice::Get();
    // TS overlays this server takes part in: as the final receiver / first relay sender of a plane, and (local servers, global plane)
    // as a merging peer.  Compressed and P3 transports bypass the overlay.
    if (enable_intra_ts && !enable_p3 && po->num_workers() > 1 && po->num_servers() == 1) ts_local_.reset(new TSNode(kLocal, app_id, app_id));
    if (enable_inter_ts && !enable_p3 && po->num_global_workers() > 1 && po->num_global_servers() == 1 && po->has_plane(kGlobal))
      ts_global_.reset(new TSNode(kGlobal, app_id, app_id));
  }
  TSNode* ts(Plane p) { return p == kLocal ? ts_local_.get() : ts_global_.get(); }
  ~KVServer() override { obj_.reset(); }
  void set_request_handle(const ReqHandle& h) { request_handle_kv_ = h; }       // requests from workers (local) / local servers (global)
  void set_response_handle(const ReqHandle& h) { response_handle_kv_ = h; }     // responses to OUR global-plane requests

  void Response(const KVMeta& req, const KVPairs& res = KVPairs()) {
    Message msg;
    msg.meta.app_id = obj_->app_id();
    msg.meta.customer_id = req.customer_id;
    msg.meta.request = false;
    msg.meta.push = req.push;
    msg.meta.head = req.cmd;
    msg.meta.timestamp = req.timestamp;
    msg.meta.recver = req.sender;
    msg.meta.key = req.key; msg.meta.version = req.version;
    FillDataMessage(&msg, res);
    Postoffice::Get()->van(static_cast<Plane>(req.plane))->Send(msg);
  }

  // ---- client side on the GLOBAL plane (local server -> global servers) ----------------------------------------------
  int Push(const SArray<Key>& keys, const SArray<char>& vals, const SArray<int>& lens, int cmd, int priority = 0, int int_key = 0,
           bool allow_dgt = false, bool allow_ts = false) {
    const int ts = obj_->NewRequest(kServerGroup, kGlobal);
    if (ts_global_ && allow_ts && keys.size() == 1) {   // inter-party TSEngine: merge with other local servers on the way to the global server
      ts_global_->Offer(int_key, cmd, keys[0], vals.data(), vals.size(),
                        TSOrigin{Postoffice::Get()->van(kGlobal)->my_node().id, ts, obj_->customer_id()});
      return ts;
    }
    KVPairs kvs; kvs.keys = keys; kvs.vals = vals; kvs.lens = lens;
    SendGlobal(ts, true, cmd, kvs, priority, int_key, allow_dgt);
    return ts;
  }
  int Pull(const SArray<Key>& keys, int cmd, int priority = 0, int int_key = 0) {
    const int ts = obj_->NewRequest(kServerGroup, kGlobal);
    KVPairs kvs; kvs.keys = keys;
    SendGlobal(ts, false, cmd, kvs, priority, int_key, false);
    return ts;
  }
  int NumResponse(int ts) { return obj_->NumResponse(ts); }
  void WaitGlobal(int ts) { obj_->WaitRequest(ts); }

  bool enable_p3 = false, enable_inter_ts = false, enable_intra_ts = false;
  int enable_dgt = 0;

 private:
  void SendGlobal(int ts, bool push, int cmd, const KVPairs& kvs, int priority, int int_key, bool allow_dgt) {
    std::vector<std::pair<bool, KVPairs>> sliced;
    DefaultSlicer(kvs, Postoffice::Get()->GetServerKeyRanges(kGlobal), &sliced);
    int skipped = 0;
    for (auto& s : sliced) if (!s.first) ++skipped;
    obj_->AddResponse(ts, skipped);
    Van* van = Postoffice::Get()->van(kGlobal);
    for (size_t i = 0; i < sliced.size(); ++i) {
      if (!sliced[i].first) continue;
      Message msg;
      msg.meta.app_id = obj_->app_id();
      msg.meta.customer_id = obj_->customer_id();
      msg.meta.request = true;
      msg.meta.push = push;
      msg.meta.head = cmd;
      msg.meta.timestamp = ts;
      msg.meta.recver = ServerRankToID(static_cast<int>(i), kGlobal);
      msg.meta.priority = priority;
      msg.meta.key = int_key;
      msg.meta.iters = 1;
      FillDataMessage(&msg, sliced[i].second);
      // DGT: only dense fp32 default pushes local server -> global server are split into ranked blocks (kv_app.h:918-919)
      if (push && allow_dgt && van->dgt_sender() != nullptr && sliced[i].second.keys.size() == 1 &&
          sliced[i].second.vals.size() > static_cast<size_t>(van->dgt_sender()->config().block_bytes)) {
        van->dgt_sender()->SendSplit(msg, int_key);
      } else {
        van->Send(msg);
      }
    }
  }
  void Process(const Message& msg) override {
    TSNode* t = ts(static_cast<Plane>(msg.meta.plane));
    if (t && t->Handle(msg)) return;
    if (msg.meta.simple_app) { ProcessSimple(msg); return; }
    KVMeta meta;
    meta.body = msg.meta.body;
    meta.cmd = msg.meta.head; meta.push = msg.meta.push; meta.sender = msg.meta.sender; meta.timestamp = msg.meta.timestamp;
    meta.customer_id = msg.meta.customer_id; meta.priority = msg.meta.priority; meta.key = msg.meta.key; meta.version = msg.meta.version;
    meta.num_merge = msg.meta.iters > 0 ? msg.meta.iters : 1; meta.plane = msg.meta.plane; meta.app_id = msg.meta.app_id;
    KVPairs data;
    ExtractKVPairs(msg, &data);
    if (msg.meta.request) {
      HIPS_CHECK(request_handle_kv_);
      request_handle_kv_(meta, data, this);
    } else if (response_handle_kv_) {
      response_handle_kv_(meta, data, this);
    }
  }
  ReqHandle request_handle_kv_, response_handle_kv_;
  std::unique_ptr<TSNode> ts_local_, ts_global_;
};

}  // namespace hips
