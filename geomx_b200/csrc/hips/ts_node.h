// TSEngine — node side: peer-merge of pushes and relay-broadcast of fresh parameters, driven by the plane's scheduler (tsengine.h).
//
// Parity (behaviour, not structure): KVWorker::ZPush under ENABLE_INTRA_TS (3rdparty/ps-lite/include/ps/kv_app.h:171-202), TS_Process
// (:1112-1179), AutoPull / AutoPullUpdate (:1040-1076, :1409-1455), KVServer::Process TS routing (:1227-1307) and the merge handlers
// WorkersMerge (src/kvstore/kvstore_dist.h:91-173, kvstore_dist_server.h:228-310).
//
//   push  : Offer(key, bytes) keeps ONE merge slot per key (sum so far, number of merged contributions, and the list of ORIGINS =
//           (node id, request timestamp) of every contribution) and ASKPUSHes the scheduler.  The scheduler pairs askers; the node that
//           is told to send ships its whole slot to the peer (which merges and asks again) or to the server.  The origin list travels
//           with the data so that the server acknowledges every contributor's own request when the round completes — no contributor
//           ever needs to know where its gradient was merged.
//   pull  : the holder of fresh parameters of (key, version) asks the scheduler for a receiver (ASKPULL, reporting the throughput it
//           measured on its previous transfer), sends the bytes, waits for the receiver's AUTOPULLREPLY, and asks again until the
//           scheduler answers -1.  Every receiver starts relaying as soon as it holds the version, so the broadcast fans out as a tree
//           whose shape follows the measured link throughputs.
//
// One TSNode per (application, plane): workers use it on the local plane (intra-party TS), local servers on the global plane (inter-
// party TS); servers of a plane use only the relay half plus AskAsServer().
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "customer.h"
#include "half.h"
#include "postoffice.h"

namespace hips {

constexpr int kTSMergeMsg = 2;   // Meta::msg_type of a peer-to-peer merge payload (1 = DGT block)
constexpr int kTSRelayMsg = 3;   // Meta::msg_type of a relayed parameter payload

struct TSOrigin { int sender = 0, timestamp = 0, customer = 0; };

inline std::string EncodeOrigins(const std::vector<TSOrigin>& v) {
  std::string s;
  for (const auto& o : v) s += std::to_string(o.sender) + "," + std::to_string(o.timestamp) + "," + std::to_string(o.customer) + ";";
  return s;
}
inline std::vector<TSOrigin> DecodeOrigins(const std::string& s) {
  std::vector<TSOrigin> out;
  size_t pos = 0;
  while (pos < s.size()) {
    const size_t end = s.find(';', pos);
    if (end == std::string::npos) break;
    TSOrigin o;
    if (sscanf(s.c_str() + pos, "%d,%d,%d", &o.sender, &o.timestamp, &o.customer) == 3) out.push_back(o);
    pos = end + 1;
  }
  return out;
}

// dtype flag of a Cantor-paired data cmd (kvstore_dist_server.h GetCommandType): y of (x, y)
inline int TSDTypeOfCmd(int cmd) {
  int w = 0;
  while ((w + 1) * (w + 2) / 2 <= cmd) ++w;
  return cmd - w * (w + 1) / 2;
}

// dst += src in the payload dtype (0 fp32, 2 fp16, 12 bf16; anything else is summed as fp32 words)
inline void TSMergeBytes(char* dst, const char* src, size_t nbytes, int dtype) {
  if (dtype == 2 || dtype == 12) {
    uint16_t* d = reinterpret_cast<uint16_t*>(dst);
    const uint16_t* s = reinterpret_cast<const uint16_t*>(src);
    const size_t n = nbytes / 2;
    if (dtype == 2) for (size_t i = 0; i < n; ++i) d[i] = FloatToHalf(HalfToFloat(d[i]) + HalfToFloat(s[i]));
    else for (size_t i = 0; i < n; ++i) d[i] = FloatToBF16(BF16ToFloat(d[i]) + BF16ToFloat(s[i]));
    return;
  }
  float* d = reinterpret_cast<float*>(dst);
  const float* s = reinterpret_cast<const float*>(src);
  const size_t n = nbytes / 4;
  for (size_t i = 0; i < n; ++i) d[i] += s[i];
}

class TSNode {
 public:
  // on_relayed(key, version, cmd, bytes): a fresh value arrived through the relay (local servers feed it to their state machine)
  using RelayedFn = std::function<void(int, int, int, const std::vector<char>&)>;

  TSNode(Plane plane, int app_id, int customer_id) : plane_(plane), app_id_(app_id), customer_id_(customer_id) {
    server_id_ = ServerRankToID(0, plane);
  }
  void set_on_relayed(const RelayedFn& f) { on_relayed_ = f; }

  // ------------------------------------------------------------------------------------------------ push merge
  void Offer(int key, int cmd, Key ps_key, const char* bytes, size_t nbytes, const TSOrigin& origin) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      Slot& s = slots_[key];
      if (s.count == 0) { s.buf.assign(bytes, bytes + nbytes); s.cmd = cmd; s.ps_key = ps_key; s.origins.clear(); }
      else TSMergeBytes(s.buf.data(), bytes, std::min(nbytes, s.buf.size()), TSDTypeOfCmd(s.cmd));
      s.count += 1;
      s.origins.push_back(origin);
    }
    AskPush(key);
  }
  // servers take part in the pairing as the final receiver: they ask once per round and again after every partial delivery
  void AskAsServer(int key) { AskPush(key); }

  // returns true if the message belonged to the TS overlay (the application must not process it any further)
  bool Handle(const Message& msg) {
    const int c = msg.meta.control.cmd;
    if (c == Control::REPLY) {
      if (msg.meta.push) ShipSlot(msg.meta.key, msg.meta.iters);
      else OnPullReply(msg);
      return true;
    }
    if (c == Control::AUTOPULLREPLY) { OnRelayAck(msg); return true; }
    if (!msg.meta.control.empty()) return false;
    if (msg.meta.msg_type == kTSMergeMsg && msg.meta.request && msg.meta.push) { OnMerge(msg); return true; }
    if (msg.meta.msg_type == kTSRelayMsg && msg.meta.request && !msg.meta.push) { OnRelayed(msg); return true; }
    return false;
  }

  // ------------------------------------------------------------------------------------------------ pull relay
  // start (or continue) broadcasting `bytes` as version `version` of `key`
  void Relay(int key, int version, int cmd, Key ps_key, const char* bytes, size_t nbytes) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      RelayBuf& r = relay_[key];
      r.version = version; r.cmd = cmd; r.ps_key = ps_key; r.bytes.assign(bytes, bytes + nbytes);
      r.last_recv = -1; r.last_tput = -1;
    }
    AskPull(key, version, -1, -1);
  }
  // blocks until version >= `version` of `key` has been relayed to this node, then copies it out
  void WaitRelayed(int key, int version, void* out, size_t nbytes) {
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { auto it = received_.find(key); return it != received_.end() && it->second.version >= version; });
    const auto& b = received_[key].bytes;
    memcpy(out, b.data(), std::min(nbytes, b.size()));
  }
  int relayed_version(int key) {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = received_.find(key);
    return it == received_.end() ? -1 : it->second.version;
  }
  // counters for tests / the profiler
  long merges_received() const { return merges_received_; }
  long relays_sent() const { return relays_sent_; }

 private:
  struct Slot { std::vector<char> buf; std::vector<TSOrigin> origins; int count = 0, cmd = 0; Key ps_key = 0; };
  struct RelayBuf {
    std::vector<char> bytes; int version = -1, cmd = 0; Key ps_key = 0;
    int last_recv = -1; long last_tput = -1;
    std::chrono::steady_clock::time_point t0;
  };
  struct Received { std::vector<char> bytes; int version = -1; };

  Van* van() { return Postoffice::Get()->van(plane_); }

  void AskPush(int key) {
    Message m;
    m.meta.recver = kScheduler;
    m.meta.control.cmd = Control::ASKPUSH;
    m.meta.request = true; m.meta.push = true;
    m.meta.key = key;
    m.meta.app_id = app_id_; m.meta.customer_id = customer_id_; m.meta.timestamp = 0;
    van()->Send(m);
  }

  void ShipSlot(int key, int dest) {
    Slot s;
    {
      std::lock_guard<std::mutex> lk(mu_);
      auto it = slots_.find(key);
      if (it == slots_.end() || it->second.count == 0) return;   // nothing to send (duplicate decision)
      s = std::move(it->second);
      slots_.erase(it);
    }
    Message msg;
    msg.meta.app_id = app_id_;
    msg.meta.customer_id = customer_id_;
    msg.meta.request = true; msg.meta.push = true;
    msg.meta.head = s.cmd;
    msg.meta.timestamp = s.origins.empty() ? 0 : s.origins[0].timestamp;
    msg.meta.recver = dest;
    msg.meta.key = key;
    msg.meta.iters = s.count;
    msg.meta.body = EncodeOrigins(s.origins);
    msg.meta.msg_type = dest == server_id_ ? 0 : kTSMergeMsg;
    SArray<Key> keys; keys.push_back(s.ps_key);
    SArray<char> vals; vals.CopyFrom(s.buf.data(), s.buf.size());
    SArray<int> lens; lens.push_back(static_cast<int>(s.buf.size()));
    msg.AddData(keys); msg.AddData(vals); msg.AddData(lens);
    van()->Send(msg);
  }

  void OnMerge(const Message& msg) {
    if (msg.data.size() < 2) return;
    const int key = msg.meta.key;
    {
      std::lock_guard<std::mutex> lk(mu_);
      Slot& s = slots_[key];
      const SArray<char>& v = msg.data[1];
      if (s.count == 0) {
        s.buf.assign(v.data(), v.data() + v.size()); s.cmd = msg.meta.head; s.origins.clear();
        SArray<Key> k(msg.data[0]); s.ps_key = k.size() ? k[0] : 0;
      } else {
        TSMergeBytes(s.buf.data(), v.data(), std::min(v.size(), s.buf.size()), TSDTypeOfCmd(s.cmd));
      }
      s.count += std::max(1, msg.meta.iters);
      for (const auto& o : DecodeOrigins(msg.meta.body)) s.origins.push_back(o);
      ++merges_received_;
    }
    AskPush(key);   // the merged slot needs a new destination
  }

  void AskPull(int key, int version, int last_recv, long last_tput) {
    Message m;
    m.meta.recver = kScheduler;
    m.meta.control.cmd = Control::ASKPULL;
    m.meta.request = true; m.meta.push = false;
    m.meta.key = key; m.meta.version = version;
    m.meta.app_id = static_cast<int>(last_tput);     // report of the previous transfer (-1: none) ...
    m.meta.customer_id = last_recv;                  // ... and who received it
    m.meta.head = app_id_;                           // echoed back so that the reply finds this application
    m.meta.body = std::to_string(customer_id_);
    m.meta.timestamp = 0;
    van()->Send(m);
  }

  void OnPullReply(const Message& msg) {
    const int key = msg.meta.key, recv = msg.meta.iters;
    Message out;
    {
      std::lock_guard<std::mutex> lk(mu_);
      auto it = relay_.find(key);
      if (it == relay_.end() || it->second.version != msg.meta.version) return;
      RelayBuf& r = it->second;
      if (recv < 0) { relay_.erase(it); return; }    // everybody holds this version
      out.meta.app_id = app_id_;
      out.meta.customer_id = customer_id_;
      out.meta.request = true; out.meta.push = false;
      out.meta.head = r.cmd;
      out.meta.timestamp = 0;
      out.meta.recver = recv;
      out.meta.key = key; out.meta.version = r.version;
      out.meta.msg_type = kTSRelayMsg;
      SArray<Key> keys; keys.push_back(r.ps_key);
      SArray<char> vals; vals.CopyFrom(r.bytes.data(), r.bytes.size());
      SArray<int> lens; lens.push_back(static_cast<int>(r.bytes.size()));
      out.AddData(keys); out.AddData(vals); out.AddData(lens);
      r.last_recv = recv;
      r.t0 = std::chrono::steady_clock::now();
      ++relays_sent_;
    }
    van()->Send(out);
  }

  void OnRelayAck(const Message& msg) {
    const int key = msg.meta.key;
    int version, last_recv; long tput;
    {
      std::lock_guard<std::mutex> lk(mu_);
      auto it = relay_.find(key);
      if (it == relay_.end() || it->second.version != msg.meta.version) return;
      RelayBuf& r = it->second;
      const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - r.t0).count();
      r.last_tput = static_cast<long>(r.bytes.size() / std::max(ms, 1e-3)) + 1;   // bytes per ms
      version = r.version; last_recv = r.last_recv; tput = r.last_tput;
    }
    AskPull(key, version, last_recv, tput);
  }

  void OnRelayed(const Message& msg) {
    if (msg.data.size() < 2) return;
    const int key = msg.meta.key, version = msg.meta.version;
    const SArray<char>& v = msg.data[1];
    std::vector<char> bytes(v.data(), v.data() + v.size());
    SArray<Key> k(msg.data[0]);
    const Key ps_key = k.size() ? k[0] : 0;
    bool fresh = false;
    {
      std::lock_guard<std::mutex> lk(mu_);
      Received& r = received_[key];
      if (version > r.version) { r.version = version; r.bytes = bytes; fresh = true; }
    }
    cv_.notify_all();
    // acknowledge to the sender (it measures the link and asks for its next receiver) ...
    Message ack;
    ack.meta.recver = msg.meta.sender;
    ack.meta.control.cmd = Control::AUTOPULLREPLY;
    ack.meta.request = true;
    ack.meta.key = key; ack.meta.version = version;
    ack.meta.app_id = msg.meta.app_id; ack.meta.customer_id = msg.meta.customer_id; ack.meta.timestamp = 0;
    van()->Send(ack);
    if (!fresh) return;
    if (on_relayed_) on_relayed_(key, version, msg.meta.head, bytes);
    // ... and relay onward ourselves
    Relay(key, version, msg.meta.head, ps_key, bytes.data(), bytes.size());
  }

  Plane plane_;
  int app_id_, customer_id_, server_id_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::map<int, Slot> slots_;
  std::map<int, RelayBuf> relay_;
  std::map<int, Received> received_;
  RelayedFn on_relayed_;
  std::atomic<long> merges_received_{0}, relays_sent_{0};   // read by the test / profiler thread while the TS threads count
};

}  // namespace hips
