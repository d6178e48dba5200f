// Node / Control / Meta / Message and their binary wire codec.
// Parity: ps-lite include/ps/internal/message.h:68-308 + src/meta.proto:8-79 including the GeoMX additions (global roles, controls
// ADD_GLOBAL_NODE / BARRIER_GLOBAL / AUTOPULLREPLY / ASKPULL / ASKPUSH / REPLY, meta fields first_key, seq, seq_begin, seq_end,
// channel, msg_type, push_op, val_bytes, total_bytes, compr, bits_num, priority, key, version, iters(num_merge), tos).
// The plane (local/global) is a property of the Van a message travels on, so the *_GLOBAL control duplicates collapse.
#pragma once
#include <string>
#include <vector>

#include "base.h"
#include "sarray.h"

namespace hips {

struct Node {
  enum Role { SERVER = 0, WORKER = 1, SCHEDULER = 2 };
  static const int kEmpty = -1;
  int role = WORKER;
  int id = kEmpty;
  std::string hostname;
  int port = kEmpty;
  bool is_recovery = false;
  int customer_id = 0;
  int udp_port = -1;    // DGT mode 1: datagram endpoint for unimportant blocks (reference: udp_port[] per channel, zmq_van.h:95-305)
  int rank_hint = -1;   // a server that already owns a rank in the other plane asks for the same rank (MultiGPS key sharding consistency)
  std::string DebugString() const {
    std::stringstream ss;
    ss << "role=" << (role == SERVER ? "server" : role == WORKER ? "worker" : "scheduler") << (id != kEmpty ? ", id=" + std::to_string(id) : "")
       << ", ip=" << hostname << ", port=" << port << ", is_recovery=" << is_recovery;
    return ss.str();
  }
};

struct Control {
  enum Command { EMPTY = 0, TERMINATE, ADD_NODE, BARRIER, ACK, HEARTBEAT, AUTOPULLREPLY, ASKPULL, ASKPUSH, REPLY };
  int cmd = EMPTY;
  std::vector<Node> node;
  int barrier_group = 0;
  uint64_t msg_sig = 0;
  bool empty() const { return cmd == EMPTY; }
};

struct Meta {
  static const int kEmpty = -1;
  int head = kEmpty;
  int app_id = kEmpty;
  int customer_id = kEmpty;
  int timestamp = kEmpty;
  int sender = kEmpty;
  int recver = kEmpty;
  bool request = false;
  bool push = false;
  bool simple_app = false;
  std::string body;
  Control control;
  // GeoMX additions
  int priority = 0;       // P3
  int key = 0;            // TSEngine / P3: original integer key
  int version = 0;        // TSEngine
  int iters = 0;          // TSEngine: num_merge (or destination id in scheduler replies)
  int first_key = 0;      // DGT
  int seq = 0, seq_begin = 0, seq_end = 0, channel = 0, msg_type = 0, push_op = 0;
  int val_bytes = 0, total_bytes = 0, bits_num = 0, tos = 0;
  int keys_len = 0, vals_len = 0, lens_len = 0;
  std::vector<float> compr;  // DGT 4-bit codebook (min, max)
  int data_size = 0;
  int plane = 0;             // set by the receiving Van (not serialised): which plane the message arrived on
};

struct Message {
  Meta meta;
  std::vector<SArray<char>> data;
  float contribution = 0.f;  // DGT (not serialised)
  template <typename V>
  void AddData(const SArray<V>& val) { data.push_back(SArray<char>(val)); meta.data_size += static_cast<int>(val.size() * sizeof(V)); }
};

// ---- codec: [meta bytes] ; framing (magic, lengths) is done by the Van ------------------------------------------------
class Writer {
 public:
  template <typename T> void Put(const T& v) { const char* p = reinterpret_cast<const char*>(&v); buf_.insert(buf_.end(), p, p + sizeof(T)); }
  void PutStr(const std::string& s) { Put<uint32_t>(static_cast<uint32_t>(s.size())); buf_.insert(buf_.end(), s.begin(), s.end()); }
  std::vector<char>& buf() { return buf_; }
 private:
  std::vector<char> buf_;
};
class Reader {
 public:
  Reader(const char* p, size_t n) : p_(p), n_(n) {}
  template <typename T> T Get() { HIPS_CHECK(o_ + sizeof(T) <= n_); T v; memcpy(&v, p_ + o_, sizeof(T)); o_ += sizeof(T); return v; }
  std::string GetStr() { uint32_t n = Get<uint32_t>(); HIPS_CHECK(o_ + n <= n_); std::string s(p_ + o_, n); o_ += n; return s; }
  size_t remaining() const { return n_ - o_; }
 private:
  const char* p_; size_t n_, o_ = 0;
};

inline void PackMeta(const Meta& m, std::vector<char>* out) {
  Writer w;
  w.Put<int32_t>(m.head); w.Put<int32_t>(m.app_id); w.Put<int32_t>(m.customer_id); w.Put<int32_t>(m.timestamp);
  w.Put<int32_t>(m.sender); w.Put<int32_t>(m.recver);
  w.Put<uint8_t>(m.request); w.Put<uint8_t>(m.push); w.Put<uint8_t>(m.simple_app);
  w.PutStr(m.body);
  w.Put<int32_t>(m.priority); w.Put<int32_t>(m.key); w.Put<int32_t>(m.version); w.Put<int32_t>(m.iters);
  w.Put<int32_t>(m.first_key); w.Put<int32_t>(m.seq); w.Put<int32_t>(m.seq_begin); w.Put<int32_t>(m.seq_end);
  w.Put<int32_t>(m.channel); w.Put<int32_t>(m.msg_type); w.Put<int32_t>(m.push_op); w.Put<int32_t>(m.val_bytes);
  w.Put<int32_t>(m.total_bytes); w.Put<int32_t>(m.bits_num); w.Put<int32_t>(m.tos);
  w.Put<int32_t>(m.keys_len); w.Put<int32_t>(m.vals_len); w.Put<int32_t>(m.lens_len);
  w.Put<uint32_t>(static_cast<uint32_t>(m.compr.size()));
  for (float c : m.compr) w.Put<float>(c);
  w.Put<int32_t>(m.control.cmd);
  if (m.control.cmd != Control::EMPTY) {
    w.Put<int32_t>(m.control.barrier_group); w.Put<uint64_t>(m.control.msg_sig);
    w.Put<uint32_t>(static_cast<uint32_t>(m.control.node.size()));
    for (const auto& n : m.control.node) {
      w.Put<int32_t>(n.role); w.Put<int32_t>(n.id); w.PutStr(n.hostname); w.Put<int32_t>(n.port);
      w.Put<uint8_t>(n.is_recovery); w.Put<int32_t>(n.customer_id); w.Put<int32_t>(n.rank_hint); w.Put<int32_t>(n.udp_port);
    }
  }
  out->swap(w.buf());
}

inline void UnpackMeta(const char* buf, size_t n, Meta* m) {
  Reader r(buf, n);
  m->head = r.Get<int32_t>(); m->app_id = r.Get<int32_t>(); m->customer_id = r.Get<int32_t>(); m->timestamp = r.Get<int32_t>();
  m->sender = r.Get<int32_t>(); m->recver = r.Get<int32_t>();
  m->request = r.Get<uint8_t>(); m->push = r.Get<uint8_t>(); m->simple_app = r.Get<uint8_t>();
  m->body = r.GetStr();
  m->priority = r.Get<int32_t>(); m->key = r.Get<int32_t>(); m->version = r.Get<int32_t>(); m->iters = r.Get<int32_t>();
  m->first_key = r.Get<int32_t>(); m->seq = r.Get<int32_t>(); m->seq_begin = r.Get<int32_t>(); m->seq_end = r.Get<int32_t>();
  m->channel = r.Get<int32_t>(); m->msg_type = r.Get<int32_t>(); m->push_op = r.Get<int32_t>(); m->val_bytes = r.Get<int32_t>();
  m->total_bytes = r.Get<int32_t>(); m->bits_num = r.Get<int32_t>(); m->tos = r.Get<int32_t>();
  m->keys_len = r.Get<int32_t>(); m->vals_len = r.Get<int32_t>(); m->lens_len = r.Get<int32_t>();
  uint32_t nc = r.Get<uint32_t>();
  HIPS_CHECK_MSG(static_cast<size_t>(nc) * sizeof(float) <= r.remaining(), "meta: codebook length exceeds the frame");   // (never allocate from an unchecked count)
  m->compr.resize(nc);
  for (uint32_t i = 0; i < nc; ++i) m->compr[i] = r.Get<float>();
  m->control.cmd = r.Get<int32_t>();
  m->control.node.clear();
  if (m->control.cmd != Control::EMPTY) {
    m->control.barrier_group = r.Get<int32_t>(); m->control.msg_sig = r.Get<uint64_t>();
    uint32_t nn = r.Get<uint32_t>();
    for (uint32_t i = 0; i < nn; ++i) {
      Node nd;
      nd.role = r.Get<int32_t>(); nd.id = r.Get<int32_t>(); nd.hostname = r.GetStr(); nd.port = r.Get<int32_t>();
      nd.is_recovery = r.Get<uint8_t>(); nd.customer_id = r.Get<int32_t>(); nd.rank_hint = r.Get<int32_t>(); nd.udp_port = r.Get<int32_t>();
      m->control.node.push_back(nd);
    }
  }
}

struct MessagePriority {
  int operator()(const Message& m) const { return m.meta.priority; }
};

}  // namespace hips
