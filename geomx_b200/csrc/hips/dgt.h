// DGT — Differential Gradient Transmission (block contribution ranking + prioritised / lossy-tolerant channels).
//
// Parity: 3rdparty/ps-lite/include/ps/kv_app.h:842-1022 (InitDGT, EvalMsgContribution: EMA(alpha) of mean|g| per (key, seq) block;
// GetChannel; block split + sort + channel assignment in KVServer::Send) and src/van.cc:290-370,707-824 (Important_scheduler /
// Unimportant_scheduler — the unimportant queue is drained only while the important queue is empty; 4-bit encode/decode with a
// min/max codebook; receiver-side reassembly into a zero-filled buffer, delivery when the block with seq == seq_end arrives).
// Modes (ENABLE_DGT): 1 = unimportant blocks travel as UDP datagrams with per-message IP_TOS (Van::SendUDP) and may be lost — the receiver
// zero-fills; DGT_UDP_LOSS % additionally drops received unimportant blocks to exercise that tolerance on a loss-free loopback, 2 = reliable but
// priority-ordered, 3 = mode 2 + 4-bit quantised unimportant blocks.  Only dense default pushes from a local server to the global
// servers are DGT-split, as in the reference (kv_app.h:918-919).
#pragma once
#include <algorithm>
#include <atomic>
#include <cmath>
#include <map>
#include <mutex>
#include <random>
#include <thread>
#include <unordered_map>
#include <vector>

#include "env.h"
#include "message.h"
#include "threadsafe_queue.h"
#include "van.h"

namespace hips {

struct DGTConfig {
  int mode = 0;          // ENABLE_DGT
  int block_bytes = 4096;
  int channels = 3;      // DMLC_UDP_CHANNEL_NUM
  float k = 0.5f;        // DMLC_K: fraction of blocks on the reliable channel
  float alpha = 0.3f;    // DGT_CONTRIBUTION_ALPHA
  int loss_pct = 0;      // DGT_UDP_LOSS (emulated datagram loss, mode 1)
  bool adaptive_k = false;  // ADAPTIVE_K_FLAG: k is a share of the contribution mass (see DGTEffectiveK)
  float k_min = 0.2f;       // DMLC_K_MIN: lower bound of the important fraction in adaptive mode
  int info = 0;             // DGT_INFO: log one line per split push (key, blocks, effective k, blocks sent)
  static DGTConfig FromEnv() {
    Environment* e = Environment::Get();
    DGTConfig c;
    c.mode = e->GetInt("ENABLE_DGT", 0);
    c.block_bytes = e->GetInt("DGT_BLOCK_SIZE", 4096);
    c.channels = std::max(1, e->GetInt("DMLC_UDP_CHANNEL_NUM", 3));
    c.k = static_cast<float>(e->GetFloat("DMLC_K", 0.5));
    c.alpha = static_cast<float>(e->GetFloat("DGT_CONTRIBUTION_ALPHA", 0.3));
    c.loss_pct = e->GetInt("DGT_UDP_LOSS", 0);
    c.adaptive_k = e->GetInt("ADAPTIVE_K_FLAG", 0) != 0;
    c.k_min = static_cast<float>(e->GetFloat("DMLC_K_MIN", 0.2));
    c.info = e->GetInt("DGT_INFO", 0);
    return c;
  }
};

// channel of the block ranked `rank` (0 = most important) out of `total`: top k fraction -> 0, the rest evenly over 1..C
inline int DGTGetChannel(int rank, int total, float k, int channels) {
  const int important = static_cast<int>(std::ceil(total * k));
  if (rank < important || channels <= 0) return 0;
  const int rest = total - important;
  const int per = std::max(1, (rest + channels - 1) / channels);
  return std::min(channels, 1 + (rank - important) / per);
}

// ADAPTIVE_K_FLAG / DMLC_K_MIN are parsed but never used by the reference (kv_app.h:844-848).  Here: with the flag set, the important set of
// one push is the shortest prefix of the contribution ranking that carries `k` of the total contribution, at least `k_min` of the blocks —
// returned as the equivalent block fraction so DGTGetChannel stays the single place that maps rank -> channel.
inline float DGTEffectiveK(const std::vector<float>& sorted_desc, float k, bool adaptive, float k_min) {
  const int n = static_cast<int>(sorted_desc.size());
  if (!adaptive || n == 0) return k;
  double total = 0;
  for (float c : sorted_desc) total += c;
  int need = static_cast<int>(std::lround(k_min * n));
  if (total > 0) {
    double cum = 0; int i = 0;
    for (; i < n; ++i) { cum += sorted_desc[i]; if (cum / total >= k - 1e-12) break; }
    need = std::max(need, std::min(n, i + 1));
  }
  need = std::max(1, std::min(n, need));
  return static_cast<float>(need) / static_cast<float>(n);
}

// 4-bit codec: codebook of 16 uniformly spaced centroids in [min, max]
inline void DGTEncode4(const float* src, size_t n, std::vector<char>* out, float* mn, float* mx) {
  float lo = src[0], hi = src[0];
  for (size_t i = 1; i < n; ++i) { lo = std::min(lo, src[i]); hi = std::max(hi, src[i]); }
  *mn = lo; *mx = hi;
  const float step = (hi - lo) / 15.f;
  out->assign((n + 1) / 2, 0);
  for (size_t i = 0; i < n; ++i) {
    int code = step > 0 ? static_cast<int>(std::lround((src[i] - lo) / step)) : 0;
    code = std::max(0, std::min(15, code));
    (*out)[i / 2] |= static_cast<char>((i & 1) ? (code << 4) : code);
  }
}
inline void DGTDecode4(const char* src, size_t n, float mn, float mx, float* dst) {
  const float step = (mx - mn) / 15.f;
  for (size_t i = 0; i < n; ++i) {
    const int b = static_cast<unsigned char>(src[i / 2]);
    const int code = (i & 1) ? (b >> 4) : (b & 15);
    dst[i] = mn + step * code;
  }
}

class DGTSender {
 public:
  explicit DGTSender(Van* van) : van_(van), cfg_(DGTConfig::FromEnv()) {
    important_ = std::thread(&DGTSender::ImportantLoop, this);
    unimportant_ = std::thread(&DGTSender::UnimportantLoop, this);
  }
  void Stop() {
    if (stop_.exchange(true)) return;
    Message t; t.meta.control.cmd = Control::TERMINATE;
    iq_.Push(t); uq_.Push(t);
    important_.join(); unimportant_.join();
  }
  const DGTConfig& config() const { return cfg_; }

  // contribution EMA of one block (kv_app.h:853-876)
  float Contribution(int key, int seq, const float* v, size_t n) {
    double s = 0;
    for (size_t i = 0; i < n; ++i) s += std::fabs(v[i]);
    const float mean = n ? static_cast<float>(s / n) : 0.f;
    std::lock_guard<std::mutex> lk(mu_);
    auto it = contri_.find({key, seq});
    float c = it == contri_.end() ? mean : cfg_.alpha * it->second + (1.f - cfg_.alpha) * mean;
    contri_[{key, seq}] = c;
    return c;
  }

  // split one dense fp32 push `base` (data = keys|vals|lens) into ranked blocks and enqueue them.  Returns #blocks sent.
  int SendSplit(const Message& base, int key) {
    const SArray<char>& vals = base.data[1];
    const int total = static_cast<int>(vals.size());
    const int bb = cfg_.block_bytes;
    const int nblk = (total + bb - 1) / bb;
    struct Blk { int seq; float c; };
    std::vector<Blk> blks(nblk);
    for (int s = 0; s < nblk; ++s) {
      const int off = s * bb, len = std::min(bb, total - off);
      blks[s] = {s, Contribution(key, s, reinterpret_cast<const float*>(vals.data() + off), len / sizeof(float))};
    }
    // sort by contribution (desc); the LAST block stays last: its arrival triggers delivery at the receiver
    std::sort(blks.begin(), blks.end() - 1, [](const Blk& a, const Blk& b) { return a.c > b.c; });
    float k_eff = cfg_.k;
    if (cfg_.adaptive_k) {
      std::vector<float> cs(nblk);
      for (int r = 0; r < nblk; ++r) cs[r] = blks[r].c;
      k_eff = DGTEffectiveK(cs, cfg_.k, true, cfg_.k_min);
    }
    last_k_eff_ = k_eff;
    int sent = 0;
    for (int r = 0; r < nblk; ++r) {
      const Blk& b = blks[r];
      const bool last = b.seq == nblk - 1;
      if (b.c == 0.f && !last) continue;  // zero-contribution blocks are dropped (kv_app.h:973-975)
      Message m;
      m.meta = base.meta;
      m.meta.msg_type = 1;
      m.meta.first_key = key;
      m.meta.seq = b.seq; m.meta.seq_begin = 0; m.meta.seq_end = nblk - 1;
      m.meta.total_bytes = total;
      const int off = b.seq * bb, len = std::min(bb, total - off);
      m.meta.val_bytes = len;
      const int ch = last ? 0 : DGTGetChannel(r, nblk, k_eff, cfg_.channels);
      m.meta.channel = ch;
      m.meta.tos = (cfg_.channels - ch) * 32;
      m.meta.priority = -ch;
      m.contribution = b.c;
      m.data.clear();
      m.data.push_back(base.data[0]);
      if (cfg_.mode == 3 && ch > 0) {
        std::vector<char> enc; float mn, mx;
        DGTEncode4(reinterpret_cast<const float*>(vals.data() + off), len / sizeof(float), &enc, &mn, &mx);
        SArray<char> e; e.CopyFrom(enc.data(), enc.size());
        m.data.push_back(e);
        m.meta.bits_num = 4; m.meta.compr = {mn, mx};
      } else {
        m.data.push_back(vals.segment(off, off + len));
        m.meta.bits_num = 32;
      }
      m.data.push_back(base.data.size() > 2 ? base.data[2] : SArray<char>());
      (ch == 0 ? iq_ : uq_).Push(m);
      ++sent;
    }
    if (cfg_.info) fprintf(stderr, "[DGT] key=%d blocks=%d k_eff=%.3f sent=%d mode=%d\n", key, nblk, k_eff, sent, cfg_.mode);
    return sent;
  }

  float last_effective_k() const { return last_k_eff_; }

 private:
  float last_k_eff_ = 0.f;
  void ImportantLoop() {
    while (true) {
      Message m; iq_.WaitAndPop(&m);
      if (m.meta.control.cmd == Control::TERMINATE) break;
      van_->SendNow(m);
    }
  }
  void UnimportantLoop() {  // van.cc:720-728: only send while the important queue is empty
    while (true) {
      Message m; uq_.WaitAndPop(&m);
      if (m.meta.control.cmd == Control::TERMINATE) break;
      while (iq_.Size() > 0 && !stop_) std::this_thread::yield();
      if (cfg_.mode == 1 && m.meta.channel > 0) van_->SendUDP(m);   // lossy datagram channel with the block's TOS
      else van_->SendNow(m);
    }
  }
  Van* van_;
  DGTConfig cfg_;
  ThreadsafeQueue<Message, MessagePriority> iq_, uq_;
  std::thread important_, unimportant_;
  std::atomic<bool> stop_{false};
  std::mutex mu_;
  std::map<std::pair<int, int>, float> contri_;
};

class DGTReceiver {
 public:
  DGTReceiver() : cfg_(DGTConfig::FromEnv()), rng_(4242) {}
  // returns true and fills `whole` when the final block (seq == seq_end) of a tensor arrived
  bool Add(const Message& blk, Message* whole) {
    const Meta& m = blk.meta;
    if (cfg_.mode == 1 && m.channel > 0 && cfg_.loss_pct > 0 && static_cast<int>(rng_() % 100) < cfg_.loss_pct) return false;  // lost datagram
    // A block comes off the wire (over UDP from anybody who can reach the port): every field is validated before it sizes or indexes
    // memory; a block that does not fit the tensor it claims to belong to is dropped like a lost one.
    if (blk.data.size() < 2 || m.total_bytes <= 0 || static_cast<size_t>(m.total_bytes) > kMaxTensorBytes) return false;
    if (m.seq < 0 || m.seq_end < m.seq || m.val_bytes < 0 || m.val_bytes % static_cast<int>(sizeof(float)) != 0) return false;
    const size_t off = static_cast<size_t>(m.seq) * static_cast<size_t>(cfg_.block_bytes);
    const size_t total = static_cast<size_t>(m.total_bytes), want = static_cast<size_t>(m.val_bytes);
    if (off > total || want > total - off) return false;
    if (m.bits_num == 4) {
      if (m.compr.size() < 2 || blk.data[1].size() < (want / sizeof(float) + 1) / 2) return false;
    } else if (blk.data[1].size() < want) {
      return false;
    }
    const auto id = std::make_tuple(m.sender, m.first_key, m.timestamp);
    if (pending_.size() >= kMaxPending && pending_.find(id) == pending_.end()) pending_.clear();   // partial tensors of senders that went away
    auto& st = pending_[id];
    if (st.buf.size() == 0) { st.buf.resize(total, 0); st.keys = blk.data[0]; st.lens = blk.data.size() > 2 ? blk.data[2] : SArray<char>(); }
    else if (st.buf.size() != total) return false;            // a block that disagrees with the first one about the tensor's size
    if (m.bits_num == 4) DGTDecode4(blk.data[1].data(), want / sizeof(float), m.compr[0], m.compr[1], reinterpret_cast<float*>(st.buf.data() + off));
    else memcpy(st.buf.data() + off, blk.data[1].data(), want);
    if (m.seq != m.seq_end) return false;
    whole->meta = m;
    whole->meta.msg_type = 0; whole->meta.channel = 0;
    whole->data.clear();
    whole->data.push_back(st.keys);
    whole->data.push_back(st.buf);
    SArray<char> lens; lens.resize(sizeof(int));
    int total_i = m.total_bytes; memcpy(lens.data(), &total_i, sizeof(int));
    whole->data.push_back(lens);
    pending_.erase(id);
    return true;
  }

 private:
  static constexpr size_t kMaxTensorBytes = size_t(1) << 31;   // one key's tensor (2 GiB): larger claims are not honoured
  static constexpr size_t kMaxPending = 4096;
  struct State { SArray<char> buf, keys, lens; };
  DGTConfig cfg_;
  std::mt19937 rng_;
  std::map<std::tuple<int, int, int>, State> pending_;
};

}  // namespace hips
