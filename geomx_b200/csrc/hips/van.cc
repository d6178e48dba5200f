// Van implementation: sockets, framing, registration, barriers, heartbeat, P3 priority sender, fault injection.
#include "van.h"

#include <arpa/inet.h>
#include <fcntl.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <sys/socket.h>
#include <sys/uio.h>
#include <sys/un.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <random>

#include "customer.h"
#include "dgt.h"
#include "network_utils.h"
#include "postoffice.h"
#include "resender.h"
#include "tsengine.h"

namespace hips {

int Verbose() {
  static int v = Environment::Get()->GetInt("PS_VERBOSE", 0);
  return v;
}

static const uint32_t kMagic = 0x48695053;  // "HiPS"
static const uint32_t kMaxMetaBytes = 16u << 20;   // node tables of a few thousand nodes fit easily; anything larger is not a HiPS frame
static const uint32_t kMaxFrameParts = 64;         // keys | vals | lens (+ a few auxiliary parts)

Van::Van(Postoffice* po, Plane plane) : po_(po), plane_(plane) {}
Van::~Van() {}

// ------------------------------------------------------------------------------------------------ sockets
static bool WriteAll(int fd, struct iovec* iov, int cnt) {
  while (cnt > 0) {
    ssize_t n = ::writev(fd, iov, cnt > 64 ? 64 : cnt);
    if (n < 0) {
      if (errno == EINTR) continue;
      return false;
    }
    while (n > 0 && cnt > 0) {
      if (static_cast<size_t>(n) >= iov->iov_len) { n -= iov->iov_len; ++iov; --cnt; }
      else { iov->iov_base = static_cast<char*>(iov->iov_base) + n; iov->iov_len -= n; n = 0; }
    }
    while (cnt > 0 && iov->iov_len == 0) { ++iov; --cnt; }
  }
  return true;
}
static bool ReadAll(int fd, void* buf, size_t n) {
  char* p = static_cast<char*>(buf);
  while (n > 0) {
    ssize_t r = ::read(fd, p, n);
    if (r == 0) return false;
    if (r < 0) {
      if (errno == EINTR) continue;
      return false;
    }
    p += r; n -= r;
  }
  return true;
}

int Van::Bind(Node* node, int max_retry) {
  Environment* env = Environment::Get();
  use_unix_ = env->GetInt("DMLC_LOCAL", 0) != 0;
  int port = node->port;
  std::mt19937 rng(static_cast<unsigned>(std::chrono::steady_clock::now().time_since_epoch().count()) + getpid());
  for (int i = 0; i <= max_retry; ++i) {
    int fd;
    if (use_unix_) {
      fd = ::socket(AF_UNIX, SOCK_STREAM, 0);
      struct sockaddr_un a; memset(&a, 0, sizeof(a)); a.sun_family = AF_UNIX;
      unix_path_ = "/tmp/hips_" + std::to_string(port);
      strncpy(a.sun_path, unix_path_.c_str(), sizeof(a.sun_path) - 1);
      ::unlink(unix_path_.c_str());
      if (::bind(fd, reinterpret_cast<struct sockaddr*>(&a), sizeof(a)) == 0 && ::listen(fd, 256) == 0) { listen_fd_ = fd; return port; }
    } else {
      fd = ::socket(AF_INET, SOCK_STREAM, 0);
      int one = 1;
      setsockopt(fd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
      struct sockaddr_in a; memset(&a, 0, sizeof(a)); a.sin_family = AF_INET; a.sin_addr.s_addr = htonl(INADDR_ANY); a.sin_port = htons(port);
      if (::bind(fd, reinterpret_cast<struct sockaddr*>(&a), sizeof(a)) == 0 && ::listen(fd, 256) == 0) { listen_fd_ = fd; return port; }
    }
    ::close(fd);
    if (i == max_retry) break;
    port = 10000 + rng() % 40000;
  }
  return -1;
}

int Van::ConnectFd(const Node& node) {
  for (int attempt = 0; attempt < 600; ++attempt) {
    int fd;
    int rc;
    if (use_unix_) {
      fd = ::socket(AF_UNIX, SOCK_STREAM, 0);
      struct sockaddr_un a; memset(&a, 0, sizeof(a)); a.sun_family = AF_UNIX;
      std::string path = "/tmp/hips_" + std::to_string(node.port);
      strncpy(a.sun_path, path.c_str(), sizeof(a.sun_path) - 1);
      rc = ::connect(fd, reinterpret_cast<struct sockaddr*>(&a), sizeof(a));
    } else {
      fd = ::socket(AF_INET, SOCK_STREAM, 0);
      int one = 1;
      setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
      struct sockaddr_in a; memset(&a, 0, sizeof(a)); a.sin_family = AF_INET; a.sin_port = htons(node.port);
      if (inet_pton(AF_INET, node.hostname.c_str(), &a.sin_addr) != 1) {
        struct hostent* he = gethostbyname(node.hostname.c_str());
        HIPS_CHECK_MSG(he != nullptr, "cannot resolve " + node.hostname);
        memcpy(&a.sin_addr, he->h_addr_list[0], he->h_length);
      }
      rc = ::connect(fd, reinterpret_cast<struct sockaddr*>(&a), sizeof(a));
    }
    if (rc == 0) return fd;
    ::close(fd);
    if (stop_) return -1;
    std::this_thread::sleep_for(std::chrono::milliseconds(100));  // peer may not be listening yet (scheduler started later)
  }
  return -1;
}

int Van::SendFrame(int fd, const Message& msg) {
  std::vector<char> meta;
  PackMeta(msg.meta, &meta);
  const uint32_t nd = static_cast<uint32_t>(msg.data.size());
  std::vector<char> hdr(12 + 8 * nd);
  uint32_t m = kMagic, ml = static_cast<uint32_t>(meta.size());
  memcpy(&hdr[0], &m, 4); memcpy(&hdr[4], &ml, 4); memcpy(&hdr[8], &nd, 4);
  size_t bytes = 0;
  for (uint32_t i = 0; i < nd; ++i) { uint64_t l = msg.data[i].size(); memcpy(&hdr[12 + 8 * i], &l, 8); bytes += l; }
  std::vector<struct iovec> iov(2 + nd);
  iov[0].iov_base = hdr.data(); iov[0].iov_len = hdr.size();
  iov[1].iov_base = meta.data(); iov[1].iov_len = meta.size();
  int cnt = 2;
  for (uint32_t i = 0; i < nd; ++i) {
    if (msg.data[i].size() == 0) continue;
    iov[cnt].iov_base = msg.data[i].data(); iov[cnt].iov_len = msg.data[i].size(); ++cnt;
  }
  if (!WriteAll(fd, iov.data(), cnt)) return -1;
  return static_cast<int>(bytes + meta.size());
}

// one self-contained datagram: same layout as a TCP frame (magic | meta len | #data | data lens | meta | data...)
static void PackDatagram(const Message& msg, std::vector<char>* out) {
  std::vector<char> meta;
  PackMeta(msg.meta, &meta);
  const uint32_t nd = static_cast<uint32_t>(msg.data.size());
  size_t total = 12 + 8 * nd + meta.size();
  for (uint32_t i = 0; i < nd; ++i) total += msg.data[i].size();
  out->resize(total);
  char* p = out->data();
  uint32_t m = kMagic, ml = static_cast<uint32_t>(meta.size());
  memcpy(p, &m, 4); memcpy(p + 4, &ml, 4); memcpy(p + 8, &nd, 4); p += 12;
  for (uint32_t i = 0; i < nd; ++i) { uint64_t l = msg.data[i].size(); memcpy(p, &l, 8); p += 8; }
  memcpy(p, meta.data(), meta.size()); p += meta.size();
  for (uint32_t i = 0; i < nd; ++i) if (msg.data[i].size()) { memcpy(p, msg.data[i].data(), msg.data[i].size()); p += msg.data[i].size(); }
}
static bool UnpackDatagram(const char* buf, size_t n, Message* msg) {
  if (n < 12) return false;
  uint32_t h[3];
  memcpy(h, buf, 12);
  if (h[0] != kMagic) return false;
  const uint32_t ml = h[1], nd = h[2];
  size_t pos = 12;
  if (nd > kMaxFrameParts || n < pos + 8ull * nd + ml) return false;
  std::vector<uint64_t> lens(nd);
  if (nd) memcpy(lens.data(), buf + pos, 8 * nd);
  pos += 8ull * nd;
  try { UnpackMeta(buf + pos, ml, &msg->meta); } catch (const std::exception&) { return false; }   // foreign / corrupt datagram
  pos += ml;
  msg->data.clear();
  for (uint32_t i = 0; i < nd; ++i) {
    if (n < pos + lens[i]) return false;
    SArray<char> d;
    if (lens[i]) d.CopyFrom(buf + pos, lens[i]);
    msg->data.push_back(d);
    pos += lens[i];
  }
  return true;
}

// The listening socket is reachable by anything that can route to this host (port scanners, health checks, a stale process of another
// job), so nothing read from a connection is trusted: a frame with the wrong magic, an implausible part count / length or a meta block
// that does not parse makes RecvFrame return false, and the caller closes THAT connection and carries on (the ZeroMQ transport of the
// reference tolerates foreign connections the same way).  Accepted sockets carry a receive timeout (PS_RECV_TIMEOUT_MS), so a peer that stalls
// in the middle of a frame cannot hold the receive thread — and with it every other peer of the plane — forever.
bool Van::RecvFrame(int fd, Message* msg) {
  uint32_t h[3];
  if (!ReadAll(fd, h, 12)) return false;
  if (h[0] != kMagic) { HIPS_VLOG(1, "plane %d: dropping a connection that does not speak the HiPS framing (magic %08x)", plane_, h[0]); return false; }
  const uint32_t ml = h[1], nd = h[2];
  if (ml > kMaxMetaBytes || nd > kMaxFrameParts) { HIPS_VLOG(1, "plane %d: implausible frame header (meta %u B, %u parts)", plane_, ml, nd); return false; }
  std::vector<uint64_t> lens(nd);
  if (nd && !ReadAll(fd, lens.data(), 8 * nd)) return false;
  uint64_t total = 0;
  for (uint32_t i = 0; i < nd; ++i) {
    if (lens[i] > max_msg_bytes_ || (total += lens[i]) > max_msg_bytes_) { HIPS_VLOG(1, "plane %d: frame larger than PS_MAX_MSG_BYTES", plane_); return false; }
  }
  std::vector<char> meta(ml);
  if (!ReadAll(fd, meta.data(), ml)) return false;
  try {
    UnpackMeta(meta.data(), ml, &msg->meta);
  } catch (const std::exception& e) {
    HIPS_VLOG(1, "plane %d: unparsable message meta (%s)", plane_, e.what());
    return false;
  }
  msg->data.clear();
  size_t bytes = ml;
  for (uint32_t i = 0; i < nd; ++i) {
    SArray<char> d;
    if (lens[i]) {
      try { d.Allocate(lens[i]); } catch (const std::exception&) { return false; }      // pooled above 64 KiB (block_pool.h)
      if (!ReadAll(fd, d.data(), lens[i])) return false;
    }
    msg->data.push_back(d);
    bytes += lens[i];
  }
  recv_bytes_ += bytes;
  return true;
}

// ------------------------------------------------------------------------------------------------ lifecycle
void Van::Start(int customer_id) {
  Environment* env = Environment::Get();
  const char* uri_key = plane_ == kLocal ? "DMLC_PS_ROOT_URI" : "DMLC_PS_GLOBAL_ROOT_URI";
  const char* port_key = plane_ == kLocal ? "DMLC_PS_ROOT_PORT" : "DMLC_PS_GLOBAL_ROOT_PORT";
  scheduler_.hostname = env->GetStr(uri_key, "127.0.0.1");
  scheduler_.port = env->GetInt(port_key, plane_ == kLocal ? 9091 : 9092);
  scheduler_.role = Node::SCHEDULER;
  scheduler_.id = kScheduler;
  const int role = po_->role_in(plane_);
  is_scheduler_ = role == Node::SCHEDULER;
  enable_p3_ = env->GetInt("ENABLE_P3", 0) != 0;
  drop_rate_ = env->GetInt("PS_DROP_MSG", 0);
  max_msg_bytes_ = static_cast<uint64_t>(env->GetFloat("PS_MAX_MSG_BYTES", 17179869184.0));   // 16 GiB: above any single tensor message
  recv_timeout_ms_ = env->GetInt("PS_RECV_TIMEOUT_MS", 60000);
  heartbeat_timeout_ = env->GetInt("PS_HEARTBEAT_TIMEOUT", 0);
  barrier_count_.assign(8, 0);

  if (is_scheduler_) {
    my_node_ = scheduler_;
  } else {
    my_node_.role = role;
    std::string ip = env->GetStr("DMLC_NODE_HOST", "");
    if (ip.empty()) {
      std::string itf = env->GetStr("DMLC_INTERFACE", "");
      if (!itf.empty()) GetIP(itf, &ip);
      if (ip.empty()) { std::string dummy; GetAvailableInterfaceAndIP(&dummy, &ip); }
      if (ip.empty() || scheduler_.hostname == "127.0.0.1" || scheduler_.hostname == "localhost") ip = "127.0.0.1";
    }
    my_node_.hostname = ip;
    int port = (plane_ == kLocal) ? env->GetInt("PORT", 0) : 0;
    my_node_.port = port ? port : GetAvailablePort();
    my_node_.id = Node::kEmpty;
    my_node_.customer_id = customer_id;
    my_node_.rank_hint = rank_hint_;
  }
  if (plane_ == kGlobal && env->GetInt("ENABLE_DGT", 0) == 1 && !is_scheduler_) {
    // DGT mode 1: unimportant gradient blocks travel as datagrams (lossy by design), one socket, per-message IP_TOS
    udp_fd_ = ::socket(AF_INET, SOCK_DGRAM, 0);
    if (udp_fd_ >= 0) {
      sockaddr_in a; memset(&a, 0, sizeof(a));
      a.sin_family = AF_INET; a.sin_addr.s_addr = htonl(INADDR_ANY); a.sin_port = 0;
      int rcvbuf = 8 << 20; ::setsockopt(udp_fd_, SOL_SOCKET, SO_RCVBUF, &rcvbuf, sizeof(rcvbuf));
      socklen_t al = sizeof(a);
      if (::bind(udp_fd_, reinterpret_cast<sockaddr*>(&a), sizeof(a)) == 0 && ::getsockname(udp_fd_, reinterpret_cast<sockaddr*>(&a), &al) == 0) {
        my_node_.udp_port = ntohs(a.sin_port);
      } else { ::close(udp_fd_); udp_fd_ = -1; }
    }
  }
  my_node_.port = Bind(&my_node_, is_scheduler_ ? 0 : 40);
  HIPS_CHECK_MSG(my_node_.port > 0, "bind failed");
  HIPS_VLOG(1, "plane %d bind to %s", plane_, my_node_.DebugString().c_str());
  HIPS_CHECK(::pipe(wake_pipe_) == 0);
  {
    std::lock_guard<std::mutex> lk(nodes_mu_);
    nodes_[kScheduler] = scheduler_;
  }
  // everything the receive thread consults must exist BEFORE it starts: a peer's first barrier request can arrive the moment this node
  // turns ready, i.e. before Start() returns (a resender created later would neither ACK nor de-duplicate that first message)
  if (env->GetInt("PS_RESEND", 0) != 0) std::atomic_store(&resender_, std::make_shared<Resender>(env->GetInt("PS_RESEND_TIMEOUT", 1000), 100, this));
  const bool ts_on = plane_ == kLocal ? env->GetInt("ENABLE_INTRA_TS", 0) != 0 : env->GetInt("ENABLE_INTER_TS", 0) != 0;
  if (is_scheduler_ && ts_on) ts_sched_.reset(new TSScheduler(this, po_->num_workers_in(plane_), plane_));
  if (plane_ == kGlobal && env->GetInt("ENABLE_DGT", 0) != 0) {
    dgt_sender_.reset(new DGTSender(this));
    dgt_receiver_.reset(new DGTReceiver());
  }
  accept_thread_.reset(new std::thread(&Van::Accepting, this));
  recv_thread_.reset(new std::thread(&Van::Receiving, this));
  if (udp_fd_ >= 0) udp_thread_.reset(new std::thread(&Van::ReceivingUDP, this));
  if (enable_p3_) prio_thread_.reset(new std::thread(&Van::PrioritySending, this));
  if (plane_ == kGlobal && env->GetInt("GEOMX_EMULATE_DELAY_MS", 0) > 0) {
    emulate_delay_us_ = env->GetInt("GEOMX_EMULATE_DELAY_MS", 0) * 1000;
    delay_thread_.reset(new std::thread(&Van::DelayedSending, this));
  }

  if (!is_scheduler_) {
    Message msg;
    msg.meta.recver = kScheduler;
    msg.meta.control.cmd = Control::ADD_NODE;
    msg.meta.control.node.push_back(my_node_);
    msg.meta.timestamp = timestamp_++;
    SendNow(msg);
  }
  while (!ready_.load() && !stop_.load()) std::this_thread::sleep_for(std::chrono::milliseconds(5));

  if (!is_scheduler_ && env->GetInt("PS_HEARTBEAT_INTERVAL", 0) > 0) heartbeat_thread_.reset(new std::thread(&Van::Heartbeat, this));
}

void Van::Stop() {
  if (std::shared_ptr<Resender> rs = std::atomic_load(&resender_)) { if (!stop_.load()) rs->WaitDrained(5000); }
  if (stop_.exchange(true)) return;
  if (prio_thread_) {
    Message exit; exit.meta.control.cmd = Control::TERMINATE; exit.meta.recver = my_node_.id; exit.meta.priority = -(1 << 30);
    send_queue_.Push(exit);
    prio_thread_->join();
  }
  if (dgt_sender_) dgt_sender_->Stop();
  if (delay_thread_) { { std::lock_guard<std::mutex> lk(delay_mu_); } delay_cv_.notify_all(); delay_thread_->join(); }
  if (heartbeat_thread_) heartbeat_thread_->join();
  char c = 1;
  if (wake_pipe_[1] >= 0) { ssize_t r = ::write(wake_pipe_[1], &c, 1); (void)r; }
  // shutdown() wakes the accept thread; the descriptor itself is closed (and the member rewritten) only after that thread has gone — it reads
  // listen_fd_ on every iteration (found by ThreadSanitizer, profiles/tsan_hips.txt)
  if (listen_fd_ >= 0) ::shutdown(listen_fd_, SHUT_RDWR);
  if (accept_thread_) accept_thread_->join();
  if (recv_thread_) recv_thread_->join();
  if (udp_thread_) { udp_thread_->join(); ::close(udp_fd_); udp_fd_ = -1; }
  if (listen_fd_ >= 0) { ::close(listen_fd_); listen_fd_ = -1; }
  // the receive threads hand every message to the resender (ACK / duplicate filter) and its timer thread sends through senders_: it goes
  // away after the receivers and before the sender table
  std::atomic_store(&resender_, std::shared_ptr<Resender>());
  if (Environment::Get()->GetInt("GEOMX_NET_STATS", 0) != 0) {   // one machine-readable line per plane (tests, capacity planning)
    fprintf(stdout, "RESULT {\"net_stats\": {\"plane\": %d, \"node\": %d, \"sent_bytes\": %zu, \"recv_bytes\": %zu, \"udp_sent\": %zu, \"udp_received\": %zu}}\n",
            static_cast<int>(plane_), my_node_.id, send_bytes_.load(), recv_bytes_.load(), udp_sent_.load(), udp_received_.load());
    fflush(stdout);
  }
  {
    std::lock_guard<std::mutex> lk(senders_mu_);
    for (auto& s : senders_) if (s.second->fd >= 0) ::close(s.second->fd);
    senders_.clear();
  }
  {
    std::lock_guard<std::mutex> lk(fds_mu_);
    for (int fd : recv_fds_) ::close(fd);
    recv_fds_.clear();
  }
  if (!unix_path_.empty()) ::unlink(unix_path_.c_str());
  ready_ = false;
}

// ------------------------------------------------------------------------------------------------ send
int Van::Send(const Message& msg) {
  if (enable_p3_ && msg.meta.control.empty() && !msg.meta.simple_app && msg.meta.request && msg.meta.push) {
    send_queue_.Push(msg);  // P3: data pushes are drained in priority order by the sender thread (van.cc:847-860)
    return 0;
  }
  return SendNow(msg);
}

void Van::PrioritySending() {
  while (true) {
    Message msg;
    send_queue_.WaitAndPop(&msg);
    if (msg.meta.control.cmd == Control::TERMINATE && msg.meta.recver == my_node_.id) break;
    SendNow(msg);
  }
}

static int64_t NowUs() { return std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int Van::SendNow(const Message& msg) {
  if (emulate_delay_us_ > 0 && msg.meta.control.empty() && ready_.load()) {
    std::lock_guard<std::mutex> lk(delay_mu_);
    delay_q_.emplace_back(NowUs() + emulate_delay_us_, msg);
    delay_cv_.notify_one();
    return 0;
  }
  return SendWire(msg);
}

void Van::DelayedSending() {
  std::unique_lock<std::mutex> lk(delay_mu_);
  while (true) {
    delay_cv_.wait(lk, [this] { return stop_.load() || !delay_q_.empty(); });
    if (delay_q_.empty()) { if (stop_.load()) return; continue; }
    const int64_t due = delay_q_.front().first, now = NowUs();
    if (now < due && !stop_.load()) { delay_cv_.wait_for(lk, std::chrono::microseconds(due - now)); continue; }
    Message m = std::move(delay_q_.front().second);
    delay_q_.pop_front();
    lk.unlock();
    SendWire(m);
    lk.lock();
  }
}

int Van::SendWire(const Message& msg) {
  const int id = msg.meta.recver;
  HIPS_CHECK(id != Meta::kEmpty);
  if (stop_.load()) return -1;                 // the transport is going down: late application traffic (e.g. a last response) is dropped
  std::shared_ptr<Sender> s;
  {
    std::lock_guard<std::mutex> lk(senders_mu_);
    auto it = senders_.find(id);
    if (it == senders_.end()) it = senders_.emplace(id, std::make_shared<Sender>()).first;
    s = it->second;
  }
  Message out = msg;
  if (out.meta.sender == Meta::kEmpty) out.meta.sender = my_node_.id;
  std::lock_guard<std::mutex> lk(s->mu);
  if (s->fd < 0) {
    Node peer;
    {
      std::lock_guard<std::mutex> nl(nodes_mu_);
      auto it = nodes_.find(id);
      HIPS_CHECK_MSG(it != nodes_.end(), "unknown node id " + std::to_string(id) + " in plane " + std::to_string(plane_));
      peer = it->second;
    }
    s->fd = ConnectFd(peer);
    if (s->fd < 0) return -1;
  }
  int n = SendFrame(s->fd, out);
  if (n < 0) {  // one reconnect attempt (peer restarted / recovery)
    ::close(s->fd);
    Node peer;
    { std::lock_guard<std::mutex> nl(nodes_mu_); peer = nodes_[id]; }
    s->fd = ConnectFd(peer);
    if (s->fd < 0) return -1;
    n = SendFrame(s->fd, out);
  }
  if (n >= 0) {
    send_bytes_ += n;
    if (ready_.load() && out.meta.control.cmd != Control::ACK && out.meta.control.cmd != Control::ADD_NODE) {
      if (std::shared_ptr<Resender> rs = std::atomic_load(&resender_)) rs->AddOutgoing(out);
    }
    if (Verbose() >= 2) HIPS_VLOG(2, "plane %d SEND %d -> %d cmd=%d req=%d push=%d ts=%d bytes=%d", plane_, out.meta.sender, id, out.meta.control.cmd,
                                   out.meta.request, out.meta.push, out.meta.timestamp, n);
  }
  return n;
}

// ------------------------------------------------------------------------------------------------ receive
void Van::Accepting() {
  while (!stop_) {
    int fd = ::accept(listen_fd_, nullptr, nullptr);
    if (fd < 0) {
      if (stop_) break;
      if (errno == EINTR) continue;
      break;
    }
    int one = 1;
    setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
    if (recv_timeout_ms_ > 0) {      // bounds a read() in the middle of a frame; idle connections are only read after poll() reports data
      struct timeval tv; tv.tv_sec = recv_timeout_ms_ / 1000; tv.tv_usec = (recv_timeout_ms_ % 1000) * 1000;
      setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
    }
    {
      std::lock_guard<std::mutex> lk(fds_mu_);
      recv_fds_.push_back(fd);
    }
    char c = 0;
    ssize_t r = ::write(wake_pipe_[1], &c, 1); (void)r;
  }
}

void Van::Receiving() {
  std::vector<Node> nodes, recovery_nodes;
  std::mt19937 rng(12345u + my_node_.port);
  while (!stop_) {
    std::vector<struct pollfd> pfds;
    {
      std::lock_guard<std::mutex> lk(fds_mu_);
      pfds.resize(recv_fds_.size() + 1);
      pfds[0].fd = wake_pipe_[0]; pfds[0].events = POLLIN;
      for (size_t i = 0; i < recv_fds_.size(); ++i) { pfds[i + 1].fd = recv_fds_[i]; pfds[i + 1].events = POLLIN; }
    }
    int rc = ::poll(pfds.data(), pfds.size(), 500);
    if (rc <= 0) continue;
    if (pfds[0].revents & POLLIN) { char buf[64]; ssize_t r = ::read(wake_pipe_[0], buf, sizeof(buf)); (void)r; }
    for (size_t i = 1; i < pfds.size(); ++i) {
      if (!(pfds[i].revents & (POLLIN | POLLHUP | POLLERR))) continue;
      Message msg;
      bool got = false;
      try { got = RecvFrame(pfds[i].fd, &msg); } catch (const std::exception& e) { HIPS_VLOG(1, "plane %d: receive error (%s)", plane_, e.what()); }
      if (!got) {  // peer closed, stalled mid-frame, or not a HiPS peer at all: drop this connection only
        std::lock_guard<std::mutex> lk(fds_mu_);
        ::close(pfds[i].fd);
        recv_fds_.erase(std::remove(recv_fds_.begin(), recv_fds_.end(), pfds[i].fd), recv_fds_.end());
        continue;
      }
      if (Verbose() >= 2) HIPS_VLOG(2, "plane %d RECV %d <- %d cmd=%d req=%d push=%d ts=%d", plane_, my_node_.id, msg.meta.sender, msg.meta.control.cmd,
                                     msg.meta.request, msg.meta.push, msg.meta.timestamp);
      // fault injection: drop received messages with probability PS_DROP_MSG % (only once the node is up)
      if (ready_.load() && drop_rate_ > 0 && msg.meta.control.cmd != Control::TERMINATE && static_cast<int>(rng() % 100) < drop_rate_) {
        HIPS_VLOG(1, "plane %d drop message from %d", plane_, msg.meta.sender);
        continue;
      }
      // registration traffic (before either side is ready) is not covered by the ACK protocol, exactly like the reference, which starts
      // its resender only once the node table is known
      if (ready_.load() && msg.meta.control.cmd != Control::ADD_NODE) {
        std::shared_ptr<Resender> rs = std::atomic_load(&resender_);
        if (rs && rs->AddIncoming(msg)) continue;
      }
      const Control& ctrl = msg.meta.control;
      if (!ctrl.empty()) {
        if (ctrl.cmd == Control::TERMINATE) { stop_ = true; break; }
        else if (ctrl.cmd == Control::ADD_NODE) ProcessAddNode(&msg, &nodes, &recovery_nodes);
        else if (ctrl.cmd == Control::BARRIER) ProcessBarrier(&msg);
        else if (ctrl.cmd == Control::HEARTBEAT) ProcessHeartbeat(&msg);
        else if (ctrl.cmd == Control::ASKPUSH || ctrl.cmd == Control::ASKPULL || ctrl.cmd == Control::AUTOPULLREPLY || ctrl.cmd == Control::REPLY) {
          if (is_scheduler_ && ts_sched_) ts_sched_->Process(msg);
          else ProcessData(&msg);  // scheduler decisions travel to the KV apps as control-tagged messages
        }
      } else {
        if (dgt_receiver_ && msg.meta.msg_type == 1) {  // DGT block: reassemble, deliver when the last block arrives
          DeliverDGT(&msg);
        } else {
          ProcessData(&msg);
        }
      }
    }
  }
}

void Van::DeliverDGT(Message* msg) {
  Message whole;
  {
    std::lock_guard<std::mutex> lk(deliver_mu_);      // blocks of one tensor arrive on two threads (TCP: important, UDP: the rest)
    if (!dgt_receiver_->Add(*msg, &whole)) return;
  }
  ProcessData(&whole);
}

void Van::ReceivingUDP() {
  std::vector<char> buf(65536);
  while (!stop_.load()) {
    struct pollfd pfd; pfd.fd = udp_fd_; pfd.events = POLLIN;
    if (::poll(&pfd, 1, 200) <= 0) continue;
    const ssize_t n = ::recvfrom(udp_fd_, buf.data(), buf.size(), 0, nullptr, nullptr);
    if (n <= 0) continue;
    Message msg;
    if (!UnpackDatagram(buf.data(), static_cast<size_t>(n), &msg)) continue;   // truncated / foreign datagram: drop, like a lost one
    ++udp_received_;
    recv_bytes_ += static_cast<size_t>(n);
    msg.meta.plane = plane_;
    if (dgt_receiver_ && msg.meta.msg_type == 1) DeliverDGT(&msg);
  }
}

int Van::SendUDP(const Message& msg) {
  Node peer;
  {
    std::lock_guard<std::mutex> nl(nodes_mu_);
    auto it = nodes_.find(msg.meta.recver);
    if (it != nodes_.end()) peer = it->second;
  }
  std::vector<char> pkt;
  Message out = msg;
  if (out.meta.sender == Meta::kEmpty) out.meta.sender = my_node_.id;
  PackDatagram(out, &pkt);
  if (udp_fd_ < 0 || peer.udp_port <= 0 || pkt.size() > 60000) return SendNow(msg);   // no datagram endpoint (or too large): reliable path
  sockaddr_in a; memset(&a, 0, sizeof(a));
  a.sin_family = AF_INET; a.sin_port = htons(static_cast<uint16_t>(peer.udp_port));
  if (::inet_pton(AF_INET, peer.hostname.c_str(), &a.sin_addr) != 1) return SendNow(msg);
  std::lock_guard<std::mutex> lk(udp_mu_);
  int tos = out.meta.tos & 0xFC;                                     // DSCP in the upper six bits (reference: ZMQ_TOS / iptables DSCP classes)
  ::setsockopt(udp_fd_, IPPROTO_IP, IP_TOS, &tos, sizeof(tos));
  const ssize_t n = ::sendto(udp_fd_, pkt.data(), pkt.size(), 0, reinterpret_cast<sockaddr*>(&a), sizeof(a));
  if (n < 0) return -1;
  ++udp_sent_;
  send_bytes_ += static_cast<size_t>(n);
  return static_cast<int>(n);
}

void Van::ProcessData(Message* msg) {
  HIPS_CHECK(msg->meta.app_id != Meta::kEmpty);
  const int app_id = msg->meta.app_id;
  const int customer_id = po_->role_in(plane_) == Node::WORKER && plane_ == kLocal ? msg->meta.customer_id : app_id;
  msg->meta.plane = plane_;
  if (!po_->DeliverTo(app_id, customer_id, *msg, po_->is_finalizing() ? 0 : 5)) {
    // the application is gone (shutdown) — e.g. a retransmission whose original was already answered; never throw on the receive thread
    HIPS_VLOG(1, "plane %d drop message for missing customer app=%d customer=%d from %d", plane_, app_id, customer_id, msg->meta.sender);
  }
}

void Van::ProcessHeartbeat(Message* msg) {
  const time_t t = time(nullptr);
  for (const auto& n : msg->meta.control.node) {
    po_->UpdateHeartbeat(n.id, t, plane_);
    if (is_scheduler_) {  // echo
      Message ack;
      ack.meta.recver = n.id;
      ack.meta.control.cmd = Control::HEARTBEAT;
      ack.meta.control.node.push_back(my_node_);
      ack.meta.timestamp = timestamp_++;
      SendNow(ack);
    }
  }
}

void Van::Heartbeat() {
  const int interval = Environment::Get()->GetInt("PS_HEARTBEAT_INTERVAL", 0);
  int waited = 0;
  while (interval > 0 && !stop_) {
    std::this_thread::sleep_for(std::chrono::milliseconds(100));
    waited += 100;
    if (waited < interval * 1000) continue;
    waited = 0;
    if (!ready_) continue;
    Message msg;
    msg.meta.recver = kScheduler;
    msg.meta.control.cmd = Control::HEARTBEAT;
    msg.meta.control.node.push_back(my_node_);
    msg.meta.timestamp = timestamp_++;
    SendNow(msg);
  }
}

void Van::ProcessBarrier(Message* msg) {
  auto& ctrl = msg->meta.control;
  if (msg->meta.request) {
    const int group = ctrl.barrier_group;
    if (barrier_count_.size() <= static_cast<size_t>(group)) barrier_count_.resize(group + 1, 0);
    ++barrier_count_[group];
    const auto& ids = po_->GetNodeIDs(group, plane_);
    HIPS_VLOG(1, "plane %d barrier count for group %d: %d / %zu", plane_, group, barrier_count_[group], ids.size());
    if (barrier_count_[group] == static_cast<int>(ids.size())) {
      barrier_count_[group] = 0;
      Message res;
      res.meta.request = false;
      res.meta.app_id = msg->meta.app_id;
      res.meta.customer_id = msg->meta.customer_id;
      res.meta.control.cmd = Control::BARRIER;
      for (int r : ids) {
        if (r == my_node_.id) { po_->Manage(res, plane_); continue; }
        res.meta.recver = r;
        res.meta.timestamp = timestamp_++;
        SendNow(res);
      }
    }
  } else {
    po_->Manage(*msg, plane_);
  }
}

// scheduler side of the registration (reference van.cc:41-163): collect all nodes, assign ranks/ids, broadcast the table.
void Van::ProcessAddNodeAtScheduler(Message* msg, std::vector<Node>* nodes, std::vector<Node>* recovery_nodes) {
  recovery_nodes->clear();
  const size_t num_nodes = po_->num_servers_in(plane_) + po_->num_workers_in(plane_);
  const time_t t = time(nullptr);
  auto& ctrl = msg->meta.control;
  if (nodes->size() < num_nodes) {
    nodes->push_back(ctrl.node[0]);
    if (nodes->size() < num_nodes) return;
    // all registered: deterministic order, then assign ranks
    // A strict weak order (role, requested rank, host, port): ranks are counted per role, so hinted servers come out in the order of
    // their hints.  (Comparing hints only between hinted nodes and addresses otherwise is NOT transitive once an un-hinted worker sits
    // between two hinted servers — std::sort then returned the servers in either order and MultiGPS key ownership diverged between planes.)
    std::sort(nodes->begin(), nodes->end(), [](const Node& a, const Node& b) {
      if (a.role != b.role) return a.role < b.role;
      const int ha = a.rank_hint >= 0 ? a.rank_hint : 1 << 30, hb = b.rank_hint >= 0 ? b.rank_hint : 1 << 30;
      if (ha != hb) return ha < hb;
      return a.hostname != b.hostname ? a.hostname < b.hostname : a.port < b.port;
    });
    for (auto& node : *nodes) {
      const int id = node.role == Node::SERVER ? ServerRankToID(num_servers_seen_, plane_) : WorkerRankToID(num_workers_seen_, plane_);
      node.id = id;
      if (node.role == Node::SERVER) ++num_servers_seen_; else ++num_workers_seen_;
      { std::lock_guard<std::mutex> lk(nodes_mu_); nodes_[id] = node; }
      po_->UpdateHeartbeat(id, t, plane_);
      HIPS_VLOG(1, "plane %d assign id=%d to %s", plane_, id, node.DebugString().c_str());
    }
    nodes->push_back(my_node_);
    Message back;
    back.meta.control.cmd = Control::ADD_NODE;
    back.meta.control.node = *nodes;
    for (const auto& n : *nodes) {
      if (n.id == my_node_.id) continue;
      back.meta.recver = n.id;
      back.meta.timestamp = timestamp_++;
      SendNow(back);
    }
    HIPS_VLOG(1, "plane %d the scheduler is connected to %d workers and %d servers", plane_, num_workers_seen_, num_servers_seen_);
    ready_ = true;
  } else {
    // a node (re)registers after the cluster is full: hand it the id of a dead node of the same role (reference van.cc:176-192)
    auto dead = po_->GetDeadNodes(heartbeat_timeout_ > 0 ? heartbeat_timeout_ : 1, plane_);
    Node nn = ctrl.node[0];
    for (auto& old : *nodes) {
      if (old.role != nn.role || old.id == my_node_.id) continue;
      if (std::find(dead.begin(), dead.end(), old.id) == dead.end()) continue;
      nn.id = old.id; nn.is_recovery = true;
      old = nn;
      { std::lock_guard<std::mutex> lk(nodes_mu_); nodes_[nn.id] = nn; }
      { std::lock_guard<std::mutex> lk(senders_mu_); auto it = senders_.find(nn.id); if (it != senders_.end() && it->second->fd >= 0) { ::close(it->second->fd); it->second->fd = -1; } }
      po_->UpdateHeartbeat(nn.id, t, plane_);
      recovery_nodes->push_back(nn);
      HIPS_VLOG(1, "plane %d replace dead node %d by %s", plane_, nn.id, nn.DebugString().c_str());
      break;
    }
    if (recovery_nodes->empty()) return;
    for (const auto& n : *nodes) {
      if (n.id == my_node_.id) continue;
      Message back;
      back.meta.control.cmd = Control::ADD_NODE;
      back.meta.control.node = (n.id == nn.id) ? *nodes : *recovery_nodes;  // the newcomer gets the full table, the others the delta
      back.meta.recver = n.id;
      back.meta.timestamp = timestamp_++;
      SendNow(back);
    }
  }
}

void Van::UpdateLocalID(Message* msg, std::vector<Node>* recovery_nodes, const std::vector<Node>& nodes) {
  for (const auto& n : msg->meta.control.node) {
    if (my_node_.id == Node::kEmpty && n.hostname == my_node_.hostname && n.port == my_node_.port && n.role == my_node_.role) {
      my_node_ = n;
      Environment::Get()->Set(plane_ == kLocal ? "DMLC_RANK" : "DMLC_GLOBAL_RANK", std::to_string(IDtoRank(n.id, plane_)));
      HIPS_VLOG(1, "plane %d my id is %d%s", plane_, n.id, n.is_recovery ? " (recovery)" : "");
    }
  }
}

void Van::ProcessAddNode(Message* msg, std::vector<Node>* nodes, std::vector<Node>* recovery_nodes) {
  if (is_scheduler_) {
    ProcessAddNodeAtScheduler(msg, nodes, recovery_nodes);
    return;
  }
  UpdateLocalID(msg, recovery_nodes, *nodes);
  {
    std::lock_guard<std::mutex> lk(nodes_mu_);
    for (const auto& n : msg->meta.control.node) {
      if (n.id == Node::kEmpty) continue;
      const bool changed = nodes_.count(n.id) && (nodes_[n.id].port != n.port || nodes_[n.id].hostname != n.hostname);
      nodes_[n.id] = n;
      if (changed) {  // recovered peer: drop the stale socket, reconnect lazily
        std::lock_guard<std::mutex> sl(senders_mu_);
        auto it = senders_.find(n.id);
        if (it != senders_.end() && it->second->fd >= 0) { ::close(it->second->fd); it->second->fd = -1; }
      }
    }
  }
  ready_ = true;
}

}  // namespace hips
