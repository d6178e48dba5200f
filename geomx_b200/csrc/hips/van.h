// Van: the transport core of ONE plane (local party plane or global plane).
//
// Parity: ps-lite include/ps/internal/van.h:33-303 + src/van.cc (registration at the scheduler :41-163, barrier counting :259-288,
// heartbeat :242-257,1128-1140, P3 priority sender thread :847-860, PS_DROP_MSG fault injection :498-500,871-877, PS_RESEND hook
// :527-533) and src/zmq_van.h (ROUTER receiver + DEALER senders, multipart zero-copy send).  Here: POSIX TCP / unix-domain sockets
// (DMLC_LOCAL=1), one listening socket + poll()-based receiver thread, lazily connected per-peer send sockets with per-socket
// mutexes (the reference serialises ALL sends behind one ZMQ mutex, zmq_van.h:380), scatter-gather writev of [header|meta|data...]
// without copying the payload.  The reference duplicates every method for the global plane (Start/StartGlobal, Receiving/
// ReceivingGlobal ...); here a Van is instantiated once per plane.
#pragma once
#include <atomic>
#include <condition_variable>
#include <functional>
#include <map>
#include <memory>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "message.h"
#include "threadsafe_queue.h"

namespace hips {

class Postoffice;
class Resender;
class TSScheduler;
class DGTSender;
class DGTReceiver;

class Van {
 public:
  Van(Postoffice* po, Plane plane);
  ~Van();
  void Start(int customer_id);
  void set_rank_hint(int r) { rank_hint_ = r; }
  void Stop();
  // thread-safe; returns the number of payload bytes sent (-1 on failure). Honour P3 priority queue when enabled.
  int Send(const Message& msg);
  const Node& my_node() const { return my_node_; }
  bool IsReady() const { return ready_.load(); }
  Plane plane() const { return plane_; }
  int GetTimestamp() { return timestamp_++; }
  size_t send_bytes() const { return send_bytes_.load(); }
  size_t recv_bytes() const { return recv_bytes_.load(); }
  const Node& scheduler() const { return scheduler_; }
  TSScheduler* ts_scheduler() { return ts_sched_.get(); }
  DGTSender* dgt_sender() { return dgt_sender_.get(); }
  // direct (unqueued) transmit used by the priority sender thread, the resender and the DGT channel schedulers
  int SendNow(const Message& msg);
  int SendWire(const Message& msg);   // the actual socket write (SendNow may defer to it through the latency emulator)
  // lossy datagram path (DGT mode 1): one UDP packet per message with the message's TOS/DSCP; falls back to TCP when no endpoint is known
  int SendUDP(const Message& msg);
  size_t udp_sent() const { return udp_sent_.load(); }
  size_t udp_received() const { return udp_received_.load(); }

 private:
  int Bind(Node* node, int max_retry);
  int ConnectFd(const Node& node);
  int SendFrame(int fd, const Message& msg);
  bool RecvFrame(int fd, Message* msg);
  void Accepting();
  void Receiving();
  void PrioritySending();
  void ReceivingUDP();
  void DeliverDGT(Message* msg);   // reassembly + delivery of a DGT block (called from the TCP and the UDP receive threads)
  void Heartbeat();
  void ProcessAddNodeAtScheduler(Message* msg, std::vector<Node>* nodes, std::vector<Node>* recovery_nodes);
  void ProcessAddNode(Message* msg, std::vector<Node>* nodes, std::vector<Node>* recovery_nodes);
  void ProcessBarrier(Message* msg);
  void ProcessHeartbeat(Message* msg);
  void ProcessData(Message* msg);
  void UpdateLocalID(Message* msg, std::vector<Node>* recovery_nodes, const std::vector<Node>& nodes);

  Postoffice* po_;
  Plane plane_;
  Node scheduler_, my_node_;
  bool is_scheduler_ = false;
  std::atomic<bool> ready_{false}, stop_{false};
  std::atomic<size_t> send_bytes_{0}, recv_bytes_{0}, udp_sent_{0}, udp_received_{0};
  int udp_fd_ = -1;
  std::mutex udp_mu_, deliver_mu_;
  std::atomic<int> timestamp_{0};
  int listen_fd_ = -1;
  int wake_pipe_[2] = {-1, -1};
  bool use_unix_ = false;
  std::string unix_path_;

  std::mutex nodes_mu_;
  std::unordered_map<int, Node> nodes_;                 // id -> address
  struct Sender { int fd = -1; std::mutex mu; };
  std::unordered_map<int, std::shared_ptr<Sender>> senders_;   // id -> connected socket (shared: a sender in flight survives Stop())
  std::mutex senders_mu_;
  std::unordered_map<std::string, int> connected_;      // "host:port" -> node id (shared-address detection)

  std::mutex fds_mu_;
  std::vector<int> recv_fds_;

  std::unique_ptr<std::thread> accept_thread_, recv_thread_, heartbeat_thread_, prio_thread_, udp_thread_;
  ThreadsafeQueue<Message, MessagePriority> send_queue_;  // P3
  // WAN emulation (GEOMX_EMULATE_DELAY_MS, data messages of the global plane): sends are released by a timer thread `delay` after they
  // were issued, in order — a one-way latency on the link between parties without blocking the issuing thread
  int emulate_delay_us_ = 0;
  std::unique_ptr<std::thread> delay_thread_;
  std::mutex delay_mu_;
  std::condition_variable delay_cv_;
  std::deque<std::pair<int64_t, Message>> delay_q_;
  void DelayedSending();
  bool enable_p3_ = false;
  int drop_rate_ = 0;
  uint64_t max_msg_bytes_ = 1ull << 34;   // PS_MAX_MSG_BYTES: frames announcing more payload than this are rejected before any allocation
  int recv_timeout_ms_ = 60000;           // PS_RECV_TIMEOUT_MS: SO_RCVTIMEO of accepted connections (0 = block forever)
  int num_servers_seen_ = 0, num_workers_seen_ = 0;
  std::vector<int> barrier_count_;
  std::shared_ptr<Resender> resender_;      // read with std::atomic_load: application threads may still send while Stop() tears down
  std::unique_ptr<TSScheduler> ts_sched_;
  std::unique_ptr<DGTSender> dgt_sender_;
  std::unique_ptr<DGTReceiver> dgt_receiver_;
  int heartbeat_timeout_ = 0;
  int rank_hint_ = -1;
  friend class Resender;
  friend class DGTSender;
};

}  // namespace hips
