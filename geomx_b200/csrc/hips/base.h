// HiPS runtime — common definitions.
//
// This is a from-scratch C++17 implementation of the capabilities of GeoMX's modified ps-lite
// (reference: 3rdparty/ps-lite/{include/ps,src}): two independent rendezvous/transport planes (local party plane and
// global plane), scheduler-mediated registration, per-tier barriers, request tracking, key-range sharding and the
// DGT / TSEngine / P3 / resend / heartbeat features.  Differences by design: plain POSIX TCP (or unix-domain) sockets with a
// length-prefixed binary codec instead of ZeroMQ + protobuf (neither is available offline, and the hot path of this framework is
// the NVSwitch fabric, csrc/kernels/hips_fabric.cu — this transport is the CPU / multi-host "plumbing" tier), ONE Van class
// instantiated per plane instead of duplicated code paths, and node-id arithmetic kept bit-compatible with the reference so that
// the server state machine can use the same sender-id predicates (postoffice.h:104-127, base.h:17-38 in the reference).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <stdexcept>
#include <string>

namespace hips {

using Key = uint64_t;
static const Key kMaxKey = UINT64_MAX;

// group ids (bit-ors), identical in both planes
static const int kScheduler = 1;
static const int kServerGroup = 2;
static const int kWorkerGroup = 4;

// planes
enum Plane { kLocal = 0, kGlobal = 1 };

// node-id arithmetic (reference postoffice.h:104-127):
//   local plane : scheduler 1, server rank r -> 100 + 2r... the reference uses 8 + 2r in BOTH planes for servers/workers of the
//   classic ps-lite and shifts the local plane by 92 so that `sender > 100` identifies a local worker/server and `sender < 100`
//   the global plane.  We keep exactly those observable properties:
//     local  : server r -> 100 + 2r (even), worker r -> 101 + 2r (odd)
//     global : global server r -> 8 + 2r (even), global worker (= local server of party r) -> 9 + 2r (odd)
inline int ServerRankToID(int rank, Plane p) { return (p == kLocal ? 100 : 8) + rank * 2; }
inline int WorkerRankToID(int rank, Plane p) { return (p == kLocal ? 101 : 9) + rank * 2; }
inline int IDtoRank(int id, Plane p) { return (id - (p == kLocal ? 100 : 8)) / 2; }
inline bool IsWorkerID(int id) { return id >= 8 && (id % 2) == 1; }
inline bool IsServerID(int id) { return id >= 8 && (id % 2) == 0; }

class Error : public std::runtime_error {
 public:
  explicit Error(const std::string& s) : std::runtime_error(s) {}
};

#define HIPS_CHECK(cond)                                                                         \
  do { if (!(cond)) throw ::hips::Error(std::string("Check failed: " #cond " at ") + __FILE__ + ":" + std::to_string(__LINE__)); } while (0)
#define HIPS_CHECK_MSG(cond, msg)                                                                \
  do { if (!(cond)) throw ::hips::Error(std::string("Check failed: " #cond " (") + (msg) + ") at " + __FILE__ + ":" + std::to_string(__LINE__)); } while (0)

int Verbose();  // PS_VERBOSE
#define HIPS_VLOG(level, ...)                      \
  do {                                             \
    if (::hips::Verbose() >= (level)) {            \
      fprintf(stderr, "[hips] " __VA_ARGS__);      \
      fputc('\n', stderr);                         \
    }                                              \
  } while (0)

}  // namespace hips
