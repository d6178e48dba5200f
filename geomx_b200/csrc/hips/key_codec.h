// Key -> (server, ps_key, length) mapping shared by workers and servers.
// Parity: KVStoreDist::EncodeDefaultKey (src/kvstore/kvstore_dist.h:721-761) and the server-side mirror (kvstore_dist_server.h:1770-1810):
// arrays smaller than MXNET_KVSTORE_BIGARRAY_BOUND go to ONE server chosen by (key*9973) % num_servers; larger arrays are partitioned
// evenly over ALL servers of the plane (MultiGPS load balancing); ps_key = server_key_range.begin() + key.
#pragma once
#include <cmath>
#include <vector>

#include "postoffice.h"

namespace hips {

struct PSKVPlan {
  std::vector<Key> keys;
  std::vector<int> lens;   // bytes per part
  std::vector<int> server; // server rank per part
  size_t size = 0;         // total bytes
};

inline PSKVPlan EncodeKey(Plane plane, int key, size_t num_elems, int num_bytes, size_t bigarray_bound) {
  const auto& krs = Postoffice::Get()->GetServerKeyRanges(plane);
  const int n = static_cast<int>(krs.size());
  PSKVPlan p;
  if (num_elems < bigarray_bound || n <= 1) {
    const int server = (key * 9973) % n;
    p.keys.push_back(krs[server].begin() + static_cast<Key>(key));
    p.lens.push_back(static_cast<int>(num_elems * num_bytes));
    p.server.push_back(server);
    p.size = num_elems * num_bytes;
  } else {
    for (int i = 0; i < n; ++i) {
      const size_t part = static_cast<size_t>(std::round(static_cast<double>(num_elems) / n * (i + 1))) -
                          static_cast<size_t>(std::round(static_cast<double>(num_elems) / n * i));
      p.keys.push_back(krs[i].begin() + static_cast<Key>(key));
      p.lens.push_back(static_cast<int>(part * num_bytes));
      p.server.push_back(i);
      p.size += part * num_bytes;
    }
  }
  return p;
}

// Compressed keys are never partitioned: a Bi-Sparse / 2-bit payload describes the WHOLE tensor ([values | indices], packed 2-bit words), so
// init, push and pull of such a key must all address the one hashed server — otherwise a server that only holds a slice would scatter
// indices beyond it away (Bi-Sparse) or dequantise with the wrong length (2-bit).  (The reference splits 2-bit payloads per server with
// EncodeCompressedKey, kvstore_dist.h:811-897, and keeps Bi-Sparse on one server, kvstore_dist_server.h:1849; one rule for both here.)
inline bool CompressionPinsKey(int gc_type /*CompressionType*/, size_t num_elems, int dtype, size_t size_lower_bound) {
  // 1 = kTwoBit (fp32 keys only, dtype 0 = kFloat32), 2 = kBiSparse (keys >= MXNET_KVSTORE_SIZE_LOWER_BOUND)
  return (gc_type == 1 && dtype == 0) || (gc_type == 2 && num_elems >= size_lower_bound);
}
inline PSKVPlan EncodeKeyPlan(Plane plane, int key, size_t num_elems, int num_bytes, size_t bigarray_bound, bool pinned) {
  return EncodeKey(plane, key, num_elems, num_bytes, pinned ? static_cast<size_t>(-1) : bigarray_bound);
}

// P3: a tensor is cut into bigarray_bound-element slices with globally increasing slice keys assigned round-robin to servers
// (kvstore_dist.h:763-799); returns (slice_key, server, begin_elem, num_elems)
struct P3Slice { int slice_key; int server; size_t begin, elems; };
inline std::vector<P3Slice> EncodeP3(int* next_slice_key, size_t num_elems, size_t slice_elems, int num_servers) {
  std::vector<P3Slice> out;
  for (size_t b = 0; b < num_elems; b += slice_elems) {
    const int sk = (*next_slice_key)++;
    out.push_back(P3Slice{sk, sk % num_servers, b, std::min(slice_elems, num_elems - b)});
  }
  return out;
}

}  // namespace hips
