// Native reader/writer of MXNet's NDArray-list (.params) format.
// Parity: src/ndarray/ndarray.cc:1583-1811 (NDArray::Save/Load V2 magic 0xF993fac9, V1 0xF993fac8, legacy magic==ndim; list magic 0x112),
// dmlc serializer of vector<string> names.  Dense arrays are returned as (dtype_flag, shape, bytes); row_sparse / csr are densified.
#pragma once
#ifndef GEOMX_NO_PYTHON   // the Python-free C library (lib/libgeomx_capi.so) compiles the runtime without the pybind11 bindings
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>
#endif

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace gxrt {
#ifndef GEOMX_NO_PYTHON   // the Python-free C library (lib/libgeomx_capi.so) compiles the runtime without the pybind11 bindings
namespace py = pybind11;
#endif

struct NDRec {
  int dtype = 0;
  std::vector<int64_t> shape;
  std::string data;
  int dev_type = 1, dev_id = 0;
};

static const uint64_t kListMagic = 0x112;
static const uint32_t kV2Magic = 0xF993fac9, kV1Magic = 0xF993fac8;
inline int FlagSize(int f) { switch (f) { case 0: case 4: return 4; case 1: case 6: return 8; case 2: return 2; default: return 1; } }

class BufReader {
 public:
  BufReader(const char* p, size_t n) : p_(p), n_(n) {}
  template <typename T> T Get() { if (o_ + sizeof(T) > n_) throw std::runtime_error("truncated NDArray file"); T v; memcpy(&v, p_ + o_, sizeof(T)); o_ += sizeof(T); return v; }
  std::string Raw(size_t n) { if (n > n_ - o_) throw std::runtime_error("truncated NDArray file"); std::string s(p_ + o_, n); o_ += n; return s; }
  size_t remaining() const { return n_ - o_; }
 private:
  const char* p_; size_t n_, o_ = 0;
};

// Files are untrusted input: ranks, extents and index arrays are validated before anything is allocated or copied.
static const uint32_t kMaxRank = 32;
inline std::vector<int64_t> ReadShape64(BufReader& r) {
  const uint32_t nd = r.Get<uint32_t>();
  if (nd > kMaxRank) throw std::runtime_error("NDArray file: implausible rank " + std::to_string(nd));
  std::vector<int64_t> s(nd);
  for (uint32_t i = 0; i < nd; ++i) { s[i] = r.Get<int64_t>(); if (s[i] < 0) throw std::runtime_error("NDArray file: negative extent"); }
  return s;
}
inline int64_t Prod(const std::vector<int64_t>& s) {
  int64_t p = 1;
  for (auto d : s) { if (d != 0 && p > (int64_t{1} << 46) / d) throw std::runtime_error("NDArray file: implausible tensor size"); p *= d; }
  return p;
}
inline int CheckedFlagSize(int f) { if (f < 0 || f > 6) throw std::runtime_error("NDArray file: unknown dtype flag " + std::to_string(f)); return FlagSize(f); }

inline NDRec ReadArray(BufReader& r) {
  NDRec out;
  const uint32_t magic = r.Get<uint32_t>();
  if (magic == kV2Magic) {
    const int32_t stype = r.Get<int32_t>();
    if (stype < 0 || stype > 2) throw std::runtime_error("NDArray file: unknown storage type " + std::to_string(stype));
    const int nad = stype == 0 ? 0 : (stype == 1 ? 1 : 2);
    std::vector<int64_t> sshape;
    if (nad) sshape = ReadShape64(r);
    out.shape = ReadShape64(r);
    if (out.shape.empty()) return out;
    out.dev_type = r.Get<int32_t>(); out.dev_id = r.Get<int32_t>();
    out.dtype = r.Get<int32_t>();
    CheckedFlagSize(out.dtype);
    std::vector<std::pair<int, std::vector<int64_t>>> aux;
    for (int i = 0; i < nad; ++i) { const int af = r.Get<int32_t>(); CheckedFlagSize(af); aux.emplace_back(af, ReadShape64(r)); }
    const std::vector<int64_t>& dshape = nad ? sshape : out.shape;
    std::string data = r.Raw(static_cast<size_t>(Prod(dshape)) * FlagSize(out.dtype));
    std::vector<std::string> auxd;
    for (auto& a : aux) auxd.push_back(r.Raw(static_cast<size_t>(Prod(a.second)) * FlagSize(a.first)));
    if (stype == 0) { out.data.swap(data); return out; }
    const size_t es = FlagSize(out.dtype);
    if (static_cast<size_t>(Prod(out.shape)) * es > (size_t{1} << 34)) throw std::runtime_error("NDArray file: sparse array too large to densify");
    out.data.assign(static_cast<size_t>(Prod(out.shape)) * es, 0);
    for (auto& a : aux) if (a.first != 6) throw std::runtime_error("NDArray file: sparse index arrays must be int64");
    if (stype == 1) {  // row_sparse: aux0 = row indices (int64)
      const int64_t rows = aux[0].second.empty() ? 0 : aux[0].second[0];
      const size_t row_bytes = out.shape.size() > 1 ? static_cast<size_t>(Prod(out.shape) / std::max<int64_t>(out.shape[0], 1)) * es : es;
      if (static_cast<size_t>(rows) * row_bytes > data.size()) throw std::runtime_error("NDArray file: row_sparse data shorter than its index");
      for (int64_t i = 0; i < rows; ++i) {
        int64_t idx; memcpy(&idx, auxd[0].data() + i * 8, 8);
        if (idx < 0 || idx >= out.shape[0]) throw std::runtime_error("NDArray file: row index out of range");
        memcpy(&out.data[idx * row_bytes], data.data() + i * row_bytes, row_bytes);
      }
    } else {           // csr: aux0 = indptr, aux1 = indices (int64)
      if (out.shape.size() != 2) throw std::runtime_error("NDArray file: csr array must be 2-D");
      const int64_t nrow = out.shape[0], ncol = out.shape[1];
      const int64_t nnz = static_cast<int64_t>(std::min(auxd[1].size() / 8, data.size() / es));
      if (static_cast<int64_t>(auxd[0].size() / 8) < nrow + 1) throw std::runtime_error("NDArray file: csr indptr shorter than rows + 1");
      for (int64_t rr = 0; rr < nrow; ++rr) {
        int64_t s, e; memcpy(&s, auxd[0].data() + rr * 8, 8); memcpy(&e, auxd[0].data() + (rr + 1) * 8, 8);
        if (s < 0 || e < s || e > nnz) throw std::runtime_error("NDArray file: csr indptr out of range");
        for (int64_t j = s; j < e; ++j) {
          int64_t c; memcpy(&c, auxd[1].data() + j * 8, 8);
          if (c < 0 || c >= ncol) throw std::runtime_error("NDArray file: csr column index out of range");
          memcpy(&out.data[(rr * ncol + c) * es], data.data() + j * es, es);
        }
      }
    }
    return out;
  }
  if (magic == kV1Magic) out.shape = ReadShape64(r);
  else {                                       // legacy: magic is ndim
    if (magic > kMaxRank) throw std::runtime_error("NDArray file: bad magic / implausible rank");
    out.shape.resize(magic); for (uint32_t i = 0; i < magic; ++i) out.shape[i] = r.Get<uint32_t>();
  }
  if (out.shape.empty()) return out;
  out.dev_type = r.Get<int32_t>(); out.dev_id = r.Get<int32_t>();
  out.dtype = r.Get<int32_t>();
  CheckedFlagSize(out.dtype);
  out.data = r.Raw(static_cast<size_t>(Prod(out.shape)) * FlagSize(out.dtype));
  return out;
}

inline std::string WriteList(const std::vector<NDRec>& arrays, const std::vector<std::string>& names) {
  std::string out;
  auto put = [&out](const void* p, size_t n) { out.append(static_cast<const char*>(p), n); };
  uint64_t hdr[3] = {kListMagic, 0, arrays.size()};
  put(hdr, 24);
  for (const auto& a : arrays) {
    uint32_t magic = kV2Magic; int32_t stype = 0;
    put(&magic, 4); put(&stype, 4);
    uint32_t nd = static_cast<uint32_t>(a.shape.size()); put(&nd, 4);
    for (auto d : a.shape) put(&d, 8);
    if (a.shape.empty()) continue;
    int32_t ctx[2] = {a.dev_type, a.dev_id}; put(ctx, 8);
    int32_t fl = a.dtype; put(&fl, 4);
    put(a.data.data(), a.data.size());
  }
  uint64_t m = names.size(); put(&m, 8);
  for (const auto& n : names) { uint64_t l = n.size(); put(&l, 8); put(n.data(), l); }
  return out;
}

#ifndef GEOMX_NO_PYTHON   // the Python-free C library (lib/libgeomx_capi.so) compiles the runtime without the pybind11 bindings
inline void BindParamsIO(py::module_& m) {
  m.def("params_load", [](py::bytes blob) {
    std::string s = blob;
    BufReader r(s.data(), s.size());
    if (r.Get<uint64_t>() != kListMagic) throw std::runtime_error("Invalid NDArray file format");
    r.Get<uint64_t>();
    const uint64_t n = r.Get<uint64_t>();
    py::list arrays;
    for (uint64_t i = 0; i < n; ++i) {
      NDRec a = ReadArray(r);
      arrays.append(py::make_tuple(a.dtype, a.shape, py::bytes(a.data), a.dev_type, a.dev_id));
    }
    const uint64_t mnames = r.Get<uint64_t>();
    std::vector<std::string> names;
    for (uint64_t i = 0; i < mnames; ++i) { const uint64_t l = r.Get<uint64_t>(); names.push_back(r.Raw(l)); }
    return py::make_tuple(arrays, names);
  }, "parse an MXNet .params blob -> ([(dtype_flag, shape, bytes, dev_type, dev_id)], [names])");
  m.def("params_save", [](const std::vector<std::tuple<int, std::vector<int64_t>, py::bytes, int, int>>& arrays, const std::vector<std::string>& names) {
    std::vector<NDRec> recs;
    for (auto& t : arrays) { NDRec a; a.dtype = std::get<0>(t); a.shape = std::get<1>(t); a.data = std::get<2>(t); a.dev_type = std::get<3>(t); a.dev_id = std::get<4>(t); recs.push_back(a); }
    return py::bytes(WriteList(recs, names));
  });
}
#endif

}  // namespace gxrt
