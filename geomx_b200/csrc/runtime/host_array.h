// The NDArray handle of the flat C ABI: host memory + shape + dtype flag, shared by c_api_runtime.cc (NDArray / serializer functions) and
// c_api_graph.cc (executor bindings, imperative invoke, autograd).  Reference role: the NDArrayHandle of include/mxnet/c_api.h:60 — there a
// pointer to a device-aware NDArray; here device tensors belong to PyTorch and the C ABI owns host arrays.
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "params_io.h"

namespace gxrt {
namespace capi {

struct AGNode;                          // autograd history of an array (c_api_graph.cc)

struct HostArray {
  gxrt::NDRec rec;
  std::vector<uint32_t> shape32;        // GetShape hands out a pointer that stays valid until the handle is freed
  std::shared_ptr<AGNode> ag;           // set while the array is a marked variable or the output of a recorded operator
  int ag_out = 0;                       // which output of that operator
  HostArray* grad = nullptr;            // marked variables: where Backward writes (not owned)
  int grad_req = 0;
  ~HostArray();
};

inline HostArray* ND(void* h) {
  if (h == nullptr) throw std::runtime_error("null NDArray handle");
  return static_cast<HostArray*>(h);
}

}  // namespace capi
}  // namespace gxrt
