// C predict API: a minimal, Python-free interface to run inference from a symbol JSON and a `.params` blob (predict.h is the runtime).
//
// Parity (GX prefix, same argument lists): include/mxnet/c_predict_api.h
//   MXPredCreate :78, MXPredCreatePartialOut :111, MXPredCreateMultiThread :144, MXPredReshape :170, MXPredGetOutputShape :185,
//   MXPredSetInput :198, MXPredForward :207, MXPredPartialForward :224, MXPredGetOutput :233, MXPredFree :242, MXNDList{Create,Get,Free} :252-277.
// dev_type 1 (cpu) runs here; dev_type 2 is refused with a message — device inference goes through the Python Executor on PyTorch tensors.
// Errors: -1 + GXRTGetLastError() (shared with c_api_runtime.cc, thread-local).
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "predict.h"

#define GX_CAPI extern "C" __attribute__((visibility("default")))

extern "C" const char* GXRTGetLastError();
void GXRTSetLastError(const std::string& msg);          // c_api_runtime.cc

namespace {
using gxrt::predict::NDList;
using gxrt::predict::Predictor;
using gxrt::predict::Shape;

template <typename F>
int Guard(F&& f) {
  try { f(); return 0; }
  catch (const std::exception& e) { GXRTSetLastError(e.what()); return -1; }
  catch (...) { GXRTSetLastError("unknown error"); return -1; }
}
struct Handle {
  std::unique_ptr<Predictor> pred;
  std::vector<uint32_t> shape_out;       // GetOutputShape hands out a pointer that stays valid until the next call on this handle
};
Handle* H(void* h) { if (!h) throw std::runtime_error("null predictor handle"); return static_cast<Handle*>(h); }

std::vector<Shape> Shapes(uint32_t n, const uint32_t* indptr, const uint32_t* data) {
  std::vector<Shape> out(n);
  for (uint32_t i = 0; i < n; ++i) {
    if (indptr[i + 1] < indptr[i] || indptr[i + 1] - indptr[i] > 8) throw std::runtime_error("input_shape_indptr is not a valid index pointer");
    out[i].assign(data + indptr[i], data + indptr[i + 1]);
  }
  return out;
}
std::unique_ptr<Predictor> Make(const char* json, const void* params, int param_size, int dev_type, uint32_t n_in, const char** keys,
                                const uint32_t* indptr, const uint32_t* shape_data, uint32_t n_out, const char** out_keys) {
  if (dev_type != 1) throw std::runtime_error("the native predictor runs on the host (dev_type 1); use the Python Executor for device inference");
  if (json == nullptr) throw std::runtime_error("null symbol JSON");
  if (param_size < 0) throw std::runtime_error("negative param_size");
  std::vector<std::string> ik, ok;
  for (uint32_t i = 0; i < n_in; ++i) ik.emplace_back(keys[i]);
  for (uint32_t i = 0; i < n_out; ++i) ok.emplace_back(out_keys[i]);
  return std::make_unique<Predictor>(std::string(json), static_cast<const char*>(params), static_cast<size_t>(param_size), ik, Shapes(n_in, indptr, shape_data), ok);
}
}  // namespace

GX_CAPI int GXPredCreate(const char* symbol_json, const void* param_bytes, int param_size, int dev_type, int /*dev_id*/, uint32_t num_input_nodes,
                         const char** input_keys, const uint32_t* input_shape_indptr, const uint32_t* input_shape_data, void** out) {
  return Guard([&] {
    auto h = std::make_unique<Handle>();
    h->pred = Make(symbol_json, param_bytes, param_size, dev_type, num_input_nodes, input_keys, input_shape_indptr, input_shape_data, 0, nullptr);
    *out = h.release();
  });
}
GX_CAPI int GXPredCreatePartialOut(const char* symbol_json, const void* param_bytes, int param_size, int dev_type, int /*dev_id*/, uint32_t num_input_nodes,
                                   const char** input_keys, const uint32_t* input_shape_indptr, const uint32_t* input_shape_data,
                                   uint32_t num_output_nodes, const char** output_keys, void** out) {
  return Guard([&] {
    auto h = std::make_unique<Handle>();
    h->pred = Make(symbol_json, param_bytes, param_size, dev_type, num_input_nodes, input_keys, input_shape_indptr, input_shape_data, num_output_nodes, output_keys);
    *out = h.release();
  });
}
// num_threads predictors over ONE copy of the graph and the parameters, each with its own inputs and activation arena
GX_CAPI int GXPredCreateMultiThread(const char* symbol_json, const void* param_bytes, int param_size, int dev_type, int /*dev_id*/, uint32_t num_input_nodes,
                                    const char** input_keys, const uint32_t* input_shape_indptr, const uint32_t* input_shape_data, int num_threads, void** out) {
  return Guard([&] {
    if (num_threads < 1) throw std::runtime_error("num_threads must be positive");
    std::vector<std::unique_ptr<Handle>> hs;
    hs.push_back(std::make_unique<Handle>());
    hs[0]->pred = Make(symbol_json, param_bytes, param_size, dev_type, num_input_nodes, input_keys, input_shape_indptr, input_shape_data, 0, nullptr);
    for (int i = 1; i < num_threads; ++i) { hs.push_back(std::make_unique<Handle>()); hs[i]->pred = hs[0]->pred->Clone(nullptr); }
    for (int i = 0; i < num_threads; ++i) out[i] = hs[i].release();
  });
}
// a NEW handle with other input shapes that shares the parameters of `handle` (which stays valid)
GX_CAPI int GXPredReshape(uint32_t num_input_nodes, const char** input_keys, const uint32_t* input_shape_indptr, const uint32_t* input_shape_data, void* handle, void** out) {
  return Guard([&] {
    std::map<std::string, Shape> shapes;
    const auto list = Shapes(num_input_nodes, input_shape_indptr, input_shape_data);
    for (uint32_t i = 0; i < num_input_nodes; ++i) shapes[input_keys[i]] = list[i];
    auto h = std::make_unique<Handle>();
    h->pred = H(handle)->pred->Clone(&shapes);
    *out = h.release();
  });
}
GX_CAPI int GXPredGetOutputShape(void* handle, uint32_t index, uint32_t** shape_data, uint32_t* shape_ndim) {
  return Guard([&] {
    Handle* h = H(handle);
    const Shape& s = h->pred->OutputShape(index);
    h->shape_out.assign(s.begin(), s.end());
    *shape_data = h->shape_out.data();
    *shape_ndim = static_cast<uint32_t>(h->shape_out.size());
  });
}
GX_CAPI int GXPredGetNumOutputs(void* handle, uint32_t* out) { return Guard([&] { *out = static_cast<uint32_t>(H(handle)->pred->NumOutputs()); }); }
GX_CAPI int GXPredSetInput(void* handle, const char* key, const float* data, uint32_t size) { return Guard([&] { H(handle)->pred->SetInput(key, data, size); }); }
GX_CAPI int GXPredForward(void* handle) { return Guard([&] { H(handle)->pred->Forward(); }); }
GX_CAPI int GXPredPartialForward(void* handle, int step, int* step_left) { return Guard([&] { H(handle)->pred->PartialForward(step, step_left); }); }
GX_CAPI int GXPredGetOutput(void* handle, uint32_t index, float* data, uint32_t size) { return Guard([&] { H(handle)->pred->GetOutput(index, data, size); }); }
// planner statistics: bytes of the activation arena and the number of operators that run
GX_CAPI int GXPredGetPlan(void* handle, uint64_t* arena_bytes, uint32_t* num_ops) {
  return Guard([&] { *arena_bytes = H(handle)->pred->ArenaBytes(); *num_ops = static_cast<uint32_t>(H(handle)->pred->NumOps()); });
}
GX_CAPI int GXPredFree(void* handle) { return Guard([&] { delete H(handle); }); }

GX_CAPI int GXNDListCreate(const char* nd_file_bytes, int nd_file_size, void** out, uint32_t* out_length) {
  return Guard([&] {
    if (nd_file_size < 0) throw std::runtime_error("negative nd_file_size");
    auto l = std::make_unique<NDList>(nd_file_bytes, static_cast<size_t>(nd_file_size));
    *out_length = static_cast<uint32_t>(l->data.size());
    *out = l.release();
  });
}
GX_CAPI int GXNDListGet(void* handle, uint32_t index, const char** out_key, const float** out_data, const uint32_t** out_shape, uint32_t* out_ndim) {
  return Guard([&] {
    if (!handle) throw std::runtime_error("null list handle");
    NDList* l = static_cast<NDList*>(handle);
    if (index >= l->data.size()) throw std::runtime_error("list index out of range");
    *out_key = l->names[index].c_str(); *out_data = l->data[index].data();
    *out_shape = l->shapes[index].data(); *out_ndim = static_cast<uint32_t>(l->shapes[index].size());
  });
}
GX_CAPI int GXNDListFree(void* handle) { return Guard([&] { delete static_cast<NDList*>(handle); }); }
