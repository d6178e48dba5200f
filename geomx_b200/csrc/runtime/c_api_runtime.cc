// Plain C API of the native runtime for non-Python front ends: host NDArray handles with the byte-exact `.params` serializer, the profiler,
// the dependency engine and the pooled host storage.  Together with csrc/hips/c_api.cc (GXKVStore*) this is the flat C ABI of the framework.
//
// Parity (names follow the reference with the GX prefix): include/mxnet/c_api.h
//   NDArray   MXNDArrayCreateEx / Free / GetShape / GetDType / GetData / SyncCopyFromCPU / SyncCopyToCPU / Save / Load   (:540-1010)
//   Profiler  MXSetProfilerConfig / MXSetProfilerState / MXDumpProfile / MXProfilePause / MXProfileSetMarker            (src/c_api/c_api_profile.cc:264-560)
//   Engine    the push/wait contract of include/mxnet/engine.h:115-314 (NewVariable / PushAsync / WaitForVar / WaitForAll) for C callbacks
//   Storage   include/mxnet/storage.h Alloc / Free of the pooled host manager (src/storage/pooled_storage_manager.h:52-172)
// Device tensors belong to PyTorch in this design, so an NDArray handle here owns HOST memory; device data crosses this ABI
// through SyncCopy*.  Every function returns 0 on success and -1 on failure; GXRTGetLastError() describes the failure of the calling thread.
#include <cstdint>
#include <cstring>
#include <fstream>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "engine.h"
#include "host_array.h"
#include "params_io.h"
#include "profiler.h"
#include "storage.h"

#define GX_CAPI extern "C" __attribute__((visibility("default")))

namespace {
thread_local std::string rt_error;
template <typename F>
int Guard(F&& f) {
  try { f(); return 0; }
  catch (const std::exception& e) { rt_error = e.what(); return -1; }
  catch (...) { rt_error = "unknown error"; return -1; }
}
using gxrt::capi::HostArray;
using gxrt::capi::ND;
// results of the last GXNDArrayLoad of this thread (the reference returns pointers into thread-local storage as well, c_api.cc MXNDArrayLoad)
thread_local std::vector<void*> load_handles;
thread_local std::vector<std::string> load_names;
thread_local std::vector<const char*> load_name_ptrs;

gx_rt::PooledHostStorage& HostPool() { static gx_rt::PooledHostStorage pool; return pool; }
std::mutex engine_mu;
std::unique_ptr<gxrt::Engine> engine;
gxrt::Engine* Eng() {
  std::lock_guard<std::mutex> lk(engine_mu);
  if (!engine) engine.reset(new gxrt::Engine(4, false));
  return engine.get();
}
}  // namespace

GX_CAPI const char* GXRTGetLastError() { return rt_error.c_str(); }
void GXRTSetLastError(const std::string& msg) { rt_error = msg; }       // for the other translation units of the C ABI (c_predict_api.cc)

// ------------------------------------------------------------------------------------------------ NDArray (host)
// dtype: mshadow flags (0 f32, 1 f64, 2 f16, 3 u8, 4 i32, 5 i8, 6 i64)
GX_CAPI int GXNDArrayCreate(const uint32_t* shape, uint32_t ndim, int dtype, void** out) {
  return Guard([&] {
    auto a = std::make_unique<HostArray>();
    a->rec.dtype = dtype;
    a->rec.shape.assign(shape, shape + ndim);
    a->rec.data.assign(static_cast<size_t>(gxrt::Prod(a->rec.shape)) * gxrt::FlagSize(dtype), '\0');
    *out = a.release();
  });
}
GX_CAPI int GXNDArrayFree(void* h) { return Guard([&] { delete ND(h); }); }
GX_CAPI int GXNDArrayGetShape(void* h, uint32_t* out_ndim, const uint32_t** out_shape) {
  return Guard([&] {
    HostArray* a = ND(h);
    a->shape32.assign(a->rec.shape.begin(), a->rec.shape.end());
    *out_ndim = static_cast<uint32_t>(a->shape32.size());
    *out_shape = a->shape32.data();
  });
}
GX_CAPI int GXNDArrayGetDType(void* h, int* out) { return Guard([&] { *out = ND(h)->rec.dtype; }); }
GX_CAPI int GXNDArrayGetData(void* h, void** out) { return Guard([&] { *out = &ND(h)->rec.data[0]; }); }
GX_CAPI int GXNDArraySyncCopyFromCPU(void* h, const void* data, size_t size_elems) {
  return Guard([&] {
    HostArray* a = ND(h);
    const size_t bytes = size_elems * gxrt::FlagSize(a->rec.dtype);
    if (bytes != a->rec.data.size()) throw std::runtime_error("SyncCopyFromCPU: size does not match the array");
    memcpy(&a->rec.data[0], data, bytes);
  });
}
GX_CAPI int GXNDArraySyncCopyToCPU(void* h, void* data, size_t size_elems) {
  return Guard([&] {
    HostArray* a = ND(h);
    const size_t bytes = size_elems * gxrt::FlagSize(a->rec.dtype);
    if (bytes != a->rec.data.size()) throw std::runtime_error("SyncCopyToCPU: size does not match the array");
    memcpy(data, a->rec.data.data(), bytes);
  });
}
// `.params` / NDArray-list file, byte-compatible with NDArray::Save (src/ndarray/ndarray.cc:1583-1811); keys may be null (unnamed list)
GX_CAPI int GXNDArraySave(const char* fname, uint32_t num, void** handles, const char** keys) {
  return Guard([&] {
    std::vector<gxrt::NDRec> recs;
    std::vector<std::string> names;
    for (uint32_t i = 0; i < num; ++i) { recs.push_back(ND(handles[i])->rec); if (keys) names.emplace_back(keys[i]); }
    const std::string blob = gxrt::WriteList(recs, names);
    std::ofstream f(fname, std::ios::binary);
    if (!f) throw std::runtime_error(std::string("cannot open ") + fname);
    f.write(blob.data(), static_cast<std::streamsize>(blob.size()));
  });
}
GX_CAPI int GXNDArrayLoad(const char* fname, uint32_t* out_size, void*** out_handles, uint32_t* out_name_size, const char*** out_names) {
  return Guard([&] {
    std::ifstream f(fname, std::ios::binary);
    if (!f) throw std::runtime_error(std::string("cannot open ") + fname);
    std::string s((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    gxrt::BufReader r(s.data(), s.size());
    if (r.Get<uint64_t>() != gxrt::kListMagic) throw std::runtime_error("Invalid NDArray file format");
    r.Get<uint64_t>();
    const uint64_t n = r.Get<uint64_t>();
    load_handles.clear(); load_names.clear(); load_name_ptrs.clear();
    for (uint64_t i = 0; i < n; ++i) { auto a = std::make_unique<HostArray>(); a->rec = gxrt::ReadArray(r); load_handles.push_back(a.release()); }
    const uint64_t m = r.Get<uint64_t>();
    for (uint64_t i = 0; i < m; ++i) { const uint64_t l = r.Get<uint64_t>(); load_names.push_back(r.Raw(l)); }
    for (auto& nm : load_names) load_name_ptrs.push_back(nm.c_str());
    *out_size = static_cast<uint32_t>(load_handles.size()); *out_handles = load_handles.data();
    *out_name_size = static_cast<uint32_t>(load_name_ptrs.size()); *out_names = load_name_ptrs.data();
  });
}

// ------------------------------------------------------------------------------------------------ profiler
// keys: filename, aggregate_stats, continuous_dump, dump_period (the subset of MXSetProfilerConfig this profiler has knobs for)
GX_CAPI int GXSetProfilerConfig(int num, const char* const* keys, const char* const* vals) {
  return Guard([&] {
    std::string fn = "profile.json"; bool agg = false, cont = false; double period = 1.0;
    for (int i = 0; i < num; ++i) {
      const std::string k = keys[i], v = vals[i];
      if (k == "filename") fn = v;
      else if (k == "aggregate_stats") agg = (v == "1" || v == "true" || v == "True");
      else if (k == "continuous_dump") cont = (v == "1" || v == "true" || v == "True");
      else if (k == "dump_period") period = std::stod(v);
    }
    hips::Profiler::Get()->SetConfig(fn, agg, cont, period);
  });
}
GX_CAPI int GXSetProfilerState(int state) { return Guard([&] { hips::Profiler::Get()->SetState(state != 0); }); }
GX_CAPI int GXProfilePause(int paused) { return Guard([&] { hips::Profiler::Get()->Pause(paused != 0); }); }
GX_CAPI int GXDumpProfile(int finished) { return Guard([&] { hips::Profiler::Get()->Dump(finished != 0); }); }
// instant marker / duration event from a non-Python front end (MXProfileSetMarker, MXProfileDurationStart/Stop collapsed into one call)
GX_CAPI int GXProfileSetMarker(const char* name, const char* category) {
  return Guard([&] { hips::Profiler::Get()->Add(name, category ? category : "marker", 'i', hips::Profiler::NowUs()); });
}
GX_CAPI int GXProfileAddDuration(const char* name, const char* category, double start_us, double dur_us) {
  return Guard([&] { hips::Profiler::Get()->Add(name, category ? category : "operator", 'X', start_us, dur_us); });
}
GX_CAPI double GXProfileNowUs() { return hips::Profiler::NowUs(); }

// ------------------------------------------------------------------------------------------------ dependency engine
typedef void (*GXEngineFn)(void* arg);
GX_CAPI int GXEngineNewVariable(int* out) { return Guard([&] { *out = Eng()->NewVariable(); }); }
// fn(arg) runs once every earlier writer of the const vars and every earlier reader/writer of the mutable vars has completed
GX_CAPI int GXEnginePushAsync(GXEngineFn fn, void* arg, const int* const_vars, int num_const, const int* mutable_vars, int num_mutable, int priority,
                              const char* name) {
  return Guard([&] {
    Eng()->Push([fn, arg] { fn(arg); }, std::vector<int>(const_vars, const_vars + num_const), std::vector<int>(mutable_vars, mutable_vars + num_mutable),
                priority, name ? name : "c_api_op");
  });
}
// same, on the worker pool of `device` (-1: CPU) chosen by `prop` (0 normal / compute, 1 copy, 2 priority) — engine.h FnProperty + exec_ctx
GX_CAPI int GXEnginePushAsyncEx(GXEngineFn fn, void* arg, const int* const_vars, int num_const, const int* mutable_vars, int num_mutable, int priority,
                                const char* name, int device, int prop) {
  return Guard([&] {
    Eng()->Push([fn, arg] { fn(arg); }, std::vector<int>(const_vars, const_vars + num_const), std::vector<int>(mutable_vars, mutable_vars + num_mutable),
                priority, name ? name : "c_api_op", device, static_cast<gxrt::FnProperty>(prop < 0 || prop > 2 ? 0 : prop));
  });
}
GX_CAPI int GXEngineDeleteVariable(int var) { return Guard([&] { Eng()->DeleteVariable(var); }); }
GX_CAPI int GXEngineWaitForVar(int var) { return Guard([&] { Eng()->WaitForVar(var); }); }
GX_CAPI int GXEngineWaitAll() { return Guard([&] { Eng()->WaitForAll(); }); }

// ------------------------------------------------------------------------------------------------ storage
GX_CAPI int GXStorageAlloc(size_t nbytes, void** out) {
  return Guard([&] { *out = HostPool().Alloc(nbytes); if (*out == nullptr) throw std::runtime_error("out of host memory"); });
}
GX_CAPI int GXStorageFree(void* p) { return Guard([&] { HostPool().Free(p); }); }
