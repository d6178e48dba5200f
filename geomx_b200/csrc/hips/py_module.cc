// pybind11 bindings of the native runtime (module geomx_b200.lib._C).
// Plays the role of the reference's flat C API for the kvstore / profiler / IO (include/mxnet/c_api.h:1949-2327 MXKVStore*, MXInitPSEnv,
// src/c_api/c_api_profile.cc) — as a typed Python module instead of ctypes over a C ABI.
#include <algorithm>
#include <pybind11/functional.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "../runtime/engine.h"
#include "../runtime/gpu_topology.h"
#include "../runtime/storage.h"
#include "../runtime/text_io.h"
#include "../runtime/io.h"
#include "../runtime/recordio.h"
#include "../runtime/params_io.h"
#include "../runtime/profiler.h"
#include "dgt.h"
#include "gradient_compression.h"
#include "half.h"
#include "key_codec.h"
#include "kvstore_dist.h"
#include "server_optim.h"
#include "tsengine.h"

namespace py = pybind11;
using namespace hips;

static py::array_t<float> WrapFloat(float* p, size_t n) {
  return py::array_t<float>({static_cast<py::ssize_t>(n)}, {static_cast<py::ssize_t>(sizeof(float))}, p, py::capsule(p, [](void*) {}));
}

// GEOMX_SEGV_TRACE=1: print the native stack of a crashing thread (addresses resolve with `addr2line -e lib/_C*.so`)
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
static void SegvTrace(int sig) {
  void* frames[64];
  const int n = backtrace(frames, 64);
  const char msg[] = "[hips] fatal signal, native stack:\n";
  ssize_t r = write(2, msg, sizeof(msg) - 1); (void)r;
  backtrace_symbols_fd(frames, n, 2);
  signal(sig, SIG_DFL);
  raise(sig);
}

PYBIND11_MODULE(_C, m) {
  if (const char* t = getenv("GEOMX_SEGV_TRACE")) { if (t[0] == '1') { signal(SIGSEGV, SegvTrace); signal(SIGABRT, SegvTrace); } }
  m.doc() = "geomx_b200 native runtime: HiPS transport + servers, compression codecs, .params IO, profiler, engine, data IO";
  py::register_exception<hips::Error>(m, "HipsError");

  // ---------------------------------------------------------------------------------------------- environment / roles
  m.def("init_ps_env", [](const std::map<std::string, std::string>& kv) { for (auto& p : kv) Environment::Get()->Set(p.first, p.second); },
        "MXInitPSEnv: override environment variables in-process");
  // message-buffer pool (block_pool.h): {cached_bytes, hits, misses, limit_bytes}; set_limit(0) disables pooling and trims the cache
  m.def("buffer_pool_stats", [] { uint64_t st[4]; BlockPool::Get()->Stats(st);
                                  return std::map<std::string, uint64_t>{{"cached_bytes", st[0]}, {"hits", st[1]}, {"misses", st[2]}, {"limit_bytes", st[3]}}; });
  m.def("buffer_pool_set_limit", [](uint64_t bytes) { BlockPool::Get()->SetLimit(static_cast<size_t>(bytes)); });
  m.def("buffer_pool_trim", [] { BlockPool::Get()->Trim(); });
  m.def("buffer_pool_class_of", [](uint64_t n) { return static_cast<uint64_t>(BlockPool::ClassOf(static_cast<size_t>(n))); });
  // allocate-and-drop `rounds` SArrays of `nbytes` (what a stream of received frames does): returns the distinct block addresses seen
  m.def("buffer_pool_probe", [](uint64_t nbytes, int rounds) { std::vector<uint64_t> seen; for (int i = 0; i < rounds; ++i) { SArray<char> a; a.Allocate(nbytes); a[0] = 1; a[nbytes - 1] = 2;
                                                               const uint64_t p = reinterpret_cast<uint64_t>(a.data()); if (std::find(seen.begin(), seen.end(), p) == seen.end()) seen.push_back(p); } return seen; });
  // the native server-side optimizer on caller-owned arrays (unit tests: big tensors take the multi-threaded path of server_optim.h)
  m.def("native_optimizer_run", [](const std::string& spec, py::array_t<float, py::array::c_style | py::array::forcecast> w0,
                                   py::array_t<float, py::array::c_style | py::array::forcecast> grads) {
    OptSpec sp = OptSpec::Parse(spec);
    if (!sp.valid()) throw std::invalid_argument("unknown optimizer spec: " + spec);
    const size_t n = static_cast<size_t>(w0.size());
    if (grads.ndim() != 2 || static_cast<size_t>(grads.shape(1)) != n) throw std::invalid_argument("grads must be [steps, n]");
    NativeOptimizer opt(sp);
    NativeOptimizer::State st;
    py::array_t<float> out(w0.size());
    memcpy(out.mutable_data(), w0.data(), n * sizeof(float));
    { py::gil_scoped_release rel; for (py::ssize_t t = 0; t < grads.shape(0); ++t) opt.Update(&st, out.mutable_data(), grads.data() + t * n, n); }
    return out;
  });
  m.def("server_threads", [] { return ServerThreads(); });
  m.def("is_worker_node", [] { Postoffice::Get()->InitEnvironment(); return Postoffice::Get()->is_worker(); });
  m.def("is_server_node", [] { Postoffice::Get()->InitEnvironment(); return Postoffice::Get()->is_server(); });
  m.def("is_scheduler_node", [] { Postoffice::Get()->InitEnvironment(); return Postoffice::Get()->is_scheduler(); });
  m.def("is_global_server_node", [] { Postoffice::Get()->InitEnvironment(); return Postoffice::Get()->is_global_server(); });
  m.def("is_global_scheduler_node", [] { Postoffice::Get()->InitEnvironment(); return Postoffice::Get()->is_global_scheduler(); });
  m.def("is_master_worker_node", [] { Postoffice::Get()->InitEnvironment(); return Postoffice::Get()->is_master_worker(); });

  // ---------------------------------------------------------------------------------------------- protocol helpers (unit tests)
  m.def("server_rank_to_id", [](int r, int plane) { return ServerRankToID(r, static_cast<Plane>(plane)); });
  m.def("worker_rank_to_id", [](int r, int plane) { return WorkerRankToID(r, static_cast<Plane>(plane)); });
  m.def("id_to_rank", [](int id, int plane) { return IDtoRank(id, static_cast<Plane>(plane)); });
  m.def("get_command_type", [](int req, int dtype) { return GetCommandType(static_cast<RequestType>(req), dtype); });
  m.def("depair_command_type", [](int cmd) { auto t = DepairDataHandleType(cmd); return std::make_pair(static_cast<int>(t.requestType), t.dtype); });
  m.def("dgt_get_channel", &DGTGetChannel);
  m.def("dgt_effective_k", &DGTEffectiveK, "important block fraction of one push; adaptive mode = share of the contribution mass");
  m.def("dgt_encode4", [](py::array_t<float, py::array::c_style> a) {
    std::vector<char> out; float mn, mx;
    DGTEncode4(a.data(), a.size(), &out, &mn, &mx);
    return py::make_tuple(py::bytes(out.data(), out.size()), mn, mx);
  });
  m.def("dgt_decode4", [](py::bytes b, size_t n, float mn, float mx) {
    std::string s = b; py::array_t<float> out(n);
    DGTDecode4(s.data(), n, mn, mx, out.mutable_data());
    return out;
  });
  m.def("pack_unpack_meta", [](int head, const std::string& body, int priority, int key, bool with_control) {
    Meta a; a.head = head; a.body = body; a.priority = priority; a.key = key; a.app_id = 3; a.timestamp = 7; a.sender = 9; a.recver = 101;
    a.request = true; a.push = true; a.compr = {0.5f, 2.f};
    if (with_control) { a.control.cmd = Control::ADD_NODE; Node n; n.hostname = "10.0.0.1"; n.port = 1234; n.id = 100; n.rank_hint = 2; a.control.node.push_back(n); }
    std::vector<char> buf; PackMeta(a, &buf);
    Meta b; UnpackMeta(buf.data(), buf.size(), &b);
    return py::make_tuple(b.head, b.body, b.priority, b.key, b.app_id, b.timestamp, b.sender, b.recver, b.request, b.push, b.compr,
                          b.control.cmd, b.control.node.empty() ? std::string() : b.control.node[0].hostname,
                          b.control.node.empty() ? -1 : b.control.node[0].rank_hint, buf.size());
  });
  // fuzzing hooks (tests/test_fuzz_codecs.py): arbitrary bytes must either decode or be rejected with an exception — never crash, never
  // allocate from an unchecked length; a decoded meta re-encodes to bytes that decode to the same meta
  m.def("fuzz_unpack_meta", [](py::bytes b) {
    const std::string s = b;
    Meta a;
    try { UnpackMeta(s.data(), s.size(), &a); } catch (const std::exception&) { return py::make_tuple(false, py::bytes()); }
    std::vector<char> again; PackMeta(a, &again);
    Meta c; UnpackMeta(again.data(), again.size(), &c);
    std::vector<char> third; PackMeta(c, &third);
    if (again != third) throw std::runtime_error("meta codec is not idempotent");
    return py::make_tuple(true, py::bytes(again.data(), again.size()));
  });
  m.def("pack_meta_fields", [](int head, int app, int customer, int ts, int sender, int recver, bool request, bool push, bool simple, const std::string& body,
                               int priority, int key, std::vector<float> compr, int ctrl_cmd, std::vector<std::tuple<int, int, std::string, int>> nodes) {
    Meta a; a.head = head; a.app_id = app; a.customer_id = customer; a.timestamp = ts; a.sender = sender; a.recver = recver; a.request = request;
    a.push = push; a.simple_app = simple; a.body = body; a.priority = priority; a.key = key; a.compr = compr; a.control.cmd = ctrl_cmd;
    for (auto& t : nodes) { Node n; n.role = std::get<0>(t); n.id = std::get<1>(t); n.hostname = std::get<2>(t); n.port = std::get<3>(t); a.control.node.push_back(n); }
    std::vector<char> buf; PackMeta(a, &buf);
    Meta b; UnpackMeta(buf.data(), buf.size(), &b);
    std::vector<std::tuple<int, int, std::string, int>> back;
    for (auto& n : b.control.node) back.emplace_back(n.role, n.id, n.hostname, n.port);
    return py::make_tuple(py::bytes(buf.data(), buf.size()),
                          py::make_tuple(b.head, b.app_id, b.customer_id, b.timestamp, b.sender, b.recver, b.request, b.push, b.simple_app, b.body, b.priority,
                                         b.key, b.compr, b.control.cmd, back));
  });
  m.def("fuzz_dgt_block", [](int seq, int seq_end, int total_bytes, int val_bytes, int bits_num, int ncompr, int payload_len, int nparts, bool reset) {
    // one reassembly step of the DGT receiver on a block whose header fields are arbitrary: must be dropped or accepted, never index out of range
    // (reset: start from a receiver without partial tensors — the fuzz tests deliberately leave some behind)
    static std::unique_ptr<DGTReceiver> rxp;
    if (reset || !rxp) rxp.reset(new DGTReceiver());
    DGTReceiver& rx = *rxp;
    Message blk, whole;
    blk.meta.sender = 9; blk.meta.first_key = 1; blk.meta.timestamp = 3; blk.meta.seq = seq; blk.meta.seq_end = seq_end;
    blk.meta.total_bytes = total_bytes; blk.meta.val_bytes = val_bytes; blk.meta.bits_num = bits_num; blk.meta.msg_type = 1;
    blk.meta.compr.assign(static_cast<size_t>(std::max(0, std::min(ncompr, 4))), 0.5f);
    SArray<char> keys; keys.resize(8);
    SArray<char> vals; vals.resize(static_cast<size_t>(std::max(0, std::min(payload_len, 1 << 16))));
    if (nparts >= 1) blk.data.push_back(keys);
    if (nparts >= 2) blk.data.push_back(vals);
    return rx.Add(blk, &whole);
  }, py::arg("seq"), py::arg("seq_end"), py::arg("total_bytes"), py::arg("val_bytes"), py::arg("bits_num"), py::arg("ncompr"), py::arg("payload_len"), py::arg("nparts"),
     py::arg("reset") = false);
  m.def("ts_pick_receiver", [](int requester, std::vector<int> idle, std::map<int, long> known, float max_greed, int trials) {
    Environment::Get()->Set("MAX_GREED_RATE_TS", std::to_string(max_greed));
    TSScheduler s(nullptr, 8, kLocal);
    for (auto& kv : known) s.Record(requester, kv.first, kv.second);
    std::vector<int> picks;
    for (int i = 0; i < trials; ++i) picks.push_back(s.PickReceiver(requester, idle));
    return picks;
  }, py::arg("requester"), py::arg("idle"), py::arg("known"), py::arg("max_greed"), py::arg("trials") = 1);
  m.def("half_roundtrip", [](float f) { return py::make_tuple(HalfToFloat(FloatToHalf(f)), BF16ToFloat(FloatToBF16(f)), FloatToHalf(f)); });

  // ---------------------------------------------------------------------------------------------- CPU compression codecs
  py::class_<GradientCompression>(m, "GradientCompression")
      .def(py::init<>())
      .def("set_params", &GradientCompression::SetParams)
      .def("encode_params", &GradientCompression::EncodeParams)
      .def("decode_params", &GradientCompression::DecodeParams)
      .def("quantize_2bit", [](GradientCompression& g, py::array_t<float, py::array::c_style> grad, py::array_t<float, py::array::c_style> residual) {
        const int64_t n = grad.size();
        py::array_t<uint32_t> out(GradientCompression::CompressedSize2Bit(n));
        g.Quantize2Bit(grad.data(), residual.mutable_data(), out.mutable_data(), n);
        return out;
      })
      .def("dequantize_2bit", [](GradientCompression& g, py::array_t<uint32_t, py::array::c_style> in, int64_t n) {
        py::array_t<float> out(n);
        g.Dequantize2Bit(in.data(), out.mutable_data(), n);
        return out;
      })
      .def("bsc_compress", [](GradientCompression& g, py::array_t<float, py::array::c_style> grad, py::array_t<float, py::array::c_style> u,
                              py::array_t<float, py::array::c_style> v) {
        int k, s, ks; GradientCompression::BSCSizes(grad.size(), g.threshold(), &k, &s, &ks);
        py::array_t<float> out(2 * k);
        g.BSCompress(grad.data(), u.mutable_data(), v.mutable_data(), out.mutable_data(), grad.size());
        return out;
      })
      .def("bsc_pull_compress", [](GradientCompression& g, py::array_t<float, py::array::c_style> dense, int mult) {
        py::array_t<float> out(GradientCompression::BSCPullSize(dense.size(), g.threshold(), mult));
        g.BSCPullCompress(dense.data(), out.mutable_data(), dense.size(), mult);
        return out;
      })
      .def_static("bsc_decompress", [](py::array_t<float, py::array::c_style> z, int64_t n) {
        py::array_t<float> out(n);
        GradientCompression::BSCDecompress(z.data(), z.size(), out.mutable_data(), n);
        return out;
      });

  // ---------------------------------------------------------------------------------------------- TSEngine helpers (unit tests)
  m.def("ts_merge_bytes", [](py::array dst, py::array src, int dtype) {
    HIPS_CHECK_MSG(dst.nbytes() == src.nbytes(), "ts_merge_bytes: size mismatch");
    hips::TSMergeBytes(static_cast<char*>(dst.mutable_data()), static_cast<const char*>(src.data()), static_cast<size_t>(dst.nbytes()), dtype);
  }, "dst += src in the payload dtype (0 fp32, 2 fp16, 12 bf16) — the merge TSEngine nodes apply to peer contributions");
  m.def("ts_origins_roundtrip", [](std::vector<std::tuple<int, int, int>> v) {
    std::vector<hips::TSOrigin> o;
    for (auto& t : v) o.push_back(hips::TSOrigin{std::get<0>(t), std::get<1>(t), std::get<2>(t)});
    std::vector<std::tuple<int, int, int>> out;
    for (auto& x : hips::DecodeOrigins(hips::EncodeOrigins(o))) out.emplace_back(x.sender, x.timestamp, x.customer);
    return out;
  });
  m.def("ts_dtype_of_cmd", &hips::TSDTypeOfCmd);

  // ---------------------------------------------------------------------------------------------- storage pools / resources
  py::class_<gx_rt::PooledHostStorage>(m, "PooledHostStorage")
      .def(py::init<size_t, size_t>(), py::arg("page") = 4096, py::arg("max_pooled") = size_t(4) << 30)
      .def("round_size", &gx_rt::PooledHostStorage::RoundSize)
      .def("alloc", [](gx_rt::PooledHostStorage& s, size_t n) { bool hit = false; void* p = s.Alloc(n, &hit); return py::make_tuple(reinterpret_cast<uintptr_t>(p), hit); })
      .def("free", [](gx_rt::PooledHostStorage& s, uintptr_t p) { return s.Free(reinterpret_cast<void*>(p)); })
      .def("size_of", [](gx_rt::PooledHostStorage& s, uintptr_t p) { return s.SizeOf(reinterpret_cast<void*>(p)); })
      .def("release_all", [](gx_rt::PooledHostStorage& s) {
        std::vector<std::pair<uintptr_t, size_t>> out;
        for (auto& kv : s.ReleaseAll()) out.emplace_back(reinterpret_cast<uintptr_t>(kv.first), kv.second);
        return out;
      })
      .def("stats", [](gx_rt::PooledHostStorage& s) {
        auto st = s.stats();
        py::dict d; d["used_bytes"] = st.used_bytes; d["pooled_bytes"] = st.pooled_bytes; d["num_alloc"] = st.num_alloc;
        d["num_pool_hits"] = st.num_pool_hits; d["num_system_alloc"] = st.num_system_alloc;
        return d;
      });
  py::class_<gx_rt::ResourceManager>(m, "ResourceManager")
      .def(py::init<>())
      .def("temp_space", [](gx_rt::ResourceManager& r, int dev, int slot, size_t n, gx_rt::PooledHostStorage& pool) {
        return reinterpret_cast<uintptr_t>(r.TempSpace(dev, slot, n, &pool));
      })
      .def("seed", &gx_rt::ResourceManager::SeedAll)
      .def("next_seed", &gx_rt::ResourceManager::NextSeed);

  // ---------------------------------------------------------------------------------------------- text data readers
  m.def("parse_csv", [](const std::string& path) {
    gx_rt::CSVData d;
    { py::gil_scoped_release nogil; d = gx_rt::ParseCSV(path); }
    py::array_t<float> a({static_cast<py::ssize_t>(d.rows), static_cast<py::ssize_t>(d.cols)});
    if (!d.values.empty()) memcpy(a.mutable_data(), d.values.data(), d.values.size() * sizeof(float));
    return a;
  }, "dense float matrix of a CSV file");
  m.def("parse_libsvm", [](const std::string& path) {
    gx_rt::LibSVMData d;
    { py::gil_scoped_release nogil; d = gx_rt::ParseLibSVM(path); }
    return py::make_tuple(py::array_t<float>(d.labels.size(), d.labels.data()), py::array_t<float>(d.values.size(), d.values.data()),
                          py::array_t<long>(d.indices.size(), d.indices.data()), py::array_t<long>(d.indptr.size(), d.indptr.data()), d.max_index);
  }, "(labels, values, indices, indptr, max_index) of a LibSVM file");

  // ---------------------------------------------------------------------------------------------- GPU topology solver
  m.def("topology_tree", [](std::vector<float> W, int n, int root) {
    HIPS_CHECK_MSG(static_cast<int>(W.size()) == n * n && root >= 0 && root < n, "topology_tree: W must be n*n and 0 <= root < n");
    gx_rt::TopologySolver solver(W, n);
    gx_rt::TopoTree t = solver.BuildTree(root);
    return py::make_tuple(t.parent, t.round, t.depth);
  }, "binary reduction tree over a link-weight matrix: (parent[], round[], depth)");
  m.def("topology_bisect", [](std::vector<float> W, int n, std::vector<int> set, int pin) {
    gx_rt::TopologySolver solver(W, n);
    std::vector<int> A, B;
    solver.Bisect(set, pin, &A, &B);
    return py::make_tuple(A, B);
  }, "Kernighan-Lin balanced bisection of a device set (the half containing `pin` first)");

  // ---------------------------------------------------------------------------------------------- KVStoreDist
  py::class_<KVStoreDist>(m, "KVStoreDist")
      .def(py::init<const std::string&>(), py::call_guard<py::gil_scoped_release>())
      .def("shutdown", &KVStoreDist::Shutdown, py::call_guard<py::gil_scoped_release>())
      .def_property_readonly("rank", &KVStoreDist::rank)
      .def_property_readonly("num_workers", &KVStoreDist::num_workers)
      .def_property_readonly("num_all_workers", &KVStoreDist::num_all_workers)
      .def_property_readonly("is_master_worker", &KVStoreDist::is_master_worker)
      .def_property_readonly("is_recovery", &KVStoreDist::is_recovery)
      .def("num_dead_node", &KVStoreDist::num_dead_node)
      .def("init", [](KVStoreDist& kv, int key, uintptr_t ptr, size_t elems, int dtype) { kv.Init(key, reinterpret_cast<const void*>(ptr), elems, dtype); },
           py::call_guard<py::gil_scoped_release>())
      .def("push", [](KVStoreDist& kv, int key, uintptr_t ptr, size_t elems, int dtype, int priority) {
        return kv.Push(key, reinterpret_cast<const void*>(ptr), elems, dtype, priority);
      }, py::call_guard<py::gil_scoped_release>())
      .def("pull", [](KVStoreDist& kv, int key, uintptr_t ptr, size_t elems, int dtype, int priority) {
        return kv.Pull(key, reinterpret_cast<void*>(ptr), elems, dtype, priority);
      }, py::call_guard<py::gil_scoped_release>())
      .def("wait", &KVStoreDist::Wait, py::call_guard<py::gil_scoped_release>())
      .def("wait_all", &KVStoreDist::WaitAll, py::call_guard<py::gil_scoped_release>())
      .def("set_gradient_compression", &KVStoreDist::SetGradientCompression, py::call_guard<py::gil_scoped_release>())
      .def("barrier", &KVStoreDist::Barrier, py::call_guard<py::gil_scoped_release>())
      .def("send_command_to_servers", &KVStoreDist::SendCommandToServers, py::call_guard<py::gil_scoped_release>())
      .def("send_bytes", &KVStoreDist::send_bytes)
      .def("recv_bytes", &KVStoreDist::recv_bytes)
      .def("ts_stats", &KVStoreDist::ts_stats)
      .def("push_rows", [](KVStoreDist& kv, int key, uintptr_t ids, size_t nrows, uintptr_t rows, size_t row_len, int priority) {
        return kv.PushRows(key, reinterpret_cast<const int64_t*>(ids), nrows, reinterpret_cast<const float*>(rows), row_len, priority);
      }, py::call_guard<py::gil_scoped_release>())
      .def("pull_rows", [](KVStoreDist& kv, int key, uintptr_t ids, size_t nrows, uintptr_t out, size_t row_len, int priority) {
        return kv.PullRows(key, reinterpret_cast<const int64_t*>(ids), nrows, reinterpret_cast<float*>(out), row_len, priority);
      }, py::call_guard<py::gil_scoped_release>())
      .def("run_server", [](KVStoreDist& kv, py::object controller, py::object updater) {
        // python objects are held through shared_ptrs whose deleter re-acquires the GIL: the std::functions are copied / destroyed
        // by server threads that do not hold it
        auto hold = [](py::object o) {
          return std::shared_ptr<py::object>(new py::object(std::move(o)), [](py::object* p) { py::gil_scoped_acquire g; delete p; });
        };
        KVStoreDistServer::Controller c = nullptr;
        KVStoreDistServer::Updater u = nullptr;
        if (!controller.is_none()) {
          auto ch = hold(controller);
          c = [ch](int head, const std::string& body) { py::gil_scoped_acquire g; (*ch)(head, py::bytes(body)); };
        }
        if (!updater.is_none()) {
          auto uh = hold(updater);
          u = [uh](int key, const float* grad, float* weight, size_t n) {
            py::gil_scoped_acquire g;
            (*uh)(key, WrapFloat(const_cast<float*>(grad), n), WrapFloat(weight, n));
          };
        }
        controller = py::none(); updater = py::none();
        py::gil_scoped_release rel;
        kv.RunServer(c, u);
        c = nullptr; u = nullptr;
      });

  // ---------------------------------------------------------------------------------------------- profiler
  m.def("profiler_set_config", [](const std::string& fn, bool agg, bool cont, double period) { Profiler::Get()->SetConfig(fn, agg, cont, period); });
  m.def("profiler_set_state", [](bool run) { Profiler::Get()->SetState(run); });
  m.def("profiler_pause", [](bool p) { Profiler::Get()->Pause(p); });
  m.def("profiler_add_event", [](const std::string& name, const std::string& cat, const std::string& ph, double ts, double dur, int pid, int tid, double value) {
    Profiler::Get()->Add(name, cat, ph.empty() ? 'X' : ph[0], ts, dur, pid, tid, value);
  });
  m.def("profiler_now_us", [] { return Profiler::NowUs(); });
  m.def("profiler_dump", [](bool finished) { Profiler::Get()->Dump(finished); });
  m.def("profiler_aggregate", [] { return Profiler::Get()->AggregateTable(); });
  m.def("profiler_clear", [] { Profiler::Get()->Clear(); });
  m.def("profiler_size", [] { return Profiler::Get()->size(); });

  // ---------------------------------------------------------------------------------------------- .params IO, data IO, engine
  gxrt::BindParamsIO(m);
  gxrt::BindIO(m);
  gxrt::BindRecordIO(m);
  gxrt::BindEngine(m);
}
