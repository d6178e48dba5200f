#include "kvstore_dist_server.h"

#include <chrono>
#include <fstream>

#include "key_codec.h"

namespace hips {

KVStoreDistServer::KVStoreDistServer() {
  Postoffice* po = Postoffice::Get();
  Environment* env = Environment::Get();
  is_global_ = po->is_global_server();
  has_global_ = po->has_plane(kGlobal) && !is_global_;
  standalone_ = !po->has_plane(kGlobal);
  use_hfa_ = env->GetInt("MXNET_KVSTORE_USE_HFA", 0) != 0;
  hfa_k2_ = std::max(1, env->GetInt("MXNET_KVSTORE_HFA_K2", 1));
  bigarray_bound_ = static_cast<size_t>(env->GetFloat("MXNET_KVSTORE_BIGARRAY_BOUND", 1000000));
  size_lower_bound_ = static_cast<size_t>(env->GetFloat("MXNET_KVSTORE_SIZE_LOWER_BOUND", 200000));
  ckpt_prefix_ = env->GetStr("GEOMX_SERVER_CKPT_PREFIX", "");
  ckpt_every_ = ckpt_prefix_.empty() ? 0 : std::max(0, env->GetInt("GEOMX_SERVER_CKPT_EVERY", 0));
  fused_tier_pull_ = env->GetInt("GEOMX_FUSED_TIER_PULL", 1) != 0 && env->GetInt("ENABLE_INTER_TS", 0) == 0;
  resume_wanted_ = !ckpt_prefix_.empty() && env->GetInt("GEOMX_SERVER_RESUME", 0) != 0 && (is_global_ || standalone_);
  const int lanes = std::max(0, env->GetInt("GEOMX_SERVER_LANES", 4));   // 0: pushes are handled on the customer thread itself
  for (int i = 0; i < lanes; ++i) lanes_.emplace_back(new Lane());
  ps_server_.reset(new KVServer(0));
  ps_server_->SimpleApp::set_request_handle([this](const SimpleData& d, SimpleApp* app) { CommandHandle(d, app); });
  ps_server_->set_request_handle([this](const KVMeta& m, const KVPairs& d, KVServer* s) { DataHandleEx(m, d, s); });
  ps_server_->set_response_handle([this](const KVMeta& m, const KVPairs& d, KVServer* s) { ResponseHandle(m, d, s); });
  // With lanes the customer thread would only forward data requests to them: let the transport's receive thread do that itself (one thread
  // wake-up less per request).  Kept off where the customer queue does real work: P3 (priority order of queued requests) and TSEngine
  // (its node logic sees every message first).
  if (!lanes_.empty() && !ps_server_->enable_p3 && ps_server_->ts(kLocal) == nullptr && ps_server_->ts(kGlobal) == nullptr &&
      env->GetInt("GEOMX_INLINE_REQUESTS", 1) != 0)
    ps_server_->get_customer()->set_inline_requests(true);
  if (has_global_ && ps_server_->ts(kGlobal) != nullptr)   // local server: fresh values of TS rounds arrive through the relay
    ps_server_->ts(kGlobal)->set_on_relayed([this](int key, int version, int cmd, const std::vector<char>& bytes) { OnRelayedFromGlobal(key, version, cmd, bytes); });
}

KVStoreDistServer::~KVStoreDistServer() {
  lanes_.clear();          // drain and join the push lanes before the transport goes away
  ps_server_.reset();
}

int KVStoreDistServer::rank_local() { return Postoffice::Get()->my_rank(kLocal); }

// ------------------------------------------------------------------------------------------------ key registry
KVStoreDistServer::KeyState* KVStoreDistServer::Find(int key) {
  std::shared_lock<std::shared_mutex> lk(reg_mu_);
  auto it = keys_.find(key);
  return it == keys_.end() ? nullptr : it->second.get();
}

KVStoreDistServer::KeyState& KVStoreDistServer::Slot(int key) {
  if (KeyState* ks = Find(key)) return *ks;
  std::unique_lock<std::shared_mutex> lk(reg_mu_);
  std::unique_ptr<KeyState>& p = keys_[key];
  if (!p) p.reset(new KeyState());
  return *p;
}

std::vector<std::pair<int, KVStoreDistServer::KeyState*>> KVStoreDistServer::Slots() {
  std::shared_lock<std::shared_mutex> lk(reg_mu_);
  std::vector<std::pair<int, KeyState*>> out;
  out.reserve(keys_.size());
  for (auto& kv : keys_) out.emplace_back(kv.first, kv.second.get());
  return out;
}

std::vector<float> KVStoreDistServer::GetStored(int key) {
  KeyState* ks = Find(key);
  if (ks == nullptr) return {};
  std::lock_guard<std::mutex> lk(ks->mu);
  const Entry& e = ks->entry;
  std::vector<float> out(e.elems);
  if (e.has_master) out = e.master;
  else ToFloat(e.data.data(), e.dtype, e.elems, out.data());
  return out;
}

void KVStoreDistServer::Send(std::vector<Reply>* out) {
  for (Reply& r : *out) {
    if (r.data.keys.size()) ps_server_->Response(r.to, r.data);
    else ps_server_->Response(r.to);
  }
  out->clear();
}

KVStoreDistServer::Reply KVStoreDistServer::StoredReply(const KVMeta& to, int key, const Entry& e) const {
  Reply r; r.to = to;
  r.data.keys.push_back(static_cast<Key>(key));
  r.data.vals.CopyFrom(e.data.data(), e.data.size());
  r.data.lens.push_back(static_cast<int>(e.data.size()));
  return r;
}

// ------------------------------------------------------------------------------------------------ dtype helpers
void KVStoreDistServer::ToFloat(const char* src, int dtype, size_t n, float* dst) {
  switch (dtype) {
    case kFloat32: memcpy(dst, src, n * 4); break;
    case kFloat64: { const double* p = reinterpret_cast<const double*>(src); for (size_t i = 0; i < n; ++i) dst[i] = static_cast<float>(p[i]); break; }
    case kFloat16: { const uint16_t* p = reinterpret_cast<const uint16_t*>(src); for (size_t i = 0; i < n; ++i) dst[i] = HalfToFloat(p[i]); break; }
    case kBfloat16: { const uint16_t* p = reinterpret_cast<const uint16_t*>(src); for (size_t i = 0; i < n; ++i) dst[i] = BF16ToFloat(p[i]); break; }
    case kInt32: { const int32_t* p = reinterpret_cast<const int32_t*>(src); for (size_t i = 0; i < n; ++i) dst[i] = static_cast<float>(p[i]); break; }
    case kInt64: { const int64_t* p = reinterpret_cast<const int64_t*>(src); for (size_t i = 0; i < n; ++i) dst[i] = static_cast<float>(p[i]); break; }
    case kInt8: { const int8_t* p = reinterpret_cast<const int8_t*>(src); for (size_t i = 0; i < n; ++i) dst[i] = p[i]; break; }
    default: { const uint8_t* p = reinterpret_cast<const uint8_t*>(src); for (size_t i = 0; i < n; ++i) dst[i] = p[i]; break; }
  }
}

void KVStoreDistServer::StoreFromFloat(Entry* e, const float* src, size_t n) {
  e->data.resize(n * DTypeSize(e->dtype));
  char* d = e->data.data();
  switch (e->dtype) {
    case kFloat32: memcpy(d, src, n * 4); break;
    case kFloat64: { double* p = reinterpret_cast<double*>(d); for (size_t i = 0; i < n; ++i) p[i] = src[i]; break; }
    case kFloat16: { uint16_t* p = reinterpret_cast<uint16_t*>(d); for (size_t i = 0; i < n; ++i) p[i] = FloatToHalf(src[i]); break; }
    case kBfloat16: { uint16_t* p = reinterpret_cast<uint16_t*>(d); for (size_t i = 0; i < n; ++i) p[i] = FloatToBF16(src[i]); break; }
    case kInt32: { int32_t* p = reinterpret_cast<int32_t*>(d); for (size_t i = 0; i < n; ++i) p[i] = static_cast<int32_t>(src[i]); break; }
    case kInt64: { int64_t* p = reinterpret_cast<int64_t*>(d); for (size_t i = 0; i < n; ++i) p[i] = static_cast<int64_t>(src[i]); break; }
    case kInt8: { int8_t* p = reinterpret_cast<int8_t*>(d); for (size_t i = 0; i < n; ++i) p[i] = static_cast<int8_t>(src[i]); break; }
    default: { uint8_t* p = reinterpret_cast<uint8_t*>(d); for (size_t i = 0; i < n; ++i) p[i] = static_cast<uint8_t>(src[i]); break; }
  }
}

// ------------------------------------------------------------------------------------------------ commands
void KVStoreDistServer::CommandHandle(const SimpleData& recved, SimpleApp* app) {
  const CommandType cmd = static_cast<CommandType>(recved.head);
  Postoffice* po = Postoffice::Get();
  switch (cmd) {
    case CommandType::kStopServer: {
      app->Response(recved);
      if (is_global_) {
        // a global server stops after every local server (global worker) voted; central workers' votes on the local plane are ignored
        bool stop = false;
        { std::lock_guard<std::mutex> lk(ctl_mu_); if (recved.plane == kGlobal) stop = (++stop_votes_ == po->num_global_workers()); }
        if (stop) exec_.Stop();
      } else {
        bool first;
        { std::lock_guard<std::mutex> lk(ctl_mu_); first = !stop_requested_; stop_requested_ = true; }
        if (first && has_global_) {  // relay to the global servers, do not wait (they stop only after all parties voted)
          ps_server_->Request(static_cast<int>(CommandType::kStopServer), "", kServerGroup, kGlobal);
        }
        exec_.Stop();          // idempotent: a second stop command is acknowledged, not queued behind a loop that has already ended
      }
      return;
    }
    case CommandType::kSyncMode: sync_mode_ = true; break;
    case CommandType::kSyncGlobalMode: sync_global_mode_ = true; break;
    case CommandType::kSetMultiPrecision: {
      if (!multi_precision_.exchange(true)) {
        for (auto& kv : Slots()) {
          std::lock_guard<std::mutex> lk(kv.second->mu);
          Entry& e = kv.second->entry;
          if (!e.has_master && e.dtype != kFloat32) { e.master.resize(e.elems); ToFloat(e.data.data(), e.dtype, e.elems, e.master.data()); e.has_master = true; }
        }
      }
      break;
    }
    case CommandType::kSetGradientCompression: {
      gc_.DecodeParams(recved.body);
      // the global server of rank 0 relays the setting to every local server (reference :337-344); workers of other parties never learn it
      if (is_global_ && recved.plane == kLocal && po->my_rank(kGlobal) == 0) {
        // fire-and-forget: the acks arrive on this very thread's queue, so waiting here would deadlock; scripts sleep ~1 s after
        // configuration ("waiting for configurations to complete", examples/cnn_bsc.py)
        ps_server_->Request(static_cast<int>(CommandType::kSetGradientCompression), recved.body, kWorkerGroup, kGlobal);
      }
      break;
    }
    case CommandType::kSetOptimizerSpec: {
      OptSpec s = OptSpec::Parse(recved.body);
      HIPS_CHECK_MSG(s.valid(), "unknown native optimizer spec: " + recved.body);
      // the per-key optimizer state (moments, step count — possibly restored from a checkpoint) stays with the keys
      std::atomic_store(&native_opt_, std::shared_ptr<const NativeOptimizer>(new NativeOptimizer(s)));
      break;
    }
    case CommandType::kSetProfilerParams: {
      // body = "k:v,k:v,...<digit>": 0 set_config, 1 state, 2 pause, 3 dump  (reference :409-456; filename gets a rank<r>_ prefix)
      HIPS_CHECK(!recved.body.empty());
      const int which = recved.body.back() - '0';
      const std::string params = recved.body.substr(0, recved.body.size() - 1);
      Profiler* pf = Profiler::Get();
      if (which == 0) {
        std::string fn = "profile.json"; bool agg = false, cont = false; double period = 1.0;
        size_t p = 0;
        while (p < params.size()) {
          size_t e = params.find(',', p); if (e == std::string::npos) e = params.size();
          const std::string kv = params.substr(p, e - p);
          const size_t c = kv.find(':');
          if (c != std::string::npos) {
            const std::string k = kv.substr(0, c), v = kv.substr(c + 1);
            if (k == "filename") fn = v; else if (k == "aggregate_stats") agg = (v == "True" || v == "1");
            else if (k == "continuous_dump") cont = (v == "True" || v == "1"); else if (k == "dump_period") period = atof(v.c_str());
          }
          p = e + 1;
        }
        const size_t slash = fn.find_last_of('/');
        const std::string prefix = "rank" + std::to_string(po->my_rank(is_global_ ? kGlobal : kLocal)) + "_";
        fn = slash == std::string::npos ? prefix + fn : fn.substr(0, slash + 1) + prefix + fn.substr(slash + 1);
        pf->SetConfig(fn, agg, cont, period);
      } else if (which == 1) pf->SetState(params == "1");
      else if (which == 2) pf->Pause(params == "1");
      else if (which == 3) pf->Dump(params == "1");
      break;
    }
    case CommandType::kSaveStates: SaveStates(recved.body); break;
    case CommandType::kLoadStates: LoadStates(recved.body); break;
    case CommandType::kController:
    default: {
      // foreign controller (pickled optimizer ...) runs on the main thread, like the reference's exec_.Exec(controller_)
      if (controller_) exec_.Exec([this, recved]() { controller_(recved.head, recved.body); });
      break;
    }
  }
  app->Response(recved);
}

// ------------------------------------------------------------------------------------------------ data
void KVStoreDistServer::DataHandleEx(const KVMeta& req, const KVPairs& data, KVServer* server) {
  const DataHandleType type = DepairDataHandleType(req.cmd);
  // the server object exists before the node has registered (its rank names the checkpoint file), so the resume happens on first traffic
  if (resume_wanted_) std::call_once(resume_once_, [this] { TryResume(); });
  if (!req.push) {
    // pulls never block (parked until the key has a value), so without lanes they are answered right here; with lanes they take the
    // key's lane, which keeps them ordered behind the pushes of that key that arrived earlier
    if (lanes_.empty()) {
      ProfileScope ps("KVStoreDistServerPull");
      HandlePull(type, req, data);
    } else {
      Lane* lane = lanes_[static_cast<size_t>(req.key < 0 ? -req.key : req.key) % lanes_.size()].get();
      lane->Post([this, type, req, data] {
        ProfileScope ps("KVStoreDistServerPull");
        HandlePull(type, req, data);
      });
    }
    return;
  }
  if (lanes_.empty()) {
    ProfileScope ps("KVStoreDistServerPush");
    HandlePush(type, req, data);
    return;
  }
  // the payload is reference counted: the lane keeps the message buffer alive
  Lane* lane = lanes_[static_cast<size_t>(req.key < 0 ? -req.key : req.key) % lanes_.size()].get();
  lane->Post([this, type, req, data] {
    ProfileScope ps("KVStoreDistServerPush");
    HandlePush(type, req, data);
  });
}

void KVStoreDistServer::ApplyUpdate(int key, KeyState* ks, const float* grad, size_t n) {
  Entry* e = &ks->entry;
  float* w = e->has_master ? e->master.data() : reinterpret_cast<float*>(e->data.data());
  const std::shared_ptr<const NativeOptimizer> opt = std::atomic_load(&native_opt_);
  if (opt) opt->Update(&ks->opt, w, grad, n);
  else if (updater_) exec_.Exec([this, key, grad, w, n]() { updater_(key, grad, w, n); });   // foreign updaters run one at a time on the main thread
  else memcpy(w, grad, n * sizeof(float));  // no optimizer on the server: store the aggregate (reference ApplyUpdates :547-550)
  if (e->has_master) StoreFromFloat(e, e->master.data(), n);
}

void KVStoreDistServer::HandlePush(const DataHandleType& type, const KVMeta& req, const KVPairs& data) {
  HIPS_CHECK(data.keys.size() == 1);
  const int key = req.key;
  Postoffice* po = Postoffice::Get();
  num_pushes_++;
  KeyState& ks = Slot(key);
  std::vector<Reply> out;                 // built under the key's lock, sent after it is released
  std::unique_lock<std::mutex> lk(ks.mu);
  Entry& e = ks.entry;
  const bool p3 = ps_server_->enable_p3;
  // inter-tier fusion (GEOMX_FUSED_TIER_PULL, default on): a global server answers a local server's dense push with the post-update value,
  // so the local server does not need a second round trip (push ack, then pull) over the slow link between parties
  const bool fuse_up = is_global_ && fused_tier_pull_ && type.requestType == RequestType::kDefaultPushPull;
  const bool fuse_bsc = is_global_ && fused_tier_pull_ && type.requestType == RequestType::kBSCompressedPushPull && sync_global_mode_;
  auto respond = [&](const KVMeta& r) {
    if (fuse_up && r.plane == kGlobal && r.sender % 2 == 1) {
      Reply rep = StoredReply(r, key, e);
      rep.data.keys = data.keys;
      out.push_back(std::move(rep));
    } else if (fuse_bsc && r.plane == kGlobal && r.sender % 2 == 1) {
      // Bi-Sparse: the answer is the re-sparsified aggregate a pull would have returned (capacity k * parties, reference :1190-1206)
      const int mult = std::max(1, Postoffice::Get()->num_global_workers());
      const float* w = e.has_master ? e.master.data() : reinterpret_cast<const float*>(e.data.data());
      std::vector<float> z(GradientCompression::BSCPullSize(static_cast<int64_t>(e.elems), gc_.threshold(), mult));
      gc_.BSCPullCompress(w, z.data(), static_cast<int64_t>(e.elems), mult);
      Reply rep; rep.to = r; rep.data.keys = data.keys;
      rep.data.vals.CopyFrom(reinterpret_cast<const char*>(z.data()), z.size() * sizeof(float));
      rep.data.lens.push_back(static_cast<int>(rep.data.vals.size()));
      out.push_back(std::move(rep));
    } else if (p3 && !is_global_) {
      Reply rep = StoredReply(r, key, e);
      rep.data.keys = data.keys;
      out.push_back(std::move(rep));
    } else {
      Reply rep; rep.to = r;
      out.push_back(std::move(rep));
    }
  };
  if (e.elems != 0 && any_skip_init_.load() && ks.skip_init_push && (standalone_ || req.plane == kLocal)) {
    // ---- resumed server: the restarted job initialises its keys again (kv.init); the checkpointed value wins, the push is only acknowledged.
    // On a global server the init comes from the master worker over the LOCAL plane of the central party — the parties may already be
    // training by then (their pulls found the restored keys initialised), so their pushes on the global plane must not be mistaken for it.
    ks.skip_init_push = false;
    respond(req);
    lk.unlock();
    Send(&out);
    AskTS(key);
    return;
  }
  if (e.elems == 0) {
    // ---- initialisation: the first push of a key defines it (reference :1237-1269)
    HIPS_CHECK(type.requestType == RequestType::kDefaultPushPull);
    const int bytes = DTypeSize(type.dtype);
    e.dtype = type.dtype;
    e.elems = data.vals.size() / bytes;
    e.data.assign(data.vals.data(), data.vals.data() + data.vals.size());
    if (e.dtype != kFloat32) {  // non-fp32 keys always carry an fp32 working copy; multi_precision keeps it as the master
      e.master.resize(e.elems); ToFloat(e.data.data(), e.dtype, e.elems, e.master.data()); e.has_master = true;
    }
    respond(req);
    const bool fetch = !(is_global_ || standalone_) && has_global_;
    if (is_global_ || standalone_) { ks.initialized = true; FlushParkedPulls(&ks, &out); }
    lk.unlock();
    Send(&out);
    if (fetch) PullFromGlobal(key, type);     // a local server adopts the global tier's value of the key
    AskTS(key);                               // TSEngine: open the first round of this key
    return;
  }
  // ---- decode the contribution to fp32
  const size_t n = e.elems;
  // dense fp32 payloads (the common case) are consumed in place from the message buffer; every other format is decoded into `incoming`
  std::vector<float> incoming;
  const float* inc = nullptr;
  const bool in_place = type.requestType == RequestType::kDefaultPushPull && type.dtype == kFloat32;
  if (!in_place) { incoming.resize(n); inc = incoming.data(); }
  if (type.requestType == RequestType::kDefaultPushPull) {
    HIPS_CHECK_MSG(data.vals.size() == n * DTypeSize(type.dtype), "push size mismatch for key " + std::to_string(key));
    if (in_place) inc = reinterpret_cast<const float*>(data.vals.data());
    else ToFloat(data.vals.data(), type.dtype, n, incoming.data());
  } else if (type.requestType == RequestType::kRowSparsePushPull) {
    // row_sparse gradient (reference DataHandleRowSparse :561-756): scatter-add the listed rows into a dense contribution — the
    // aggregation / tier logic below is storage-agnostic
    HIPS_CHECK_MSG(e.dtype == kFloat32 && data.vals.size() >= 2 * sizeof(int64_t), "row_sparse push needs an fp32 key");
    const int64_t* hdr = reinterpret_cast<const int64_t*>(data.vals.data());
    const size_t nrows = static_cast<size_t>(hdr[0]), row_len = static_cast<size_t>(hdr[1]);
    HIPS_CHECK_MSG(data.vals.size() == (2 + nrows) * sizeof(int64_t) + nrows * row_len * sizeof(float), "row_sparse push size mismatch");
    const int64_t* ids = hdr + 2;
    const float* rows = reinterpret_cast<const float*>(ids + nrows);
    std::fill(incoming.begin(), incoming.end(), 0.f);
    for (size_t r = 0; r < nrows; ++r) {
      const size_t base = static_cast<size_t>(ids[r]) * row_len;
      HIPS_CHECK_MSG(ids[r] >= 0 && base + row_len <= n, "row id out of range in row_sparse push");
      for (size_t j = 0; j < row_len; ++j) incoming[base + j] += rows[r * row_len + j];
    }
  } else if (type.requestType == RequestType::kCompressedPushPull) {
    HIPS_CHECK_MSG(data.vals.size() == static_cast<size_t>(GradientCompression::CompressedSize2Bit(static_cast<int64_t>(n))) * sizeof(uint32_t),
                   "2-bit push of key " + std::to_string(key) + " does not match the stored tensor (compressed keys are not partitioned: "
                   "set_gradient_compression before kv.init on every worker)");
    gc_.Dequantize2Bit(reinterpret_cast<const uint32_t*>(data.vals.data()), incoming.data(), static_cast<int64_t>(n));
  } else {
    HIPS_CHECK_MSG(data.vals.size() % (2 * sizeof(float)) == 0 && data.vals.size() / (2 * sizeof(float)) <= n,
                   "Bi-Sparse push of key " + std::to_string(key) + " is larger than the stored tensor");
    GradientCompression::BSCDecompress(reinterpret_cast<const float*>(data.vals.data()), data.vals.size() / sizeof(float), incoming.data(), n);
  }
  // central-party workers only train when DMLC_ENABLE_CENTRAL_WORKER=1 (reference :1274-1275)
  if (is_global_ && req.plane == kLocal && !po->enable_central_workers()) { respond(req); lk.unlock(); Send(&out); return; }

  const bool sync = (is_global_ ? sync_global_mode_.load() : sync_mode_.load());
  if (!sync) {
    // ---- MixedSync / async: apply this contribution immediately (reference :1582-1609)
    ApplyUpdate(key, &ks, inc, n);
    const std::vector<KVMeta> who = ExpandOrigins(req);
    for (size_t i = 0; i < who.size(); ++i) {
      if (i > 0 && who[i].sender == who[i - 1].sender && who[i].timestamp == who[i - 1].timestamp) continue;
      respond(who[i]);
    }
    lk.unlock();
    Send(&out);
    AskTS(key);
    return;
  }
  UpdateBuf& ub = ks.ub;
  if (ub.request.empty()) ub.merged.assign(inc, inc + n);
  else { float* m = ub.merged.data(); ParallelFor(n, size_t(1) << 18, [m, inc](size_t lo, size_t hi) { for (size_t i = lo; i < hi; ++i) m[i] += inc[i]; }); }
  for (const KVMeta& r : ExpandOrigins(req)) ub.request.push_back(r);
  size_t expected;
  if (is_global_) expected = po->num_global_workers() + (po->enable_central_workers() ? po->num_workers() : 0);
  else expected = po->num_workers();
  if (ub.request.size() < expected) {
    lk.unlock();
    AskTS(key);      // TSEngine: a partial delivery — this server is free to receive the next (merged) contribution
    return;
  }
  if (is_global_ || standalone_) {
    ApplyUpdate(key, &ks, ub.merged.data(), n);
    // the round counter moves (and a due periodic checkpoint is written) BEFORE any push of this round is acknowledged: until the acks
    // are out no worker can start round N+1 of any key, so a snapshot taken now holds exactly the round-N state of every key
    const bool checkpoint_due = BumpRound(&ks);
    std::vector<KVMeta> reqs; reqs.swap(ub.request);
    for (size_t i = 0; i < reqs.size(); ++i) {
      if (i > 0 && reqs[i].sender == reqs[i - 1].sender && reqs[i].timestamp == reqs[i - 1].timestamp) continue;  // merged duplicates (TS)
      respond(reqs[i]);
    }
    lk.unlock();
    if (checkpoint_due) SaveStates(ckpt_prefix_);
    Send(&out);
    RoundCompleted(key, /*bumped=*/true);
    return;
  }
  const bool local_round_done = FinishLocalAggregation(key, &ks, type, &out);
  lk.unlock();
  Send(&out);
  if (local_round_done) RoundCompleted(key);
}

// local server: all workers of the party have pushed `key`
bool KVStoreDistServer::FinishLocalAggregation(int key, KeyState* ks, const DataHandleType& type, std::vector<Reply>* out) {
  Entry& e = ks->entry;
  UpdateBuf* ub = &ks->ub;
  const size_t n = e.elems;
  float* w = e.has_master ? e.master.data() : reinterpret_cast<float*>(e.data.data());
  memcpy(w, ub->merged.data(), n * sizeof(float));   // only aggregate (ApplyUpdates on a non-global server)
  if (e.has_master) StoreFromFloat(&e, w, n);
  // HFA: every K2-th local round of a key goes to the global tier.  The reference counts local rounds on key 0 (:1324) and relies on one
  // thread serving the keys in arrival order; keys are served by concurrent lanes here, so every key counts its OWN rounds — all keys are
  // pushed once per iteration, hence the counters agree and the decision cannot depend on which lane ran first.
  ++ks->local_rounds;
  if (key == 0) local_iters_ = ks->local_rounds;
  if (use_hfa_ && (ks->local_rounds % hfa_k2_ != 0)) {   // local synchronisation only
    for (size_t i = 0; i < ub->request.size(); ++i) {
      const std::vector<KVMeta>& q = ub->request;
      if (i > 0 && q[i].sender == q[i - 1].sender && q[i].timestamp == q[i - 1].timestamp) continue;
      if (ps_server_->enable_p3) out->push_back(StoredReply(q[i], key, e));
      else { Reply rep; rep.to = q[i]; out->push_back(std::move(rep)); }
    }
    ub->request.clear();
    return true;
  }
  if (use_hfa_) {                                    // push the party's progress since the last global sync
    auto& ms = ks->milestone;
    HIPS_CHECK_MSG(ms.size() == n, "HFA milestone not initialised for key " + std::to_string(key));
    const float inv = 1.f / Postoffice::Get()->num_global_workers();
    for (size_t i = 0; i < n; ++i) w[i] = (w[i] - ms[i]) * inv;
    if (e.has_master) StoreFromFloat(&e, w, n);
  }
  ks->round.waiting.swap(ub->request);
  PushToGlobal(key, ks, type);
  return false;
}

void KVStoreDistServer::PushToGlobal(int key, KeyState* st, const DataHandleType& type) {
  Entry& e = st->entry;
  GlobalRound& r = st->round;
  const size_t n = e.elems;
  const float* w = e.has_master ? e.master.data() : reinterpret_cast<const float*>(e.data.data());
  const int num_gs = Postoffice::Get()->num_global_servers();
  const auto& krs = Postoffice::Get()->GetServerKeyRanges(kGlobal);
  SArray<Key> keys; SArray<char> vals; SArray<int> lens;
  int cmd;
  const CompressionType ct = gc_.type();
  if (ct == CompressionType::kBiSparse && n >= size_lower_bound_ ) {
    int k, sample, ks; GradientCompression::BSCSizes(static_cast<int64_t>(n), gc_.threshold(), &k, &sample, &ks);
    auto& u = st->bsc_u; auto& v = st->bsc_v;
    if (u.size() != n) { u.assign(n, 0.f); v.assign(n, 0.f); }
    std::vector<float> out(2 * static_cast<size_t>(k));
    gc_.BSCompress(w, u.data(), v.data(), out.data(), static_cast<int64_t>(n));
    vals.CopyFrom(reinterpret_cast<const char*>(out.data()), out.size() * sizeof(float));
    keys.push_back(krs[(key * 9973) % num_gs].begin() + static_cast<Key>(key));   // always one global server (reference :1849)
    lens.push_back(static_cast<int>(vals.size()));
    cmd = GetCommandType(RequestType::kBSCompressedPushPull, kFloat32);
  } else if (ct == CompressionType::kTwoBit && e.dtype == kFloat32) {
    auto& res = st->residual_2bit;
    if (res.size() != n) res.assign(n, 0.f);
    std::vector<uint32_t> words(GradientCompression::CompressedSize2Bit(static_cast<int64_t>(n)));
    gc_.Quantize2Bit(w, res.data(), words.data(), static_cast<int64_t>(n));
    vals.CopyFrom(reinterpret_cast<const char*>(words.data()), words.size() * sizeof(uint32_t));
    keys.push_back(krs[(key * 9973) % num_gs].begin() + static_cast<Key>(key));
    lens.push_back(static_cast<int>(vals.size()));
    cmd = GetCommandType(RequestType::kCompressedPushPull, kFloat32);
  } else {
    PSKVPlan plan = EncodeKeyPlan(kGlobal, key, n, DTypeSize(e.dtype), bigarray_bound_, CompressionPinsKey(static_cast<int>(gc_.type()), n, e.dtype, size_lower_bound_));
    for (Key k : plan.keys) keys.push_back(k);
    for (int l : plan.lens) lens.push_back(l);
    vals.CopyFrom(e.data.data(), e.data.size());
    cmd = GetCommandType(RequestType::kDefaultPushPull, e.dtype);
  }
  r.cmd = cmd;
  const bool allow_dgt = ps_server_->enable_dgt != 0 && cmd == GetCommandType(RequestType::kDefaultPushPull, kFloat32);
  // inter-party TSEngine: dense pushes are merged with other parties' aggregates on their way; the fresh value comes back by relay
  r.via_ts = ps_server_->ts(kGlobal) != nullptr && !allow_dgt && keys.size() == 1 && DepairDataHandleType(cmd).requestType == RequestType::kDefaultPushPull;
  // responses carry the key and the request's timestamp; they are matched against push_ts under this key's lock (held here), so an
  // answer that arrives before Push() returns still finds the round
  r.push_ts = ps_server_->Push(keys, vals, lens, cmd, -key, key, allow_dgt, r.via_ts);
}

void KVStoreDistServer::PullFromGlobal(int key, const DataHandleType& type) {
  KeyState& ksr = Slot(key);
  std::lock_guard<std::mutex> lk(ksr.mu);
  Entry& e = ksr.entry;
  GlobalRound& r = ksr.round;
  const size_t n = e.elems;
  const int num_gs = Postoffice::Get()->num_global_servers();
  const auto& krs = Postoffice::Get()->GetServerKeyRanges(kGlobal);
  SArray<Key> keys;
  int cmd;
  if (gc_.type() == CompressionType::kBiSparse && n >= size_lower_bound_ && r.push_ts >= 0) {
    keys.push_back(krs[(key * 9973) % num_gs].begin() + static_cast<Key>(key));
    cmd = GetCommandType(RequestType::kBSCompressedPushPull, kFloat32);
    r.parts_expected = 1;
  } else if (gc_.type() == CompressionType::kTwoBit && e.dtype == kFloat32 && r.push_ts >= 0) {
    keys.push_back(krs[(key * 9973) % num_gs].begin() + static_cast<Key>(key));
    cmd = GetCommandType(RequestType::kDefaultPushPull, e.dtype);
    r.parts_expected = 1;
  } else {
    PSKVPlan plan = EncodeKeyPlan(kGlobal, key, n, DTypeSize(e.dtype), bigarray_bound_, CompressionPinsKey(static_cast<int>(gc_.type()), n, e.dtype, size_lower_bound_));
    for (Key k : plan.keys) keys.push_back(k);
    cmd = GetCommandType(RequestType::kDefaultPushPull, e.dtype);
    r.parts_expected = static_cast<int>(plan.keys.size());
  }
  r.parts.clear();
  r.cmd = cmd;
  r.pull_ts = ps_server_->Pull(keys, cmd, -key, key);
}

// responses to OUR requests on the global plane (local server side); they carry the key and the timestamp of the request they answer
void KVStoreDistServer::ResponseHandle(const KVMeta& res, const KVPairs& data, KVServer* server) {
  const int key = res.key;
  KeyState* ks = Find(key);
  if (ks == nullptr) return;
  std::vector<Reply> out;
  std::unique_lock<std::mutex> lk(ks->mu);
  GlobalRound& r = ks->round;
  Entry& e = ks->entry;
  auto finish = [&](std::vector<float>* fresh) {       // the key's new value is complete: store it, release the workers, open the next round
    const bool was_round = r.push_ts >= 0;
    ApplyFreshFromGlobal(key, ks, fresh, &out);
    lk.unlock();
    Send(&out);
    if (was_round) RoundCompleted(key); else AskTS(key);
  };
  const auto by_key = [](const std::pair<Key, std::vector<char>>& a, const std::pair<Key, std::vector<char>>& b) { return a.first < b.first; };
  if (res.push) {
    if (res.timestamp != r.push_ts) return;             // not the round in flight (late duplicate)
    // push ack: once every global server acknowledged, fetch the fresh value (reference :941-957) — unless the acks already carried it
    // (fused inter-tier pull: every global server that owns a slice of the key answers with the post-update slice)
    if (data.vals.size() > 0 && data.keys.size()) {
      std::vector<char> bytes(data.vals.data(), data.vals.data() + data.vals.size());
      r.parts.emplace_back(data.keys[0], std::move(bytes));
    }
    if (server->NumResponse(res.timestamp) != Postoffice::Get()->num_global_servers() - 1) return;
    if (r.via_ts) { r.parts.clear(); return; }   // TSEngine: the global server relays the fresh value (OnRelayedFromGlobal)
    const DataHandleType type = DepairDataHandleType(r.cmd);
    size_t got = 0;
    for (auto& p : r.parts) got += p.second.size();
    if (type.requestType == RequestType::kBSCompressedPushPull && r.parts.size() == 1 && got > 0) {
      // fused Bi-Sparse answer: [values | indices] of the re-sparsified aggregate
      std::vector<float> recved(e.elems);
      const auto& z = r.parts[0].second;
      GradientCompression::BSCDecompress(reinterpret_cast<const float*>(z.data()), z.size() / sizeof(float), recved.data(), e.elems);
      r.parts.clear();
      finish(&recved);
      return;
    }
    if (type.requestType == RequestType::kDefaultPushPull && got > 0 && got == e.elems * DTypeSize(e.dtype)) {
      std::sort(r.parts.begin(), r.parts.end(), by_key);
      std::vector<char> whole;
      for (auto& p : r.parts) whole.insert(whole.end(), p.second.begin(), p.second.end());
      r.parts.clear();
      std::vector<float> recved(e.elems);
      ToFloat(whole.data(), e.dtype, e.elems, recved.data());
      finish(&recved);
      return;
    }
    r.parts.clear();
    lk.unlock();
    PullFromGlobal(key, type);
    return;
  }
  // pull response part
  if (res.timestamp != r.pull_ts) return;
  std::vector<char> bytes(data.vals.data(), data.vals.data() + data.vals.size());
  r.parts.emplace_back(data.keys.size() ? data.keys[0] : 0, std::move(bytes));
  if (static_cast<int>(r.parts.size()) < r.parts_expected) return;
  std::sort(r.parts.begin(), r.parts.end(), by_key);
  const size_t n = e.elems;
  const DataHandleType type = DepairDataHandleType(r.cmd);
  std::vector<float> recved(n);
  if (type.requestType == RequestType::kBSCompressedPushPull) {
    const auto& z = r.parts[0].second;
    GradientCompression::BSCDecompress(reinterpret_cast<const float*>(z.data()), z.size() / sizeof(float), recved.data(), n);
  } else {
    std::vector<char> whole;
    for (auto& p : r.parts) whole.insert(whole.end(), p.second.begin(), p.second.end());
    HIPS_CHECK_MSG(whole.size() == n * DTypeSize(e.dtype), "pull response size mismatch for key " + std::to_string(key));
    ToFloat(whole.data(), e.dtype, n, recved.data());
  }
  r.parts.clear();
  finish(&recved);
}

// the value of `key` after a global round (or the initial value) reached this local server: HFA algebra, store, release the workers
void KVStoreDistServer::ApplyFreshFromGlobal(int key, KeyState* ks, std::vector<float>* recved_p, std::vector<Reply>* out) {
  std::vector<float>& recved = *recved_p;
  Entry& e = ks->entry;
  GlobalRound& r = ks->round;
  const size_t n = e.elems;
  float* w = e.has_master ? e.master.data() : reinterpret_cast<float*>(e.data.data());
  if (use_hfa_) {
    // HandleHFAAccumulate (reference :959-972): the first pulled value becomes the milestone, afterwards stored = milestone + sum(deltas)
    auto& ms = ks->milestone;
    if (ms.size() != n) { memcpy(w, recved.data(), n * 4); ms.assign(w, w + n); }
    else { for (size_t i = 0; i < n; ++i) { w[i] = ms[i] + recved[i]; ms[i] = w[i]; } }
  } else {
    memcpy(w, recved.data(), n * sizeof(float));
  }
  if (e.has_master) StoreFromFloat(&e, w, n);
  ks->initialized = true;
  FlushParkedPulls(ks, out);
  std::vector<KVMeta> waiting; waiting.swap(r.waiting);
  r.push_ts = r.pull_ts = -1;
  r.via_ts = false;
  for (size_t i = 0; i < waiting.size(); ++i) {
    if (i > 0 && waiting[i].sender == waiting[i - 1].sender && waiting[i].timestamp == waiting[i - 1].timestamp) continue;
    if (ps_server_->enable_p3) out->push_back(StoredReply(waiting[i], key, e));
    else { Reply rep; rep.to = waiting[i]; out->push_back(std::move(rep)); }
  }
}

// inter-party TSEngine: the global server's relay delivered the fresh value of a round this local server pushed through the overlay
void KVStoreDistServer::OnRelayedFromGlobal(int key, int version, int cmd, const std::vector<char>& bytes) {
  KeyState* ks = Find(key);
  if (ks == nullptr) return;
  std::vector<Reply> out;
  std::unique_lock<std::mutex> lk(ks->mu);
  Entry& e = ks->entry;
  if (!ks->round.via_ts) return;    // not waiting for a relayed round (e.g. a duplicate)
  const size_t n = e.elems;
  if (bytes.size() != n * DTypeSize(e.dtype)) return;
  std::vector<float> recved(n);
  ToFloat(bytes.data(), e.dtype, n, recved.data());
  ApplyFreshFromGlobal(key, ks, &recved, &out);
  lk.unlock();
  Send(&out);
  RoundCompleted(key);
}

// ------------------------------------------------------------------------------------------------ TSEngine hooks
std::vector<KVMeta> KVStoreDistServer::ExpandOrigins(const KVMeta& req) {
  std::vector<KVMeta> out;
  const std::vector<TSOrigin> origins = DecodeOrigins(req.body);
  if (origins.empty()) {
    for (int i = 0; i < std::max(1, req.num_merge); ++i) out.push_back(req);
    return out;
  }
  for (const TSOrigin& o : origins) {
    KVMeta r = req;
    r.sender = o.sender; r.timestamp = o.timestamp; r.customer_id = o.customer; r.num_merge = 1; r.body.clear();
    out.push_back(r);
  }
  return out;
}

// this server is ready to receive (more) contributions of `key`: join the scheduler's pairing queue on the planes it serves
void KVStoreDistServer::AskTS(int key) {
  Postoffice* po = Postoffice::Get();
  if (TSNode* t = ps_server_->ts(kLocal)) { if (!is_global_ || po->enable_central_workers()) t->AskAsServer(key); }
  if (is_global_) if (TSNode* t = ps_server_->ts(kGlobal)) t->AskAsServer(key);
}

// ks->mu held: advance the round counter of the key.  Returns true when the SLOWEST key thereby reaches a multiple of the checkpoint period:
// the caller then writes the snapshot after releasing the key and BEFORE acknowledging the round — until the acks are out no worker can start
// round N+1 of any key, so every key holds exactly its round-N value while the snapshot is taken.
bool KVStoreDistServer::BumpRound(KeyState* ks) {
  const int version = ++ks->version;
  if (ckpt_every_ <= 0) return false;
  int slowest = version;
  for (auto& kv : Slots()) slowest = std::min(slowest, kv.second->version.load());
  std::lock_guard<std::mutex> lk(round_mu_);
  if (slowest > ckpt_key_ && slowest % ckpt_every_ == 0) { ckpt_key_ = slowest; return true; }
  return false;
}

// a synchronisation round of `key` finished on this server: bump the version (unless the caller already did), start the relay broadcast,
// open the next round
void KVStoreDistServer::RoundCompleted(int key, bool bumped) {
  std::vector<char> bytes;
  int version, cmd;
  bool checkpoint_due = false;
  const bool relays = ps_server_->ts(kLocal) != nullptr || (is_global_ && ps_server_->ts(kGlobal) != nullptr);
  {
    KeyState& ks = Slot(key);
    std::lock_guard<std::mutex> lk(ks.mu);
    if (!bumped) checkpoint_due = BumpRound(&ks);
    version = ks.version.load();
    if (relays) bytes = ks.entry.data;
    cmd = GetCommandType(RequestType::kDefaultPushPull, ks.entry.dtype);
  }
  if (checkpoint_due) SaveStates(ckpt_prefix_);
  Postoffice* po = Postoffice::Get();
  if (TSNode* t = ps_server_->ts(kLocal)) {
    if (!is_global_ || po->enable_central_workers()) t->Relay(key, version, cmd, static_cast<Key>(key), bytes.data(), bytes.size());
  }
  if (is_global_) if (TSNode* t = ps_server_->ts(kGlobal)) t->Relay(key, version, cmd, static_cast<Key>(key), bytes.data(), bytes.size());
  AskTS(key);
}

// A pull never blocks a thread: while the key has no value yet (the reference spins with sleep(100ms), :1719-1724) the request is parked with
// the key and answered by whoever initialises it.
void KVStoreDistServer::HandlePull(const DataHandleType& type, const KVMeta& req, const KVPairs& data) {
  const int key = req.key;
  KeyState& ks = Slot(key);
  std::vector<Reply> out;
  {
    std::lock_guard<std::mutex> lk(ks.mu);
    if (!ks.initialized) { ks.parked_pulls.push_back(ParkedPull{type, req, data}); return; }
    out.push_back(PullReply(&ks, type, req, data));
  }
  Send(&out);
}

// ks->mu held: the key just received its value — answer the pulls that arrived before it
void KVStoreDistServer::FlushParkedPulls(KeyState* ks, std::vector<Reply>* out) {
  for (ParkedPull& p : ks->parked_pulls) out->push_back(PullReply(ks, p.type, p.req, p.data));
  ks->parked_pulls.clear();
}

// ks->mu held
KVStoreDistServer::Reply KVStoreDistServer::PullReply(KeyState* ks, const DataHandleType& type, const KVMeta& req, const KVPairs& data) {
  Entry& e = ks->entry;
  Reply rep;
  rep.to = req;
  KVPairs& res = rep.data;
  res.keys = data.keys;
  if (res.keys.size() == 0) res.keys.push_back(static_cast<Key>(req.key));
  if (type.requestType == RequestType::kRowSparsePushPull) {
    // row_sparse pull: only the requested rows travel back (reference DataHandleRowSparse pull branch :700-756)
    HIPS_CHECK_MSG(data.vals.size() >= 2 * sizeof(int64_t), "row_sparse pull without row ids");
    const int64_t* hdr = reinterpret_cast<const int64_t*>(data.vals.data());
    const size_t nrows = static_cast<size_t>(hdr[0]), row_len = static_cast<size_t>(hdr[1]);
    const int64_t* ids = hdr + 2;
    std::vector<float> full(e.elems);
    if (e.has_master) full = e.master; else ToFloat(e.data.data(), e.dtype, e.elems, full.data());
    std::vector<float> out(nrows * row_len);
    for (size_t r = 0; r < nrows; ++r) {
      const size_t base = static_cast<size_t>(ids[r]) * row_len;
      HIPS_CHECK_MSG(ids[r] >= 0 && base + row_len <= e.elems, "row id out of range in row_sparse pull");
      memcpy(out.data() + r * row_len, full.data() + base, row_len * sizeof(float));
    }
    res.vals.CopyFrom(reinterpret_cast<const char*>(out.data()), out.size() * sizeof(float));
  } else if (type.requestType == RequestType::kBSCompressedPushPull) {
    // Bi-Sparse pull: re-sparsify the aggregate, capacity k * num_parties (reference :1190-1206)
    const int mult = std::max(1, Postoffice::Get()->num_global_workers());
    const float* w = e.has_master ? e.master.data() : reinterpret_cast<const float*>(e.data.data());
    std::vector<float> out(GradientCompression::BSCPullSize(static_cast<int64_t>(e.elems), gc_.threshold(), mult));
    gc_.BSCPullCompress(w, out.data(), static_cast<int64_t>(e.elems), mult);
    res.vals.CopyFrom(reinterpret_cast<const char*>(out.data()), out.size() * sizeof(float));
  } else {
    res.vals.CopyFrom(e.data.data(), e.data.size());
  }
  res.lens.push_back(static_cast<int>(res.vals.size()));
  return rep;
}

// ------------------------------------------------------------------------------------------------ server-state checkpoint
static void WriteVec(std::ofstream& f, const std::vector<float>& v) { uint64_t n = v.size(); f.write(reinterpret_cast<const char*>(&n), 8); if (n) f.write(reinterpret_cast<const char*>(v.data()), n * 4); }
static void ReadVec(std::ifstream& f, std::vector<float>* v) { uint64_t n = 0; f.read(reinterpret_cast<char*>(&n), 8); v->resize(n); if (n) f.read(reinterpret_cast<char*>(v->data()), n * 4); }

// <prefix>.server<r>g for global servers, <prefix>.server<r>l for local / stand-alone ones.  A local server of a two-tier job is named by
// its rank on the GLOBAL plane (= its party): every party's only server has local rank 0.
std::string KVStoreDistServer::StatePath(const std::string& prefix) const {
  Postoffice* po = Postoffice::Get();
  const int r = po->my_rank((is_global_ || has_global_) ? kGlobal : kLocal);
  return prefix + ".server" + std::to_string(r) + (is_global_ ? "g" : "l");
}

// One key at a time under its own lock.  The periodic snapshot is taken at a quiescent point (see BumpRound); an explicit kSaveStates
// command is issued by the job between rounds.
void KVStoreDistServer::SaveStates(const std::string& prefix) {
  const std::string path = StatePath(prefix);
  const std::string tmp = path + ".tmp";
  std::ofstream f(tmp, std::ios::binary);
  const uint64_t magic = 0x4869505353544154ull;  // "HiPSSTAT"
  f.write(reinterpret_cast<const char*>(&magic), 8);
  const auto slots = Slots();
  uint64_t nk = slots.size(); f.write(reinterpret_cast<const char*>(&nk), 8);
  for (auto& kv : slots) {
    KeyState& ks = *kv.second;
    std::lock_guard<std::mutex> lk(ks.mu);
    const Entry& e = ks.entry;
    int32_t key = kv.first, dtype = e.dtype; uint64_t elems = e.elems;
    f.write(reinterpret_cast<const char*>(&key), 4); f.write(reinterpret_cast<const char*>(&dtype), 4); f.write(reinterpret_cast<const char*>(&elems), 8);
    uint64_t nb = e.data.size(); f.write(reinterpret_cast<const char*>(&nb), 8); f.write(e.data.data(), nb);
    WriteVec(f, e.master);
    WriteVec(f, ks.milestone);
    WriteVec(f, ks.bsc_u);
    WriteVec(f, ks.bsc_v);
    WriteVec(f, ks.residual_2bit);
    int32_t t = ks.opt.t; f.write(reinterpret_cast<const char*>(&t), 4);
    WriteVec(f, ks.opt.a); WriteVec(f, ks.opt.b);
  }
  int64_t li = local_iters_.load(); f.write(reinterpret_cast<const char*>(&li), 8);
  f.close();
  HIPS_CHECK_MSG(std::rename(tmp.c_str(), path.c_str()) == 0, "cannot move " + tmp + " to " + path);   // readers never see a torn file
}

// GEOMX_SERVER_RESUME=1: a (re)started global / stand-alone server adopts the last periodic checkpoint if there is one.  The job's scripts
// still call kv.init for every key; those pushes are acknowledged without touching the restored values (KeyState::skip_init_push).
void KVStoreDistServer::TryResume() {
  const std::string path = StatePath(ckpt_prefix_);
  std::ifstream probe(path, std::ios::binary);
  if (!probe.good()) return;
  probe.close();
  LoadStates(ckpt_prefix_);
  const auto slots = Slots();
  for (auto& kv : slots) { std::lock_guard<std::mutex> lk(kv.second->mu); kv.second->skip_init_push = true; }
  any_skip_init_ = !slots.empty();
  fprintf(stderr, "[hips] server resumed %zu keys from %s\n", slots.size(), path.c_str());
}

void KVStoreDistServer::LoadStates(const std::string& prefix) {
  const std::string path = StatePath(prefix);
  std::ifstream f(path, std::ios::binary);
  HIPS_CHECK_MSG(f.good(), "cannot open " + path);
  uint64_t magic = 0, nk = 0;
  f.read(reinterpret_cast<char*>(&magic), 8); f.read(reinterpret_cast<char*>(&nk), 8);
  HIPS_CHECK(magic == 0x4869505353544154ull);
  std::vector<Reply> late;       // pulls that were waiting for a key this file brings
  for (uint64_t i = 0; i < nk; ++i) {
    int32_t key, dtype; uint64_t elems, nb;
    f.read(reinterpret_cast<char*>(&key), 4); f.read(reinterpret_cast<char*>(&dtype), 4); f.read(reinterpret_cast<char*>(&elems), 8);
    f.read(reinterpret_cast<char*>(&nb), 8);
    KeyState& ks = Slot(key);
    std::lock_guard<std::mutex> lk(ks.mu);
    Entry& e = ks.entry;
    e.dtype = dtype; e.elems = elems; e.data.resize(nb); f.read(e.data.data(), nb);
    ReadVec(f, &e.master); e.has_master = !e.master.empty();
    ReadVec(f, &ks.milestone);
    ReadVec(f, &ks.bsc_u);
    ReadVec(f, &ks.bsc_v);
    ReadVec(f, &ks.residual_2bit);
    int32_t t; f.read(reinterpret_cast<char*>(&t), 4);
    ks.opt.t = t; ReadVec(f, &ks.opt.a); ReadVec(f, &ks.opt.b);     // used by whichever native optimizer is (or will be) configured
    ks.initialized = true;
    std::vector<Reply> parked;
    FlushParkedPulls(&ks, &parked);
    late.insert(late.end(), std::make_move_iterator(parked.begin()), std::make_move_iterator(parked.end()));
  }
  Send(&late);
  int64_t li = 0; f.read(reinterpret_cast<char*>(&li), 8); local_iters_ = li;
  for (auto& kv : Slots()) { std::lock_guard<std::mutex> lk(kv.second->mu); kv.second->local_rounds = li; }
}

}  // namespace hips
