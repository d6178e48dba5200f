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
  ps_server_.reset(new KVServer(0));
  ps_server_->SimpleApp::set_request_handle([this](const SimpleData& d, SimpleApp* app) { CommandHandle(d, app); });
  ps_server_->set_request_handle([this](const KVMeta& m, const KVPairs& d, KVServer* s) { DataHandleEx(m, d, s); });
  ps_server_->set_response_handle([this](const KVMeta& m, const KVPairs& d, KVServer* s) { ResponseHandle(m, d, s); });
  if (has_global_ && ps_server_->ts(kGlobal) != nullptr)   // local server: fresh values of TS rounds arrive through the relay
    ps_server_->ts(kGlobal)->set_on_relayed([this](int key, int version, int cmd, const std::vector<char>& bytes) { OnRelayedFromGlobal(key, version, cmd, bytes); });
}

KVStoreDistServer::~KVStoreDistServer() { ps_server_.reset(); }

int KVStoreDistServer::rank_local() { return Postoffice::Get()->my_rank(kLocal); }

std::vector<float> KVStoreDistServer::GetStored(int key) {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = store_.find(key);
  if (it == store_.end()) return {};
  std::vector<float> out(it->second.elems);
  if (it->second.has_master) out = it->second.master;
  else ToFloat(it->second.data.data(), it->second.dtype, it->second.elems, out.data());
  return out;
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
        { std::lock_guard<std::mutex> lk(mu_); if (recved.plane == kGlobal) stop = (++stop_votes_ == po->num_global_workers()); }
        if (stop) exec_.Stop();
      } else {
        bool first;
        { std::lock_guard<std::mutex> lk(mu_); first = !stop_requested_; stop_requested_ = true; }
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
      std::lock_guard<std::mutex> lk(mu_);
      if (!multi_precision_) {
        multi_precision_ = true;
        for (auto& kv : store_) {
          Entry& e = kv.second;
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
      std::lock_guard<std::mutex> lk(mu_);
      OptSpec s = OptSpec::Parse(recved.body);
      HIPS_CHECK_MSG(s.valid(), "unknown native optimizer spec: " + recved.body);
      native_opt_.reset(new NativeOptimizer(s));
      for (auto& kv : resumed_opt_) native_opt_->states()[kv.first] = kv.second;   // state checkpointed by a previous incarnation
      resumed_opt_.clear();
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
  ProfileScope ps(req.push ? "KVStoreDistServerPush" : "KVStoreDistServerPull");
  if (req.push) HandlePush(type, req, data);
  else HandlePull(type, req, data);
}

void KVStoreDistServer::ApplyUpdate(int key, Entry* e, const float* grad, size_t n) {
  float* w = e->has_master ? e->master.data() : reinterpret_cast<float*>(e->data.data());
  if (native_opt_) native_opt_->Update(key, w, grad, n);
  else if (updater_) exec_.Exec([this, key, grad, w, n]() { updater_(key, grad, w, n); });
  else memcpy(w, grad, n * sizeof(float));  // no optimizer on the server: store the aggregate (reference ApplyUpdates :547-550)
  if (e->has_master) StoreFromFloat(e, e->master.data(), n);
}

void KVStoreDistServer::HandlePush(const DataHandleType& type, const KVMeta& req, const KVPairs& data) {
  HIPS_CHECK(data.keys.size() == 1);
  const int key = req.key;
  Postoffice* po = Postoffice::Get();
  num_pushes_++;
  std::unique_lock<std::mutex> lk(mu_);
  Entry& e = store_[key];
  const bool p3 = ps_server_->enable_p3;
  // inter-tier fusion (GEOMX_FUSED_TIER_PULL, default on): a global server answers a local server's dense push with the post-update value,
  // so the local server does not need a second round trip (push ack, then pull) over the slow link between parties
  const bool fuse_up = is_global_ && fused_tier_pull_ && type.requestType == RequestType::kDefaultPushPull;
  const bool fuse_bsc = is_global_ && fused_tier_pull_ && type.requestType == RequestType::kBSCompressedPushPull && sync_global_mode_;
  auto respond = [&](const KVMeta& r) {
    if (fuse_up && r.plane == kGlobal && r.sender % 2 == 1) {
      KVPairs res; res.keys = data.keys;
      res.vals.CopyFrom(e.data.data(), e.data.size());
      res.lens.push_back(static_cast<int>(e.data.size()));
      ps_server_->Response(r, res);
    } else if (fuse_bsc && r.plane == kGlobal && r.sender % 2 == 1) {
      // Bi-Sparse: the answer is the re-sparsified aggregate a pull would have returned (capacity k * parties, reference :1190-1206)
      const int mult = std::max(1, Postoffice::Get()->num_global_workers());
      const float* w = e.has_master ? e.master.data() : reinterpret_cast<const float*>(e.data.data());
      std::vector<float> out(GradientCompression::BSCPullSize(static_cast<int64_t>(e.elems), gc_.threshold(), mult));
      gc_.BSCPullCompress(w, out.data(), static_cast<int64_t>(e.elems), mult);
      KVPairs res; res.keys = data.keys;
      res.vals.CopyFrom(reinterpret_cast<const char*>(out.data()), out.size() * sizeof(float));
      res.lens.push_back(static_cast<int>(res.vals.size()));
      ps_server_->Response(r, res);
    } else if (p3 && !is_global_) {
      KVPairs res; res.keys = data.keys;
      res.vals.CopyFrom(e.data.data(), e.data.size());
      res.lens.push_back(static_cast<int>(e.data.size()));
      ps_server_->Response(r, res);
    } else ps_server_->Response(r);
  };
  if (e.elems != 0 && !skip_init_push_.empty() && (standalone_ || req.plane == kLocal)) {
    // ---- resumed server: the restarted job initialises its keys again (kv.init); the checkpointed value wins, the push is only acknowledged.
    // On a global server the init comes from the master worker over the LOCAL plane of the central party — the parties may already be
    // training by then (their pulls found the restored keys initialised), so their pushes on the global plane must not be mistaken for it.
    auto sk = skip_init_push_.find(key);
    if (sk != skip_init_push_.end() && sk->second) {
      sk->second = false;
      respond(req);
      lk.unlock();
      AskTS(key);
      return;
    }
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
    if (is_global_ || standalone_) { initialized_[key] = true; init_cv_.notify_all(); }
    else if (has_global_) { lk.unlock(); PullFromGlobal(key, type); AskTS(key); return; }
    lk.unlock();
    AskTS(key);      // TSEngine: open the first round of this key
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
  if (is_global_ && req.plane == kLocal && !po->enable_central_workers()) { respond(req); return; }

  const bool sync = (is_global_ ? sync_global_mode_ : sync_mode_);
  if (!sync) {
    // ---- MixedSync / async: apply this contribution immediately (reference :1582-1609)
    ApplyUpdate(key, &e, inc, n);
    const std::vector<KVMeta> who = ExpandOrigins(req);
    for (size_t i = 0; i < who.size(); ++i) {
      if (i > 0 && who[i].sender == who[i - 1].sender && who[i].timestamp == who[i - 1].timestamp) continue;
      respond(who[i]);
    }
    lk.unlock();
    AskTS(key);
    return;
  }
  UpdateBuf& ub = update_buf_[key];
  if (ub.request.empty()) ub.merged.assign(inc, inc + n);
  else { float* m = ub.merged.data(); for (size_t i = 0; i < n; ++i) m[i] += inc[i]; }
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
    ApplyUpdate(key, &e, ub.merged.data(), n);
    // the round counter moves (and a due periodic checkpoint is written) BEFORE any push of this round is acknowledged: once the acks are
    // out, workers may start round N+1 of other keys, and a snapshot taken later could mix round-N and round-N+1 state
    BumpRoundLocked(key);
    std::vector<KVMeta> reqs; reqs.swap(ub.request);
    for (size_t i = 0; i < reqs.size(); ++i) {
      if (i > 0 && reqs[i].sender == reqs[i - 1].sender && reqs[i].timestamp == reqs[i - 1].timestamp) continue;  // merged duplicates (TS)
      respond(reqs[i]);
    }
    lk.unlock();
    RoundCompleted(key, /*bumped=*/true);
    return;
  }
  const bool local_round_done = FinishLocalAggregation(key, type, &ub);
  lk.unlock();
  if (local_round_done) RoundCompleted(key);
}

// local server: all workers of the party have pushed `key`
bool KVStoreDistServer::FinishLocalAggregation(int key, const DataHandleType& type, UpdateBuf* ub) {
  Entry& e = store_[key];
  const size_t n = e.elems;
  float* w = e.has_master ? e.master.data() : reinterpret_cast<float*>(e.data.data());
  memcpy(w, ub->merged.data(), n * sizeof(float));   // only aggregate (ApplyUpdates on a non-global server)
  if (e.has_master) StoreFromFloat(&e, w, n);
  if (key == 0) ++local_iters_;                      // HFA counts local rounds on key 0 (reference :1324)
  auto ack_all = [&](std::vector<KVMeta>* reqs) {
    for (size_t i = 0; i < reqs->size(); ++i) {
      if (i > 0 && (*reqs)[i].sender == (*reqs)[i - 1].sender && (*reqs)[i].timestamp == (*reqs)[i - 1].timestamp) continue;
      if (ps_server_->enable_p3) {
        KVPairs res; res.keys.push_back(static_cast<Key>(key)); res.vals.CopyFrom(e.data.data(), e.data.size()); res.lens.push_back(static_cast<int>(e.data.size()));
        ps_server_->Response((*reqs)[i], res);
      } else ps_server_->Response((*reqs)[i]);
    }
    reqs->clear();
  };
  if (use_hfa_ && (local_iters_ % hfa_k2_ != 0)) {   // local synchronisation only
    ack_all(&ub->request);
    return true;
  }
  if (use_hfa_) {                                    // push the party's progress since the last global sync
    auto& ms = milestone_[key];
    HIPS_CHECK_MSG(ms.size() == n, "HFA milestone not initialised for key " + std::to_string(key));
    const float inv = 1.f / Postoffice::Get()->num_global_workers();
    for (size_t i = 0; i < n; ++i) w[i] = (w[i] - ms[i]) * inv;
    if (e.has_master) StoreFromFloat(&e, w, n);
  }
  GlobalRound& r = rounds_[key];
  r.waiting.swap(ub->request);
  PushToGlobal(key, type);
  return false;
}

void KVStoreDistServer::PushToGlobal(int key, const DataHandleType& type) {
  Entry& e = store_[key];
  GlobalRound& r = rounds_[key];
  const size_t n = e.elems;
  const float* w = e.has_master ? e.master.data() : reinterpret_cast<const float*>(e.data.data());
  const int num_gs = Postoffice::Get()->num_global_servers();
  const auto& krs = Postoffice::Get()->GetServerKeyRanges(kGlobal);
  SArray<Key> keys; SArray<char> vals; SArray<int> lens;
  int cmd;
  const CompressionType ct = gc_.type();
  if (ct == CompressionType::kBiSparse && n >= size_lower_bound_ ) {
    int k, sample, ks; GradientCompression::BSCSizes(static_cast<int64_t>(n), gc_.threshold(), &k, &sample, &ks);
    auto& u = bsc_u_[key]; auto& v = bsc_v_[key];
    if (u.size() != n) { u.assign(n, 0.f); v.assign(n, 0.f); }
    std::vector<float> out(2 * static_cast<size_t>(k));
    gc_.BSCompress(w, u.data(), v.data(), out.data(), static_cast<int64_t>(n));
    vals.CopyFrom(reinterpret_cast<const char*>(out.data()), out.size() * sizeof(float));
    keys.push_back(krs[(key * 9973) % num_gs].begin() + static_cast<Key>(key));   // always one global server (reference :1849)
    lens.push_back(static_cast<int>(vals.size()));
    cmd = GetCommandType(RequestType::kBSCompressedPushPull, kFloat32);
  } else if (ct == CompressionType::kTwoBit && e.dtype == kFloat32) {
    auto& res = residual_2bit_[key];
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
  r.push_ts = ps_server_->Push(keys, vals, lens, cmd, -key, key, allow_dgt, r.via_ts);
  ts_key_[r.push_ts] = key;
}

void KVStoreDistServer::PullFromGlobal(int key, const DataHandleType& type) {
  std::lock_guard<std::mutex> lk(mu_);
  Entry& e = store_[key];
  GlobalRound& r = rounds_[key];
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
  ts_key_[r.pull_ts] = key;
}

// responses to OUR requests on the global plane (local server side)
void KVStoreDistServer::ResponseHandle(const KVMeta& res, const KVPairs& data, KVServer* server) {
  std::unique_lock<std::mutex> lk(mu_);
  auto it = ts_key_.find(res.timestamp);
  if (it == ts_key_.end()) return;
  const int key = it->second;
  GlobalRound& r = rounds_[key];
  if (res.push) {
    // push ack: once every global server acknowledged, fetch the fresh value (reference :941-957) — unless the acks already carried it
    // (fused inter-tier pull: every global server that owns a slice of the key answers with the post-update slice)
    if (data.vals.size() > 0 && data.keys.size()) {
      std::vector<char> bytes(data.vals.data(), data.vals.data() + data.vals.size());
      r.parts.emplace_back(data.keys[0], std::move(bytes));
    }
    if (server->NumResponse(res.timestamp) != Postoffice::Get()->num_global_servers() - 1) return;
    ts_key_.erase(it);
    if (r.via_ts) { r.parts.clear(); return; }   // TSEngine: the global server relays the fresh value (OnRelayedFromGlobal)
    const DataHandleType type = DepairDataHandleType(r.cmd);
    Entry& e = store_[key];
    size_t got = 0;
    for (auto& p : r.parts) got += p.second.size();
    if (type.requestType == RequestType::kBSCompressedPushPull && r.parts.size() == 1 && got > 0) {
      // fused Bi-Sparse answer: [values | indices] of the re-sparsified aggregate
      std::vector<float> recved(e.elems);
      const auto& z = r.parts[0].second;
      GradientCompression::BSCDecompress(reinterpret_cast<const float*>(z.data()), z.size() / sizeof(float), recved.data(), e.elems);
      r.parts.clear();
      const bool was_round = r.push_ts >= 0;
      ApplyFreshFromGlobal(key, &recved);
      lk.unlock();
      if (was_round) RoundCompleted(key); else AskTS(key);
      return;
    }
    if (type.requestType == RequestType::kDefaultPushPull && got > 0 && got == e.elems * DTypeSize(e.dtype)) {
      std::sort(r.parts.begin(), r.parts.end(), [](const std::pair<Key, std::vector<char>>& a, const std::pair<Key, std::vector<char>>& b) { return a.first < b.first; });
      std::vector<char> whole;
      for (auto& p : r.parts) whole.insert(whole.end(), p.second.begin(), p.second.end());
      r.parts.clear();
      std::vector<float> recved(e.elems);
      ToFloat(whole.data(), e.dtype, e.elems, recved.data());
      const bool was_round = r.push_ts >= 0;
      ApplyFreshFromGlobal(key, &recved);
      lk.unlock();
      if (was_round) RoundCompleted(key); else AskTS(key);
      return;
    }
    r.parts.clear();
    lk.unlock();
    PullFromGlobal(key, type);
    return;
  }
  // pull response part
  std::vector<char> bytes(data.vals.data(), data.vals.data() + data.vals.size());
  r.parts.emplace_back(data.keys.size() ? data.keys[0] : 0, std::move(bytes));
  if (static_cast<int>(r.parts.size()) < r.parts_expected) return;
  ts_key_.erase(it);
  std::sort(r.parts.begin(), r.parts.end(), [](const std::pair<Key, std::vector<char>>& a, const std::pair<Key, std::vector<char>>& b) { return a.first < b.first; });
  Entry& e = store_[key];
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
  const bool was_round = r.push_ts >= 0;
  ApplyFreshFromGlobal(key, &recved);
  lk.unlock();
  if (was_round) RoundCompleted(key); else AskTS(key);
}

// the value of `key` after a global round (or the initial value) reached this local server: HFA algebra, store, release the workers
void KVStoreDistServer::ApplyFreshFromGlobal(int key, std::vector<float>* recved_p) {
  std::vector<float>& recved = *recved_p;
  Entry& e = store_[key];
  GlobalRound& r = rounds_[key];
  const size_t n = e.elems;
  float* w = e.has_master ? e.master.data() : reinterpret_cast<float*>(e.data.data());
  if (use_hfa_) {
    // HandleHFAAccumulate (reference :959-972): the first pulled value becomes the milestone, afterwards stored = milestone + sum(deltas)
    auto& ms = milestone_[key];
    if (ms.size() != n) { memcpy(w, recved.data(), n * 4); ms.assign(w, w + n); }
    else { for (size_t i = 0; i < n; ++i) { w[i] = ms[i] + recved[i]; ms[i] = w[i]; } }
  } else {
    memcpy(w, recved.data(), n * sizeof(float));
  }
  if (e.has_master) StoreFromFloat(&e, w, n);
  initialized_[key] = true;
  init_cv_.notify_all();
  std::vector<KVMeta> waiting; waiting.swap(r.waiting);
  r.push_ts = r.pull_ts = -1;
  r.via_ts = false;
  for (size_t i = 0; i < waiting.size(); ++i) {
    if (i > 0 && waiting[i].sender == waiting[i - 1].sender && waiting[i].timestamp == waiting[i - 1].timestamp) continue;
    if (ps_server_->enable_p3) {
      KVPairs out; out.keys.push_back(static_cast<Key>(key)); out.vals.CopyFrom(e.data.data(), e.data.size()); out.lens.push_back(static_cast<int>(e.data.size()));
      ps_server_->Response(waiting[i], out);
    } else ps_server_->Response(waiting[i]);
  }
}

// inter-party TSEngine: the global server's relay delivered the fresh value of a round this local server pushed through the overlay
void KVStoreDistServer::OnRelayedFromGlobal(int key, int version, int cmd, const std::vector<char>& bytes) {
  std::unique_lock<std::mutex> lk(mu_);
  auto it = store_.find(key);
  if (it == store_.end()) return;
  Entry& e = it->second;
  GlobalRound& r = rounds_[key];
  if (!r.via_ts) return;            // not waiting for a relayed round (e.g. a duplicate)
  const size_t n = e.elems;
  if (bytes.size() != n * DTypeSize(e.dtype)) return;
  std::vector<float> recved(n);
  ToFloat(bytes.data(), e.dtype, n, recved.data());
  ApplyFreshFromGlobal(key, &recved);
  lk.unlock();
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

// a synchronisation round of `key` finished on this server: bump the version, start the relay broadcast, open the next round
// caller holds mu_: advance the round counter of `key`; when the SLOWEST key thereby reaches a multiple of the checkpoint period, write the
// snapshot right here — at this instant every key holds exactly its round-N value and nothing of round N has been acknowledged yet
void KVStoreDistServer::BumpRoundLocked(int key) {
  const int version = ++round_version_[key];
  if (ckpt_every_ <= 0) return;
  int slowest = version;
  for (auto& kv : store_) { auto it = round_version_.find(kv.first); slowest = std::min(slowest, it == round_version_.end() ? 0 : it->second); }
  if (slowest > ckpt_key_ && slowest % ckpt_every_ == 0) { ckpt_key_ = slowest; SaveStatesLocked(ckpt_prefix_); }
}

void KVStoreDistServer::RoundCompleted(int key, bool bumped) {
  std::vector<char> bytes;
  int version, cmd;
  {
    std::lock_guard<std::mutex> lk(mu_);
    if (!bumped) BumpRoundLocked(key);
    version = round_version_[key];
    Entry& e = store_[key];
    bytes = e.data;
    cmd = GetCommandType(RequestType::kDefaultPushPull, e.dtype);
  }
  Postoffice* po = Postoffice::Get();
  if (TSNode* t = ps_server_->ts(kLocal)) {
    if (!is_global_ || po->enable_central_workers()) t->Relay(key, version, cmd, static_cast<Key>(key), bytes.data(), bytes.size());
  }
  if (is_global_) if (TSNode* t = ps_server_->ts(kGlobal)) t->Relay(key, version, cmd, static_cast<Key>(key), bytes.data(), bytes.size());
  AskTS(key);
}

void KVStoreDistServer::HandlePull(const DataHandleType& type, const KVMeta& req, const KVPairs& data) {
  const int key = req.key;
  std::unique_lock<std::mutex> lk(mu_);
  // the pull thread waits until the key has been initialised (reference spins with sleep(100ms) :1719-1724)
  init_cv_.wait(lk, [this, key] { auto it = initialized_.find(key); return it != initialized_.end() && it->second; });
  Entry& e = store_[key];
  KVPairs res;
  res.keys = data.keys;
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
  lk.unlock();
  ps_server_->Response(req, res);
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

void KVStoreDistServer::SaveStates(const std::string& prefix) {
  std::lock_guard<std::mutex> lk(mu_);
  SaveStatesLocked(prefix);
}

void KVStoreDistServer::SaveStatesLocked(const std::string& prefix) {
  const std::string path = StatePath(prefix);
  const std::string tmp = path + ".tmp";
  std::ofstream f(tmp, std::ios::binary);
  const uint64_t magic = 0x4869505353544154ull;  // "HiPSSTAT"
  f.write(reinterpret_cast<const char*>(&magic), 8);
  uint64_t nk = store_.size(); f.write(reinterpret_cast<const char*>(&nk), 8);
  for (auto& kv : store_) {
    int32_t key = kv.first, dtype = kv.second.dtype; uint64_t elems = kv.second.elems;
    f.write(reinterpret_cast<const char*>(&key), 4); f.write(reinterpret_cast<const char*>(&dtype), 4); f.write(reinterpret_cast<const char*>(&elems), 8);
    uint64_t nb = kv.second.data.size(); f.write(reinterpret_cast<const char*>(&nb), 8); f.write(kv.second.data.data(), nb);
    WriteVec(f, kv.second.master);
    WriteVec(f, milestone_.count(key) ? milestone_[key] : std::vector<float>());
    WriteVec(f, bsc_u_.count(key) ? bsc_u_[key] : std::vector<float>());
    WriteVec(f, bsc_v_.count(key) ? bsc_v_[key] : std::vector<float>());
    WriteVec(f, residual_2bit_.count(key) ? residual_2bit_[key] : std::vector<float>());
    NativeOptimizer::State st;
    if (native_opt_ && native_opt_->states().count(key)) st = native_opt_->states()[key];
    int32_t t = st.t; f.write(reinterpret_cast<const char*>(&t), 4);
    WriteVec(f, st.a); WriteVec(f, st.b);
  }
  int64_t li = local_iters_; f.write(reinterpret_cast<const char*>(&li), 8);
  f.close();
  HIPS_CHECK_MSG(std::rename(tmp.c_str(), path.c_str()) == 0, "cannot move " + tmp + " to " + path);   // readers never see a torn file
}

// GEOMX_SERVER_RESUME=1: a (re)started global / stand-alone server adopts the last periodic checkpoint if there is one.  The job's scripts
// still call kv.init for every key; those pushes are acknowledged without touching the restored values (skip_init_push_).
void KVStoreDistServer::TryResume() {
  const std::string path = StatePath(ckpt_prefix_);
  std::ifstream probe(path, std::ios::binary);
  if (!probe.good()) return;
  probe.close();
  LoadStates(ckpt_prefix_);
  std::lock_guard<std::mutex> lk(mu_);
  for (auto& kv : store_) skip_init_push_[kv.first] = true;
  fprintf(stderr, "[hips] server resumed %zu keys from %s\n", store_.size(), path.c_str());
}

void KVStoreDistServer::LoadStates(const std::string& prefix) {
  std::lock_guard<std::mutex> lk(mu_);
  const std::string path = StatePath(prefix);
  std::ifstream f(path, std::ios::binary);
  HIPS_CHECK_MSG(f.good(), "cannot open " + path);
  uint64_t magic = 0, nk = 0;
  f.read(reinterpret_cast<char*>(&magic), 8); f.read(reinterpret_cast<char*>(&nk), 8);
  HIPS_CHECK(magic == 0x4869505353544154ull);
  for (uint64_t i = 0; i < nk; ++i) {
    int32_t key, dtype; uint64_t elems, nb;
    f.read(reinterpret_cast<char*>(&key), 4); f.read(reinterpret_cast<char*>(&dtype), 4); f.read(reinterpret_cast<char*>(&elems), 8);
    f.read(reinterpret_cast<char*>(&nb), 8);
    Entry& e = store_[key];
    e.dtype = dtype; e.elems = elems; e.data.resize(nb); f.read(e.data.data(), nb);
    ReadVec(f, &e.master); e.has_master = !e.master.empty();
    std::vector<float> v;
    ReadVec(f, &v); if (!v.empty()) milestone_[key] = v;
    ReadVec(f, &v); if (!v.empty()) bsc_u_[key] = v;
    ReadVec(f, &v); if (!v.empty()) bsc_v_[key] = v;
    ReadVec(f, &v); if (!v.empty()) residual_2bit_[key] = v;
    int32_t t; f.read(reinterpret_cast<char*>(&t), 4);
    NativeOptimizer::State st; st.t = t; ReadVec(f, &st.a); ReadVec(f, &st.b);
    if (!st.a.empty() || st.t != 0) {
      if (native_opt_) native_opt_->states()[key] = st;
      else resumed_opt_[key] = st;                     // the optimizer spec arrives later (set_optimizer command)
    }
    initialized_[key] = true;
  }
  int64_t li = 0; f.read(reinterpret_cast<char*>(&li), 8); local_iters_ = li;
  init_cv_.notify_all();
}

}  // namespace hips
