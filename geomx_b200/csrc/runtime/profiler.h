// Native profiler: lock-light event recorder that dumps chrome://tracing JSON.
// Parity: src/profiler/profiler.{h,cc} (Profiler singleton, states run/stop + pause, DeviceStats per-thread queues, DumpProfile emitting
// {"traceEvents":[...]} with ph codes B/E/X/i/C and pid = device, :155-254; aggregate table aggregate_stats.cc; continuous dump timer
// :258-296) and the server-side command path kSetProfilerParams (kvstore_dist_server.h:409-456: filename prefixed by rank<r>_).
#pragma once
#include <atomic>
#include <chrono>
#include <cstdio>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace hips {

class Profiler {
 public:
  static Profiler* Get() { static Profiler p; return &p; }
  struct Event { std::string name, cat; char ph; double ts_us, dur_us; int pid, tid; double value; };

  void SetConfig(const std::string& filename, bool aggregate, bool continuous_dump, double dump_period) {
    std::lock_guard<std::mutex> lk(mu_);
    filename_ = filename; aggregate_ = aggregate; continuous_ = continuous_dump; period_ = dump_period;
  }
  void SetState(bool run) {
    running_ = run;
    if (run && continuous_ && !dumper_.joinable()) {
      stop_dumper_ = false;
      dumper_ = std::thread([this] {
        while (!stop_dumper_) { std::this_thread::sleep_for(std::chrono::milliseconds(static_cast<int>(period_ * 1000))); if (running_) Dump(false); }
      });
    }
  }
  void Pause(bool p) { paused_ = p; }
  bool active() const { return running_ && !paused_; }
  static double NowUs() {
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
  }
  void Add(const std::string& name, const std::string& cat, char ph, double ts_us, double dur_us = 0, int pid = 0, int tid = 0, double value = 0) {
    if (!active()) return;
    std::lock_guard<std::mutex> lk(mu_);
    events_.push_back(Event{name, cat, ph, ts_us, dur_us, pid, tid, value});
  }
  size_t size() { std::lock_guard<std::mutex> lk(mu_); return events_.size(); }
  std::string filename() { std::lock_guard<std::mutex> lk(mu_); return filename_; }

  static std::string Escape(const std::string& s) {
    std::string o;
    for (char c : s) { if (c == '"' || c == '\\') o.push_back('\\'); o.push_back(c); }
    return o;
  }
  void Dump(bool finished) {
    std::vector<Event> evs;
    std::string fn;
    { std::lock_guard<std::mutex> lk(mu_); evs = events_; fn = filename_; }
    FILE* f = fopen(fn.c_str(), "w");
    if (!f) return;
    fprintf(f, "{\"traceEvents\":[\n");
    bool first = true;
    for (const auto& e : evs) {
      if (!first) fprintf(f, ",\n");
      first = false;
      fprintf(f, "{\"name\":\"%s\",\"cat\":\"%s\",\"ph\":\"%c\",\"ts\":%.3f,\"pid\":%d,\"tid\":%d", Escape(e.name).c_str(), Escape(e.cat).c_str(), e.ph,
              e.ts_us, e.pid, e.tid);
      if (e.ph == 'X') fprintf(f, ",\"dur\":%.3f", e.dur_us);
      if (e.ph == 'C') fprintf(f, ",\"args\":{\"%s\":%g}", Escape(e.name).c_str(), e.value);
      if (e.ph == 'i') fprintf(f, ",\"s\":\"p\"");
      fprintf(f, "}");
    }
    fprintf(f, "\n],\"displayTimeUnit\":\"ms\"}\n");
    fclose(f);
    if (finished) { running_ = false; }
  }
  // aggregate table: name -> (count, total, min, max)
  std::string AggregateTable() {
    std::map<std::string, std::vector<double>> agg;
    { std::lock_guard<std::mutex> lk(mu_); for (auto& e : events_) if (e.ph == 'X') agg[e.name].push_back(e.dur_us); }
    std::string out = "Name                                      Count      Total(us)      Min(us)      Max(us)      Avg(us)\n";
    char line[256];
    for (auto& kv : agg) {
      double tot = 0, mn = 1e300, mx = 0;
      for (double d : kv.second) { tot += d; mn = d < mn ? d : mn; mx = d > mx ? d : mx; }
      snprintf(line, sizeof(line), "%-40s %6zu %14.1f %12.1f %12.1f %12.1f\n", kv.first.c_str(), kv.second.size(), tot, mn, mx, tot / kv.second.size());
      out += line;
    }
    return out;
  }
  void Clear() { std::lock_guard<std::mutex> lk(mu_); events_.clear(); }
  ~Profiler() { stop_dumper_ = true; if (dumper_.joinable()) dumper_.join(); }

 private:
  std::mutex mu_;
  std::vector<Event> events_;
  std::string filename_ = "profile.json";
  bool aggregate_ = false, continuous_ = false;
  double period_ = 1.0;
  std::atomic<bool> running_{false}, paused_{false}, stop_dumper_{false};
  std::thread dumper_;
};

struct ProfileScope {
  ProfileScope(const char* name, const char* cat = "kvstore") : name_(name), cat_(cat), on_(Profiler::Get()->active()), t0_(on_ ? Profiler::NowUs() : 0) {}
  ~ProfileScope() { if (on_) Profiler::Get()->Add(name_, cat_, 'X', t0_, Profiler::NowUs() - t0_); }
  const char* name_; const char* cat_; bool on_; double t0_;
};

}  // namespace hips
