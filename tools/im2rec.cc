// Native pass-through packer: image list (.lst) -> RecordIO database (.rec) + index (.idx), no image codec and no Python involved.
//
//   g++ -O2 -std=c++17 -pthread tools/im2rec.cc -o im2rec
//   ./im2rec data/train.lst images/ data/train.rec [pack_label=0] [nsplit=1] [part=0] [threads=4]
//
// Parity: tools/im2rec.cc of the reference in its `unchanged=1` mode (the file bytes go into the record as they are); label packing, nsplit /
// part partitioning and the record layout are the same:   IRHeader { uint32 flag; float label; uint64 image_id[2]; } [flag floats] payload
// (src/io/image_recordio.h:40-80) inside dmlc RecordIO framing.  Resizing / re-encoding needs a codec: tools/im2rec.py does that with PIL.
// Differences: files are read by a pool of threads that run ahead of the writer through a bounded window (the reference reads and encodes on
// one thread), and an index file (key <TAB> offset per record, what MXIndexedRecordIO reads) is always written.
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

namespace {
constexpr uint32_t kMagic = 0xced7230a;

struct Item { uint64_t index; std::vector<float> labels; std::string path; };

bool ParseLine(const std::string& line, Item* it) {
  std::vector<std::string> f;
  std::stringstream ss(line);
  std::string tok;
  while (std::getline(ss, tok, '\t')) f.push_back(tok);
  if (f.size() < 3) return false;
  try {
    it->index = std::stoull(f[0]);
    it->labels.clear();
    for (size_t i = 1; i + 1 < f.size(); ++i) it->labels.push_back(std::stof(f[i]));
  } catch (...) { return false; }
  it->path = f.back();
  while (!it->path.empty() && (it->path.back() == '\r' || it->path.back() == '\n' || it->path.back() == ' ')) it->path.pop_back();
  return !it->path.empty();
}

// one logical record in dmlc framing; payloads that contain the magic word at an aligned offset are split into continuation chunks
void WriteRecord(FILE* f, const std::string& buf) {
  auto chunk = [&](uint32_t cflag, const char* p, uint32_t len) {
    const uint32_t head[2] = {kMagic, (cflag << 29) | len};
    static const char zero[4] = {0, 0, 0, 0};
    fwrite(head, 4, 2, f);
    if (len) fwrite(p, 1, len, f);
    if (len & 3) fwrite(zero, 1, 4 - (len & 3), f);
  };
  const uint32_t n = static_cast<uint32_t>(buf.size()), aligned = n & ~3u;
  uint32_t start = 0; bool first = true;
  for (uint32_t i = 0; i < aligned; i += 4) {
    uint32_t w; memcpy(&w, buf.data() + i, 4);
    if (w != kMagic) continue;
    chunk(first ? 1u : 2u, buf.data() + start, i - start);
    start = i + 4; first = false;
  }
  chunk(first ? 0u : 3u, buf.data() + start, n - start);
}

std::string Pack(const Item& it, bool pack_label, const std::string& bytes) {
  std::string out;
  const bool multi = pack_label || it.labels.size() > 1;
  const uint32_t flag = multi ? static_cast<uint32_t>(it.labels.size()) : 0;
  const float label = multi || it.labels.empty() ? 0.f : it.labels[0];
  const uint64_t id[2] = {it.index, 0};
  out.append(reinterpret_cast<const char*>(&flag), 4); out.append(reinterpret_cast<const char*>(&label), 4); out.append(reinterpret_cast<const char*>(id), 16);
  if (multi) out.append(reinterpret_cast<const char*>(it.labels.data()), it.labels.size() * 4);
  out += bytes;
  return out;
}
}  // namespace

int main(int argc, char** argv) {
  if (argc < 4) {
    fprintf(stderr, "usage: %s image.lst image_root/ output.rec [pack_label=0] [nsplit=1] [part=0] [threads=4]\n", argv[0]);
    return 1;
  }
  std::map<std::string, long> opt = {{"pack_label", 0}, {"nsplit", 1}, {"part", 0}, {"threads", 4}};
  for (int i = 4; i < argc; ++i) {
    const char* eq = strchr(argv[i], '=');
    if (!eq || !opt.count(std::string(argv[i], eq - argv[i]))) { fprintf(stderr, "unknown option %s\n", argv[i]); return 1; }
    opt[std::string(argv[i], eq - argv[i])] = atol(eq + 1);
  }
  if (opt["nsplit"] < 1 || opt["part"] < 0 || opt["part"] >= opt["nsplit"] || opt["threads"] < 1) { fprintf(stderr, "bad nsplit / part / threads\n"); return 1; }
  std::string root = argv[2];
  if (!root.empty() && root.back() != '/') root += '/';
  std::vector<Item> items;
  {
    std::ifstream lst(argv[1]);
    if (!lst) { fprintf(stderr, "cannot open %s\n", argv[1]); return 1; }
    std::string line; Item it; size_t ln = 0;
    while (std::getline(lst, line)) { ++ln; if (line.empty()) continue; if (ParseLine(line, &it)) items.push_back(it); else fprintf(stderr, "%s:%zu: malformed line skipped\n", argv[1], ln); }
  }
  const size_t per = (items.size() + opt["nsplit"] - 1) / opt["nsplit"], lo = std::min(items.size(), per * opt["part"]), hi = std::min(items.size(), lo + per);
  std::string rec_name = argv[3];
  if (opt["nsplit"] > 1) rec_name += ".part" + std::to_string(opt["part"]);
  const size_t dot = rec_name.rfind(".rec");
  const std::string idx_name = (dot == std::string::npos ? rec_name : rec_name.substr(0, dot)) + ".idx" + (opt["nsplit"] > 1 ? ".part" + std::to_string(opt["part"]) : "");
  FILE* rec = fopen(rec_name.c_str(), "wb");
  FILE* idx = fopen(idx_name.c_str(), "w");
  if (!rec || !idx) { fprintf(stderr, "cannot open the output files\n"); return 1; }

  // readers run ahead of the writer inside a bounded window, the writer emits strictly in list order
  const size_t n = hi - lo, window = static_cast<size_t>(opt["threads"]) * 8;
  std::vector<std::string> blob(n);
  std::vector<char> state(n, 0);                  // 0 pending, 1 read, 2 failed
  std::mutex mu; std::condition_variable cv;
  std::atomic<size_t> next{0};
  size_t written = 0;
  std::vector<std::thread> readers;
  for (long t = 0; t < opt["threads"]; ++t) readers.emplace_back([&] {
    while (true) {
      const size_t i = next.fetch_add(1);
      if (i >= n) return;
      { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return i < written + window; }); }
      std::ifstream f(root + items[lo + i].path, std::ios::binary);
      std::string bytes;
      const bool ok = static_cast<bool>(f);
      if (ok) bytes.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
      std::lock_guard<std::mutex> lk(mu);
      blob[i] = std::move(bytes); state[i] = ok ? 1 : 2;
      cv.notify_all();
    }
  });
  size_t packed = 0;
  for (size_t i = 0; i < n; ++i) {
    std::string bytes; char st;
    { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return state[i] != 0; }); bytes = std::move(blob[i]); st = state[i]; written = i + 1; cv.notify_all(); }
    if (st == 2) { fprintf(stderr, "cannot read %s, skipped\n", (root + items[lo + i].path).c_str()); continue; }
    const std::string record = Pack(items[lo + i], opt["pack_label"] != 0, bytes);
    if (record.size() >= (1u << 29)) { fprintf(stderr, "%s is too large for one record, skipped\n", items[lo + i].path.c_str()); continue; }
    fprintf(idx, "%llu\t%ld\n", static_cast<unsigned long long>(items[lo + i].index), ftell(rec));
    WriteRecord(rec, record);
    if (++packed % 1000 == 0) fprintf(stderr, "%zu images packed\n", packed);
  }
  for (auto& t : readers) t.join();
  fclose(rec); fclose(idx);
  printf("packed %zu of %zu images -> %s\n", packed, n, rec_name.c_str());
  return 0;
}
