// RecordIO and data-iterator groups of the flat C ABI.
//
// Parity: include/mxnet/c_api.h
//   :2180-2260  MXRecordIOWriterCreate / Free / WriteRecord / Tell,  MXRecordIOReaderCreate / Free / ReadRecord / Seek / Tell
//               (format of 3rdparty/dmlc-core/include/dmlc/recordio.h: magic 0xced7230a, lrecord = cflag << 29 | length, payload padded to 4 bytes,
//                payloads that contain the magic word are split into continuation chunks)
//   :1760-1860  MXListDataIters / MXDataIterGetIterInfo / CreateIter / Free / Next / BeforeFirst / GetData / GetLabel / GetIndex / GetPadNum
//               over src/io/iter_mnist.cc (MNISTIter) and src/io/iter_csv.cc (CSVIter); batches come back as host NDArray handles
//               (host_array.h) that stay valid until the next call to Next / the iterator is freed.
// ImageRecordIter's JPEG pipeline lives in the Python front end (image.py over the native record reader, recordio.h); a C front end reads the
// records through the reader below and decodes with its own codec.
#include <array>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <numeric>
#include <random>
#include <string>
#include <vector>

#include "host_array.h"
#include "io.h"
#include "text_io.h"

#define GX_CAPI extern "C" __attribute__((visibility("default")))

void GXRTSetLastError(const std::string& msg);

namespace {
using gxrt::capi::HostArray;

template <typename F>
int Guard(F&& f) {
  try { f(); return 0; }
  catch (const std::exception& e) { GXRTSetLastError(e.what()); return -1; }
  catch (...) { GXRTSetLastError("unknown error"); return -1; }
}

constexpr uint32_t kMagic = 0xced7230a;
constexpr uint32_t kLenMask = (1u << 29) - 1;

struct RecWriter {
  FILE* f = nullptr;
  ~RecWriter() { if (f) fclose(f); }
  void Chunk(uint32_t cflag, const char* p, uint32_t len) {
    const uint32_t head[2] = {kMagic, (cflag << 29) | len};
    static const char zero[4] = {0, 0, 0, 0};
    if (fwrite(head, 4, 2, f) != 2 || (len && fwrite(p, 1, len, f) != len) || (((4 - (len & 3)) & 3) && fwrite(zero, 1, (4 - (len & 3)) & 3, f) != ((4 - (len & 3)) & 3)))
      throw std::runtime_error("RecordIO: write failed");
  }
  void Write(const char* buf, size_t size) {
    if (size >= (1u << 29)) throw std::runtime_error("RecordIO: a record must be smaller than 2^29 bytes");
    // the magic word may not appear at a 4-byte aligned position inside a chunk: split there (the reader re-inserts it)
    const uint32_t n = static_cast<uint32_t>(size), aligned = n & ~3u;
    uint32_t start = 0; bool first = true;
    for (uint32_t i = 0; i < aligned; i += 4) {
      uint32_t w; memcpy(&w, buf + i, 4);
      if (w != kMagic) continue;
      Chunk(first ? 1u : 2u, buf + start, i - start);
      start = i + 4; first = false;
    }
    Chunk(first ? 0u : 3u, buf + start, n - start);
  }
};

struct RecReader {
  FILE* f = nullptr;
  std::string buf;
  ~RecReader() { if (f) fclose(f); }
  // false at end of file
  bool Read() {
    buf.clear();
    bool more = true, first = true;
    while (more) {
      uint32_t head[2];
      const size_t got = fread(head, 4, 2, f);
      if (got == 0 && first) return false;
      if (got != 2) throw std::runtime_error("RecordIO: truncated record header");
      if (head[0] != kMagic) throw std::runtime_error("RecordIO: bad magic word");
      const uint32_t cflag = head[1] >> 29, len = head[1] & kLenMask, padded = (len + 3) & ~3u;
      if (first ? (cflag != 0 && cflag != 1) : (cflag != 2 && cflag != 3)) throw std::runtime_error("RecordIO: chunk flags out of sequence");
      if (!first) buf.append(reinterpret_cast<const char*>(&kMagic), 4);
      const size_t at = buf.size();
      buf.resize(at + padded);
      if (padded && fread(&buf[at], 1, padded, f) != padded) throw std::runtime_error("RecordIO: truncated record payload");
      buf.resize(at + len);
      more = cflag == 1 || cflag == 2;
      first = false;
    }
    return true;
  }
};

// ---------------------------------------------------------------------------------------------- data iterators
struct IterInfo { const char* name; const char* doc; std::vector<std::array<const char*, 3>> params; };
const std::vector<IterInfo>& Iters() {
  static const std::vector<IterInfo> v = {
      {"MNISTIter", "batches of the MNIST idx files (src/io/iter_mnist.cc)",
       {{"image", "string, required", "idx3 image file"}, {"label", "string, required", "idx1 label file"}, {"batch_size", "int, optional, default=128", "batch size"},
        {"shuffle", "boolean, optional, default=1", "shuffle every epoch"}, {"flat", "boolean, optional, default=0", "(batch, 784) instead of (batch, 1, 28, 28)"},
        {"seed", "int, optional, default=0", "shuffle seed"}, {"num_parts", "int, optional, default=1", "number of partitions"},
        {"part_index", "int, optional, default=0", "which partition to read"}}},
      {"CSVIter", "batches of dense rows from CSV files (src/io/iter_csv.cc)",
       {{"data_csv", "string, required", "data file"}, {"data_shape", "Shape(tuple), required", "shape of one example"},
        {"label_csv", "string, optional", "label file (zeros when absent)"}, {"label_shape", "Shape(tuple), optional, default=(1,)", "shape of one label"},
        {"batch_size", "int, required", "batch size"}, {"round_batch", "boolean, optional, default=1", "wrap around to fill the last batch"}}}};
  return v;
}

struct DataIter {
  std::vector<float> data, label;          // the whole set, row-major
  std::vector<int64_t> dshape, lshape;     // per example
  int64_t n = 0, batch = 1, cursor = 0;    // cursor: next example
  bool shuffle = false, round_batch = true;   // round_batch: a short last batch wraps around to the first examples (pad reports how many)
  bool drop_last = false;                     // MNISTIter: a short last batch is not produced (iter_mnist.cc:96)
  std::vector<int64_t> order;
  std::mt19937 rng;
  int pad = 0;
  std::unique_ptr<HostArray> bdata, blabel;
  std::vector<uint64_t> index;

  void Alloc() {
    auto mk = [&](const std::vector<int64_t>& per) {
      auto a = std::make_unique<HostArray>();
      a->rec.dtype = 0; a->rec.shape.push_back(batch);
      int64_t numel = batch;
      for (auto d : per) { a->rec.shape.push_back(d); numel *= d; }
      a->rec.data.assign(static_cast<size_t>(numel) * 4, '\0');
      return a;
    };
    bdata = mk(dshape); blabel = mk(lshape);
    if (lshape.size() == 1 && lshape[0] == 1) blabel->rec.shape.pop_back();       // labels of width 1 are (batch,)
    order.resize(static_cast<size_t>(n)); std::iota(order.begin(), order.end(), 0);
    Reset();
  }
  void Reset() { cursor = 0; if (shuffle) std::shuffle(order.begin(), order.end(), rng); }
  bool Next() {
    if (cursor >= n || (drop_last && cursor + batch > n)) return false;
    const int64_t dper = std::accumulate(dshape.begin(), dshape.end(), int64_t{1}, std::multiplies<int64_t>());
    const int64_t lper = std::accumulate(lshape.begin(), lshape.end(), int64_t{1}, std::multiplies<int64_t>());
    const int64_t have = std::min(batch, n - cursor);
    pad = static_cast<int>(batch - have);
    index.clear();
    float* bd = reinterpret_cast<float*>(&bdata->rec.data[0]); float* bl = reinterpret_cast<float*>(&blabel->rec.data[0]);
    for (int64_t b = 0; b < batch; ++b) {
      const int64_t src = order[static_cast<size_t>((cursor + b) % n)];
      memcpy(bd + b * dper, data.data() + src * dper, dper * 4);
      memcpy(bl + b * lper, label.data() + src * lper, lper * 4);
      index.push_back(static_cast<uint64_t>(src));
    }
    cursor += batch;
    return true;
  }
};

DataIter* IT(void* h) { if (!h) throw std::runtime_error("null DataIter handle"); return static_cast<DataIter*>(h); }

std::vector<int64_t> ParseTuple(const std::string& s) {
  std::vector<int64_t> out; size_t i = 0;
  while (i < s.size()) {
    if (std::isdigit(static_cast<unsigned char>(s[i]))) { size_t j = i; while (j < s.size() && std::isdigit(static_cast<unsigned char>(s[j]))) ++j; out.push_back(std::stoll(s.substr(i, j - i))); i = j; }
    else ++i;
  }
  return out;
}
bool Truthy(const std::string& v) { return v == "1" || v == "True" || v == "true"; }

DataIter* MakeMNIST(const std::map<std::string, std::string>& kw) {
  auto get = [&](const char* k, const char* def) { auto it = kw.find(k); return it == kw.end() ? std::string(def) : it->second; };
  if (!kw.count("image") || !kw.count("label")) throw std::runtime_error("MNISTIter: image and label are required");
  auto img = gxrt::ReadIdx(get("image", "")); auto lab = gxrt::ReadIdx(get("label", ""));
  if (img.first.size() != 3 || lab.first.size() != 1 || img.first[0] != lab.first[0]) throw std::runtime_error("MNISTIter: image / label files do not match");
  auto it = std::make_unique<DataIter>();
  const int64_t total = img.first[0], H = img.first[1], W = img.first[2];
  if (static_cast<int64_t>(img.second.size()) < total * H * W || static_cast<int64_t>(lab.second.size()) < total) throw std::runtime_error("MNISTIter: idx payload shorter than its header says");
  const int64_t parts = std::max<int64_t>(std::stoll(get("num_parts", "1")), 1), part = std::stoll(get("part_index", "0"));
  if (part < 0 || part >= parts) throw std::runtime_error("MNISTIter: part_index out of range");
  const int64_t per = total / parts, lo = part * per, hi = part + 1 == parts ? total : lo + per;
  it->n = hi - lo;
  it->batch = std::stoll(get("batch_size", "128"));
  if (it->batch < 1 || it->n < 1) throw std::runtime_error("MNISTIter: empty partition or non-positive batch size");
  it->shuffle = Truthy(get("shuffle", "1"));
  it->rng.seed(static_cast<uint32_t>(std::stoll(get("seed", "0"))));
  it->dshape = Truthy(get("flat", "0")) ? std::vector<int64_t>{H * W} : std::vector<int64_t>{1, H, W};
  it->lshape = {1};
  it->data.resize(static_cast<size_t>(it->n * H * W)); it->label.resize(static_cast<size_t>(it->n));
  const unsigned char* px = reinterpret_cast<const unsigned char*>(img.second.data()) + lo * H * W;
  for (size_t i = 0; i < it->data.size(); ++i) it->data[i] = px[i] * (1.f / 256.f);                  // iter_mnist.cc:117 scales by 1/256
  const unsigned char* lb = reinterpret_cast<const unsigned char*>(lab.second.data()) + lo;
  for (size_t i = 0; i < it->label.size(); ++i) it->label[i] = lb[i];
  it->drop_last = true;
  it->Alloc();
  return it.release();
}

DataIter* MakeCSV(const std::map<std::string, std::string>& kw) {
  auto get = [&](const char* k, const char* def) { auto it = kw.find(k); return it == kw.end() ? std::string(def) : it->second; };
  if (!kw.count("data_csv") || !kw.count("data_shape") || !kw.count("batch_size")) throw std::runtime_error("CSVIter: data_csv, data_shape and batch_size are required");
  auto it = std::make_unique<DataIter>();
  it->dshape = ParseTuple(get("data_shape", ""));
  it->lshape = ParseTuple(get("label_shape", "(1,)"));
  if (it->dshape.empty() || it->lshape.empty()) throw std::runtime_error("CSVIter: empty data_shape / label_shape");
  const int64_t dper = std::accumulate(it->dshape.begin(), it->dshape.end(), int64_t{1}, std::multiplies<int64_t>());
  const int64_t lper = std::accumulate(it->lshape.begin(), it->lshape.end(), int64_t{1}, std::multiplies<int64_t>());
  gx_rt::CSVData d = gx_rt::ParseCSV(get("data_csv", ""));
  if (d.cols != dper) throw std::runtime_error("CSVIter: rows have " + std::to_string(d.cols) + " columns, data_shape needs " + std::to_string(dper));
  it->n = d.rows; it->data = std::move(d.values);
  if (kw.count("label_csv")) {
    gx_rt::CSVData l = gx_rt::ParseCSV(get("label_csv", ""));
    if (l.rows != d.rows || l.cols != lper) throw std::runtime_error("CSVIter: label file does not match the data file / label_shape");
    it->label = std::move(l.values);
  } else it->label.assign(static_cast<size_t>(it->n * lper), 0.f);
  it->batch = std::stoll(get("batch_size", "1"));
  if (it->batch < 1 || it->n < 1) throw std::runtime_error("CSVIter: empty file or non-positive batch size");
  it->round_batch = Truthy(get("round_batch", "1"));
  it->Alloc();
  return it.release();
}
}  // namespace

// ================================================================================================ RecordIO
GX_CAPI int GXRecordIOWriterCreate(const char* uri, void** out) {
  return Guard([&] { auto w = std::make_unique<RecWriter>(); w->f = fopen(uri, "wb"); if (!w->f) throw std::runtime_error(std::string("cannot open ") + uri); *out = w.release(); });
}
GX_CAPI int GXRecordIOWriterFree(void* h) { return Guard([&] { delete static_cast<RecWriter*>(h); }); }
GX_CAPI int GXRecordIOWriterWriteRecord(void* h, const char* buf, size_t size) {
  return Guard([&] { if (!h) throw std::runtime_error("null writer"); static_cast<RecWriter*>(h)->Write(buf, size); });
}
GX_CAPI int GXRecordIOWriterTell(void* h, size_t* pos) {
  return Guard([&] { if (!h) throw std::runtime_error("null writer"); const long p = ftell(static_cast<RecWriter*>(h)->f); if (p < 0) throw std::runtime_error("RecordIO: tell failed"); *pos = static_cast<size_t>(p); });
}
GX_CAPI int GXRecordIOReaderCreate(const char* uri, void** out) {
  return Guard([&] { auto r = std::make_unique<RecReader>(); r->f = fopen(uri, "rb"); if (!r->f) throw std::runtime_error(std::string("cannot open ") + uri); *out = r.release(); });
}
GX_CAPI int GXRecordIOReaderFree(void* h) { return Guard([&] { delete static_cast<RecReader*>(h); }); }
// *buf == nullptr and *size == 0 at end of file; the buffer belongs to the reader and is valid until the next read
GX_CAPI int GXRecordIOReaderReadRecord(void* h, const char** buf, size_t* size) {
  return Guard([&] {
    if (!h) throw std::runtime_error("null reader");
    RecReader* r = static_cast<RecReader*>(h);
    if (r->Read()) { *buf = r->buf.data(); *size = r->buf.size(); } else { *buf = nullptr; *size = 0; }
  });
}
GX_CAPI int GXRecordIOReaderSeek(void* h, size_t pos) {
  return Guard([&] { if (!h) throw std::runtime_error("null reader"); if (fseek(static_cast<RecReader*>(h)->f, static_cast<long>(pos), SEEK_SET) != 0) throw std::runtime_error("RecordIO: seek failed"); });
}
GX_CAPI int GXRecordIOReaderTell(void* h, size_t* pos) {
  return Guard([&] { if (!h) throw std::runtime_error("null reader"); const long p = ftell(static_cast<RecReader*>(h)->f); if (p < 0) throw std::runtime_error("RecordIO: tell failed"); *pos = static_cast<size_t>(p); });
}

// ================================================================================================ data iterators
GX_CAPI int GXListDataIters(uint32_t* out_size, void*** out_array) {
  return Guard([&] {
    static thread_local std::vector<void*> v;
    v.clear(); for (auto& i : Iters()) v.push_back(const_cast<IterInfo*>(&i));
    *out_size = static_cast<uint32_t>(v.size()); *out_array = v.data();
  });
}
GX_CAPI int GXDataIterGetIterInfo(void* creator, const char** name, const char** description, uint32_t* num_args, const char*** arg_names,
                                  const char*** arg_type_infos, const char*** arg_descriptions) {
  return Guard([&] {
    if (!creator) throw std::runtime_error("null creator");
    const IterInfo* i = static_cast<IterInfo*>(creator);
    static thread_local std::vector<const char*> n, t, d;
    n.clear(); t.clear(); d.clear();
    for (auto& p : i->params) { n.push_back(p[0]); t.push_back(p[1]); d.push_back(p[2]); }
    *name = i->name; *description = i->doc; *num_args = static_cast<uint32_t>(n.size()); *arg_names = n.data(); *arg_type_infos = t.data(); *arg_descriptions = d.data();
  });
}
GX_CAPI int GXDataIterCreateIter(void* creator, uint32_t num_param, const char** keys, const char** vals, void** out) {
  return Guard([&] {
    if (!creator) throw std::runtime_error("null creator");
    std::map<std::string, std::string> kw;
    for (uint32_t i = 0; i < num_param; ++i) kw[keys[i]] = vals[i];
    const std::string name = static_cast<IterInfo*>(creator)->name;
    *out = name == "MNISTIter" ? MakeMNIST(kw) : MakeCSV(kw);
  });
}
GX_CAPI int GXDataIterFree(void* h) { return Guard([&] { delete IT(h); }); }
GX_CAPI int GXDataIterBeforeFirst(void* h) { return Guard([&] { IT(h)->Reset(); }); }
GX_CAPI int GXDataIterNext(void* h, int* out) { return Guard([&] { *out = IT(h)->Next() ? 1 : 0; }); }
GX_CAPI int GXDataIterGetData(void* h, void** out) { return Guard([&] { *out = IT(h)->bdata.get(); }); }
GX_CAPI int GXDataIterGetLabel(void* h, void** out) { return Guard([&] { *out = IT(h)->blabel.get(); }); }
GX_CAPI int GXDataIterGetIndex(void* h, uint64_t** out_index, uint64_t* out_size) {
  return Guard([&] { DataIter* i = IT(h); *out_index = i->index.data(); *out_size = i->index.size(); });
}
GX_CAPI int GXDataIterGetPadNum(void* h, int* pad) { return Guard([&] { *pad = IT(h)->pad; }); }
