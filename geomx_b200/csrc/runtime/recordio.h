// Native RecordIO reader: memory-maps a .rec file, finds the logical records and hands out (re-assembled) payloads without going through
// Python file objects.  Format (dmlc-core include/dmlc/recordio.h): every chunk is  uint32 magic 0xced7230a | uint32 (cflag << 29 | length) |
// payload | pad to 4 bytes;  a payload that contains the magic word is stored as several chunks (cflag 1 first, 2 middle, 3 last) and the
// magic words between them are re-inserted on reading.  Capability parity: dmlc::RecordIOReader / RecordIOChunkReader used by the reference's
// ImageRecordIter (src/io/iter_image_recordio_2.cc): sequential scan, random access by offset, multi-threaded batch reads.
#pragma once
#include <fcntl.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace gxrt {
namespace py = pybind11;

class RecordFile {
 public:
  static constexpr uint32_t kMagic = 0xced7230a;
  explicit RecordFile(const std::string& path) : path_(path) {
    fd_ = ::open(path.c_str(), O_RDONLY);
    if (fd_ < 0) throw std::runtime_error("cannot open " + path);
    struct stat st;
    if (::fstat(fd_, &st) != 0) { ::close(fd_); throw std::runtime_error("cannot stat " + path); }
    size_ = static_cast<size_t>(st.st_size);
    if (size_ > 0) {
      void* p = ::mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
      if (p == MAP_FAILED) { ::close(fd_); throw std::runtime_error("cannot map " + path); }
      base_ = static_cast<const uint8_t*>(p);
      ::madvise(p, size_, MADV_WILLNEED);
    }
  }
  ~RecordFile() {
    if (base_ != nullptr) ::munmap(const_cast<uint8_t*>(base_), size_);
    if (fd_ >= 0) ::close(fd_);
  }
  RecordFile(const RecordFile&) = delete;
  RecordFile& operator=(const RecordFile&) = delete;

  size_t size() const { return size_; }

  // byte offsets of all logical records (the first chunk of each), in file order
  std::vector<uint64_t> Scan() const {
    std::vector<uint64_t> out;
    size_t pos = 0;
    while (pos + 8 <= size_) {
      uint32_t magic, lrec;
      memcpy(&magic, base_ + pos, 4); memcpy(&lrec, base_ + pos + 4, 4);
      if (magic != kMagic) throw std::runtime_error("invalid RecordIO file " + path_ + " (bad magic at " + std::to_string(pos) + ")");
      const uint32_t cflag = lrec >> 29, len = lrec & ((1u << 29) - 1);
      if (cflag == 0 || cflag == 1) out.push_back(pos);
      pos += 8 + ((static_cast<size_t>(len) + 3) & ~size_t(3));
    }
    return out;
  }

  // payload of the logical record that starts at `off`
  std::string Read(uint64_t off) const {
    std::string out;
    size_t pos = static_cast<size_t>(off);
    bool first = true;
    while (true) {
      if (pos + 8 > size_) throw std::runtime_error("RecordIO: offset " + std::to_string(off) + " runs past the end of " + path_);
      uint32_t magic, lrec;
      memcpy(&magic, base_ + pos, 4); memcpy(&lrec, base_ + pos + 4, 4);
      if (magic != kMagic) throw std::runtime_error("RecordIO: no record starts at offset " + std::to_string(pos) + " of " + path_);
      const uint32_t cflag = lrec >> 29, len = lrec & ((1u << 29) - 1);
      if (pos + 8 + len > size_) throw std::runtime_error("RecordIO: truncated record at offset " + std::to_string(pos));
      if (first && cflag != 0 && cflag != 1) throw std::runtime_error("RecordIO: offset " + std::to_string(off) + " is inside a multi-chunk record");
      if (!first) { const uint32_t m = kMagic; out.append(reinterpret_cast<const char*>(&m), 4); }
      out.append(reinterpret_cast<const char*>(base_ + pos + 8), len);
      if (cflag == 0 || cflag == 3) return out;
      first = false;
      pos += 8 + ((static_cast<size_t>(len) + 3) & ~size_t(3));
    }
  }

  // several records at once, copied by `threads` workers (the GIL is released by the binding)
  std::vector<std::string> ReadMany(const std::vector<uint64_t>& offs, int threads) const {
    std::vector<std::string> out(offs.size());
    const int nt = std::max(1, std::min<int>(threads, static_cast<int>(offs.size())));
    std::vector<std::thread> pool;
    std::vector<std::string> errors(nt);
    for (int t = 0; t < nt; ++t) {
      pool.emplace_back([&, t] {
        try { for (size_t i = t; i < offs.size(); i += nt) out[i] = Read(offs[i]); }
        catch (const std::exception& e) { errors[t] = e.what(); }
      });
    }
    for (auto& th : pool) th.join();
    for (auto& e : errors) if (!e.empty()) throw std::runtime_error(e);
    return out;
  }

 private:
  std::string path_;
  int fd_ = -1;
  const uint8_t* base_ = nullptr;
  size_t size_ = 0;
};

inline void BindRecordIO(py::module_& m) {
  py::class_<RecordFile>(m, "RecordFile")
      .def(py::init<const std::string&>())
      .def_property_readonly("size", &RecordFile::size)
      .def("scan", &RecordFile::Scan, py::call_guard<py::gil_scoped_release>())
      .def("read", [](const RecordFile& f, uint64_t off) { std::string s; { py::gil_scoped_release g; s = f.Read(off); } return py::bytes(s); })
      .def("read_many", [](const RecordFile& f, const std::vector<uint64_t>& offs, int threads) {
        std::vector<std::string> v;
        { py::gil_scoped_release g; v = f.ReadMany(offs, threads); }
        py::list out;
        for (auto& s : v) out.append(py::bytes(s));
        return out;
      }, py::arg("offsets"), py::arg("threads") = 4);
}

}  // namespace gxrt
