// Native text-format readers behind mx.io.CSVIter / mx.io.LibSVMIter.
// Parity: src/io/iter_csv.cc (dense float rows, fixed column count) and src/io/iter_libsvm.cc (label idx:val ... -> CSR), both of which the
// reference parses with dmlc-core's threaded text parsers.  One pass over an mmap-free buffered read, strtof-based, no per-token allocation.
#pragma once
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace gx_rt {

inline std::string ReadWholeFile(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f.good()) throw std::runtime_error("cannot open " + path);
  return std::string((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

struct CSVData { std::vector<float> values; long rows = 0, cols = 0; };

inline CSVData ParseCSV(const std::string& path) {
  const std::string buf = ReadWholeFile(path);
  CSVData out;
  const char* p = buf.c_str();
  const char* end = p + buf.size();
  long cols_this = 0;
  while (p < end) {
    while (p < end && (*p == ' ' || *p == '\t' || *p == '\r')) ++p;
    if (p >= end) break;
    if (*p == '\n') {
      if (cols_this) { if (!out.cols) out.cols = cols_this; else if (cols_this != out.cols) throw std::runtime_error("ragged CSV row in " + path); ++out.rows; cols_this = 0; }
      ++p; continue;
    }
    char* next = nullptr;
    const float v = strtof(p, &next);
    if (next == p) throw std::runtime_error("bad number in " + path);
    out.values.push_back(v); ++cols_this;
    p = next;
    while (p < end && (*p == ' ' || *p == '\t' || *p == '\r')) ++p;
    if (p < end && *p == ',') ++p;
  }
  if (cols_this) { if (!out.cols) out.cols = cols_this; else if (cols_this != out.cols) throw std::runtime_error("ragged CSV row in " + path); ++out.rows; }
  return out;
}

struct LibSVMData { std::vector<float> labels, values; std::vector<long> indices, indptr; long max_index = -1; };

inline LibSVMData ParseLibSVM(const std::string& path) {
  const std::string buf = ReadWholeFile(path);
  LibSVMData out;
  out.indptr.push_back(0);
  const char* p = buf.c_str();
  const char* end = p + buf.size();
  while (p < end) {
    while (p < end && (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n')) ++p;
    if (p >= end) break;
    if (*p == '#') { while (p < end && *p != '\n') ++p; continue; }
    char* next = nullptr;
    out.labels.push_back(strtof(p, &next));
    if (next == p) throw std::runtime_error("bad label in " + path);
    p = next;
    while (p < end && *p != '\n') {
      while (p < end && (*p == ' ' || *p == '\t' || *p == '\r')) ++p;
      if (p >= end || *p == '\n' || *p == '#') break;
      const long idx = strtol(p, &next, 10);
      if (next == p || *next != ':') throw std::runtime_error("bad idx:val pair in " + path);
      p = next + 1;
      const float v = strtof(p, &next);
      if (next == p) throw std::runtime_error("bad value in " + path);
      p = next;
      out.indices.push_back(idx); out.values.push_back(v);
      if (idx > out.max_index) out.max_index = idx;
    }
    while (p < end && *p != '\n') ++p;
    out.indptr.push_back(static_cast<long>(out.indices.size()));
  }
  return out;
}

}  // namespace gx_rt
