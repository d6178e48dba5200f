// CPU gradient compression used by the host-side servers/workers: 2-bit with residual, Bi-Sparse (BSC).
// Parity: src/kvstore/gradient_compression.{h,cc} (SetParams :46-58, Encode/DecodeParams :82-100, Quantize/Dequantize 2bit :118-189 with
// the bit layout of gradient_compression-inl.h:40-127, BSCompress :191-269, BSCPullCompress :271-308, BSCDecompress :310-336).
// Contract notes: see geomx_b200/kvstore/compression.py (sampling is a deterministic stride instead of std::shuffle(seed 42)).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <queue>
#include <string>
#include <vector>

#include "base.h"

namespace hips {

enum class CompressionType { kNone = 0, kTwoBit = 1, kBiSparse = 2 };

class GradientCompression {
 public:
  static constexpr float kPadVal = -65530.f;
  static constexpr float kPadIdx = -1.f;
  CompressionType type() const { return type_; }
  float threshold() const { return threshold_; }
  void SetParams(const std::string& type, float threshold) {
    if (type == "none") type_ = CompressionType::kNone;
    else if (type == "2bit") type_ = CompressionType::kTwoBit;
    else if (type == "bsc") type_ = CompressionType::kBiSparse;
    else throw Error("Unknown type for gradient compression " + type);
    if (type_ != CompressionType::kNone) HIPS_CHECK_MSG(threshold > 0, "threshold must be greater than 0");
    threshold_ = threshold;
  }
  std::string EncodeParams() const { return std::to_string(static_cast<int>(type_)) + "," + std::to_string(threshold_); }
  void DecodeParams(const std::string& s) {
    const size_t c = s.find(',');
    HIPS_CHECK(c != std::string::npos);
    type_ = static_cast<CompressionType>(std::stoi(s.substr(0, c)));
    threshold_ = std::stof(s.substr(c + 1));
  }
  static int64_t CompressedSize2Bit(int64_t n) { return (n + 15) / 16; }

  // out: ceil(n/16) 32-bit words
  void Quantize2Bit(const float* grad, float* residual, uint32_t* out, int64_t n) const {
    const float thr = threshold_;
    const int64_t words = CompressedSize2Bit(n);
    for (int64_t w = 0; w < words; ++w) {
      uint32_t word = 0;
      for (int j = 0; j < 16; ++j) {
        const int64_t i = w * 16 + j;
        if (i >= n) break;
        float r = residual[i] + grad[i];
        uint32_t code = 0;
        if (r >= thr) { code = 3; r -= thr; }
        else if (r <= -thr) { code = 2; r += thr; }
        residual[i] = r;
        word |= code << (((j >> 2) << 3) + (6 - 2 * (j & 3)));
      }
      out[w] = word;
    }
  }
  void Dequantize2Bit(const uint32_t* in, float* out, int64_t n, bool accumulate = false) const {
    const float thr = threshold_;
    for (int64_t i = 0; i < n; ++i) {
      const int j = static_cast<int>(i & 15);
      const uint32_t code = (in[i >> 4] >> (((j >> 2) << 3) + (6 - 2 * (j & 3)))) & 3u;
      const float v = code == 3 ? thr : (code == 2 ? -thr : 0.f);
      out[i] = accumulate ? out[i] + v : v;
    }
  }

  static void BSCSizes(int64_t n, float thr, int* k, int* sample, int* k_sample) {
    *k = static_cast<int>(static_cast<float>(n) * thr);
    int s = (n * 0.005 * thr >= 10) ? static_cast<int>(n * 0.005) : static_cast<int>(10 / thr);
    s = std::max<int64_t>(1, std::min<int64_t>(s, n));
    *sample = s;
    *k_sample = std::max(1, static_cast<int>(s * thr));
  }
  // out: 2k floats [vals | idx]; u, v: error-feedback state (momentum 0.9)
  void BSCompress(const float* grad, float* u, float* v, float* out, int64_t n) const {
    int k, sample, k_sample;
    BSCSizes(n, threshold_, &k, &sample, &k_sample);
    for (int64_t i = 0; i < n; ++i) { u[i] = u[i] * 0.9f + grad[i]; v[i] += u[i]; }
    if (k == 0) return;
    const int64_t stride = std::max<int64_t>(1, n / sample);
    std::priority_queue<float, std::vector<float>, std::greater<float>> q;
    for (int j = 0; j < sample; ++j) {
      const float a = std::fabs(v[j * stride]);
      if (static_cast<int>(q.size()) < k_sample || a > q.top()) {
        if (static_cast<int>(q.size()) == k_sample) q.pop();
        q.push(a);
      }
    }
    const float boundary = q.top();
    int cnt = 0;
    for (int64_t i = 0; i < n && cnt < k; ++i) {
      if (std::fabs(v[i]) >= boundary) {
        out[cnt] = v[i]; out[k + cnt] = static_cast<float>(i);
        v[i] = 0; u[i] = 0; ++cnt;
      }
    }
    for (; cnt < k; ++cnt) { out[cnt] = kPadVal; out[k + cnt] = kPadIdx; }
  }
  // keep non-zeros in index order, capacity k = n * thr * multiplier
  void BSCPullCompress(const float* dense, float* out, int64_t n, int multiplier) const {
    const int k = static_cast<int>(static_cast<float>(n) * threshold_ * multiplier);
    int cnt = 0;
    for (int64_t i = 0; i < n && cnt < k; ++i)
      if (dense[i] != 0) { out[cnt] = dense[i]; out[k + cnt] = static_cast<float>(i); ++cnt; }
    for (; cnt < k; ++cnt) { out[cnt] = kPadVal; out[k + cnt] = kPadIdx; }
  }
  static int64_t BSCPullSize(int64_t n, float thr, int multiplier) { return 2 * static_cast<int64_t>(static_cast<float>(n) * thr * multiplier); }
  static void BSCDecompress(const float* zipped, int64_t zipped_len, float* out, int64_t n, bool accumulate = false) {
    const int64_t k = zipped_len / 2;
    if (!accumulate) memset(out, 0, n * sizeof(float));
    for (int64_t j = 0; j < k; ++j) {
      const int64_t idx = static_cast<int64_t>(zipped[k + j]);
      if (idx >= 0 && idx < n) { if (accumulate) out[idx] += zipped[j]; else out[idx] = zipped[j]; }
    }
  }

 private:
  CompressionType type_ = CompressionType::kNone;
  float threshold_ = 0.5f;
};

}  // namespace hips
