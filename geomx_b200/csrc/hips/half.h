// Software fp16 / bf16 <-> fp32 conversion for host-side servers (parity: 3rdparty/ps-lite/src/half_float/umHalf.{h,inl}, used by
// MergeMsg_HALF van.cc:310-328).  Round-to-nearest-even, handles subnormals / inf / nan.
#pragma once
#include <cstdint>
#include <cstring>

namespace hips {

inline float HalfToFloat(uint16_t h) {
  const uint32_t sign = (h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1Fu, man = h & 0x3FFu, bits;
  if (exp == 0) {
    if (man == 0) bits = sign;
    else {
      exp = 127 - 15 + 1;
      while ((man & 0x400u) == 0) { man <<= 1; --exp; }
      man &= 0x3FFu;
      bits = sign | (exp << 23) | (man << 13);
    }
  } else if (exp == 31) bits = sign | 0x7F800000u | (man << 13);
  else bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
  float f; memcpy(&f, &bits, 4);
  return f;
}
inline uint16_t FloatToHalf(float f) {
  uint32_t x; memcpy(&x, &f, 4);
  const uint32_t sign = (x >> 16) & 0x8000u;
  int32_t exp = static_cast<int32_t>((x >> 23) & 0xFF) - 127 + 15;
  uint32_t man = x & 0x7FFFFFu;
  if (((x >> 23) & 0xFF) == 0xFF) return static_cast<uint16_t>(sign | 0x7C00u | (man ? 0x200u : 0));
  if (exp >= 31) return static_cast<uint16_t>(sign | 0x7C00u);
  if (exp <= 0) {
    if (exp < -10) return static_cast<uint16_t>(sign);
    man |= 0x800000u;
    const int shift = 14 - exp;
    uint32_t half_man = man >> shift;
    const uint32_t rem = man & ((1u << shift) - 1), halfway = 1u << (shift - 1);
    if (rem > halfway || (rem == halfway && (half_man & 1))) ++half_man;
    return static_cast<uint16_t>(sign | half_man);
  }
  uint32_t half = sign | (static_cast<uint32_t>(exp) << 10) | (man >> 13);
  const uint32_t rem = man & 0x1FFFu;
  if (rem > 0x1000u || (rem == 0x1000u && (half & 1))) ++half;
  return static_cast<uint16_t>(half);
}
inline float BF16ToFloat(uint16_t h) { uint32_t b = static_cast<uint32_t>(h) << 16; float f; memcpy(&f, &b, 4); return f; }
inline uint16_t FloatToBF16(float f) {
  uint32_t x; memcpy(&x, &f, 4);
  if ((x & 0x7FFFFFFFu) > 0x7F800000u) return static_cast<uint16_t>((x >> 16) | 0x40u);
  x += 0x7FFFu + ((x >> 16) & 1u);
  return static_cast<uint16_t>(x >> 16);
}

}  // namespace hips
