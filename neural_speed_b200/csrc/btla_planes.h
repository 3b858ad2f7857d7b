// btla_planes.h -- the bit-plane layout BesTLA uses for weight codes that are not 4 or 8 bits wide (host code only).
//
// StorageWeightKBlockNInteger::resize (bestla/bestla/bestla_storage.h:724-745) sizes mQBuf as the SUM of power-of-two planes and
// compressBit{2,3,5,6,7}Weight (bestla_prologue_b.h:512-564) + compress_{2,3,5,6,7}bit (kernel_ref.h:178-345) fill them over the
// FLAT element index e of the tile-interleaved [NPad x KPad] buffer (the same order the 4-bit nibbles use):
//     code u = q + 2^(bits-1)                      (q = the signed integer the RTN quantiser produced)
//     bits 7:  4-bit plane @0 | 2-bit plane @E/2 | 1-bit plane @E/2+E/4      u = b4 | b2 << 4 | b1 << 6
//     bits 6:  4-bit plane @0 | 2-bit plane @E/2                             u = b4 | b2 << 4
//     bits 5:  4-bit plane @0 | 1-bit plane @E/2                             u = b4 | b1 << 4
//     bits 3:  2-bit plane @0 | 1-bit plane @E/4                             u = b2 | b1 << 2
//     bits 2:  2-bit plane @0
// inside a plane element e sits at bit (e % (8/w)) * w of byte e / (8/w) (utils::bit4x2 / bit2x4 / bit1x8 bit-fields, x86 order).
#pragma once
#include <cstddef>
#include <cstdint>

namespace ns_planes {

struct Layout {
  int nplanes;
  int width[3];
  size_t off[3];  // byte offset of each plane
  size_t bytes;   // total
};

inline bool layout(int bits, size_t E, Layout* L) {
  static const int W[8][3] = {{0, 0, 0}, {1, 0, 0}, {2, 0, 0}, {2, 1, 0}, {4, 0, 0}, {4, 1, 0}, {4, 2, 0}, {4, 2, 1}};
  if (bits < 1 || bits > 7) return false;
  size_t at = 0;
  L->nplanes = 0;
  for (int i = 0; i < 3 && W[bits][i]; ++i) {
    L->width[i] = W[bits][i];
    L->off[i] = at;
    at += (E * (size_t)W[bits][i] + 7) / 8;  // utils::updiv(KPad * NPad * w, 8)
    L->nplanes = i + 1;
  }
  L->bytes = at;
  return true;
}

inline int get(const uint8_t* q, const Layout& L, size_t e) {
  int u = 0, sh = 0;
  for (int i = 0; i < L.nplanes; ++i) {
    const int w = L.width[i], per = 8 / w;
    u |= ((q[L.off[i] + e / per] >> ((e % per) * w)) & ((1 << w) - 1)) << sh;
    sh += w;
  }
  return u;
}

inline void put(uint8_t* q, const Layout& L, size_t e, int u) {
  for (int i = 0; i < L.nplanes; ++i) {
    const int w = L.width[i], per = 8 / w, mask = (1 << w) - 1;
    uint8_t& b = q[L.off[i] + e / per];
    b = (uint8_t)((b & ~(mask << ((e % per) * w))) | ((u & mask) << ((e % per) * w)));
    u >>= w;
  }
}

}  // namespace ns_planes
