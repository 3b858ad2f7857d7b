/*
 * oracle/ref_btla.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Compiles the reference's own scalar BesTLA kernels straight from /root/reference
 * (bestla/bestla/kernel_ref.h + bestla.h + bestla_utils.h are #included where they lie; nothing is
 * copied) and exposes them with a C ABI.  These three headers are the only part of BesTLA that builds
 * without the un-vendored xbyak v7.06 dependency (bestla/CMakeLists.txt:29-35); they hold every piece
 * of low-bit arithmetic that defines results on the hot path (RTN quantiser, NF4, s4 pack/unpack,
 * activation quantisers, reference GEMVs).  Used to pin oracle/oracle_btla.cpp and as the
 * "reference-scalar" CPU timing for BesTLA configs.
 */
#include <cstdint>
#include <cstring>
#include <cmath>
#include "kernel_ref.h"

using namespace bestla;
#define REF_API extern "C" __attribute__((visibility("default")))

/* kernel_ref.h:1608 quantize_f32_sign_int_rowblock. src is [row=K][col=N] (ld_src), dst int8 [K][N] (ld_dst),
 * scales/zps [K/blocksize][ld_dst]. qtype is the raw BTLA_DTYPE value (bestla.h:38-87). */
REF_API int ref_btla_quantize_f32_sign_int_rowblock(const float* src, int8_t* dst, int row, int col, int ld_src,
                                                    int ld_dst, float* scales, int8_t* zps, int blocksize,
                                                    uint32_t qtype) {
  return (int)kernel::ref::quantize_f32_sign_int_rowblock(src, dst, row, col, ld_src, ld_dst, scales, zps, blocksize,
                                                          (BTLA_DTYPE)qtype);
}

/* kernel_ref.h:1802 quantize_f32_f4_rowblock<F4_NF4> */
REF_API int ref_btla_quantize_f32_nf4_rowblock(const float* src, int8_t* dst, int row, int col, int ld_src, int ld_dst,
                                               float* scales, int blocksize) {
  return (int)kernel::ref::quantize_f32_f4_rowblock<BTLA_DTYPE::F4_NF4>(src, dst, row, col, ld_src, ld_dst, scales,
                                                                       blocksize);
}
REF_API float ref_btla_nf4_unpack(int8_t v) { return kernel::ref::nf4_unpack(v); }
REF_API int8_t ref_btla_nf4_quantize(float x) { return kernel::ref::nf4_quantize(x); }

/* kernel_ref.h:1825 / :1886 activation quantisers */
REF_API int ref_btla_quantize_fp_u8_colblock(int row, int col, const float* src, int ld_src, uint8_t* dst, int ld_dst,
                                             float* scales, int ld_scale, uint8_t* zps, int blocksize, float* blkreduce) {
  return (int)kernel::ref::quantize_fp_u8_colblock<float>(row, col, src, ld_src, dst, ld_dst, scales, ld_scale, zps,
                                                          blocksize, blkreduce);
}
REF_API int ref_btla_quantize_fp_s8_colblock(int row, int col, const float* src, int ld_src, int8_t* dst, int ld_dst,
                                             float* scales, int ld_scale, int blocksize, float* reduce) {
  return (int)kernel::ref::quantize_fp_s8_colblock<float>(row, col, src, ld_src, dst, ld_dst, scales, ld_scale,
                                                          blocksize, reduce);
}

/* kernel_ref.h:40 padding_interleave (int8), :155 compress_s8_s4, :471 decompress_s4_s8 */
REF_API int ref_btla_padding_interleave_s8(const int8_t* src, int8_t* dst, int row, int col, int rowpad, int colpad,
                                           int src_step, int dst_step, int NTile, int RowPack) {
  return (int)kernel::ref::padding_interleave<int8_t, int8_t>(src, dst, row, col, rowpad, colpad, src_step, dst_step,
                                                              NTile, RowPack);
}
REF_API int ref_btla_compress_s8_s4(const int8_t* src, uint8_t* dst, size_t size) {
  return (int)kernel::ref::compress_s8_s4(src, reinterpret_cast<utils::int4x2*>(dst), size);
}
REF_API int ref_btla_decompress_s4_s8(const uint8_t* src, int8_t* dst, size_t elt) {
  return (int)kernel::ref::decompress_s4_s8(reinterpret_cast<utils::int4x2*>(const_cast<uint8_t*>(src)), dst, elt,
                                            nullptr, 0);
}

/* kernel_ref.h:1802 quantize_f32_f4_rowblock<F4_T> and :1416 f4_unpack<F4_T> for the three 4-bit float codebooks.
 * kind: 0 = F4_NF4, 1 = F4_BNB, 2 = F4_E2M1 */
REF_API int ref_btla_quantize_f32_f4_rowblock(int kind, const float* src, int8_t* dst, int row, int col, int ld_src, int ld_dst,
                                              float* scales, int blocksize) {
  if (kind == 1)
    return (int)kernel::ref::quantize_f32_f4_rowblock<BTLA_DTYPE::F4_BNB>(src, dst, row, col, ld_src, ld_dst, scales, blocksize);
  if (kind == 2)
    return (int)kernel::ref::quantize_f32_f4_rowblock<BTLA_DTYPE::F4_E2M1>(src, dst, row, col, ld_src, ld_dst, scales, blocksize);
  return (int)kernel::ref::quantize_f32_f4_rowblock<BTLA_DTYPE::F4_NF4>(src, dst, row, col, ld_src, ld_dst, scales, blocksize);
}
REF_API float ref_btla_f4_unpack(int kind, int8_t v) {
  if (kind == 1) return kernel::ref::f4_unpack<BTLA_DTYPE::F4_BNB>(v);
  if (kind == 2) return kernel::ref::f4_unpack<BTLA_DTYPE::F4_E2M1>(v);
  return kernel::ref::f4_unpack<BTLA_DTYPE::F4_NF4>(v);
}

/* kernel_ref.h:178-345 compress_{7,6,5,3,2}bit with the plane pointers placed as compressBit{7,6,5,3,2}Weight do
 * (bestla_prologue_b.h:512-564: bit1_offset / bit2_offset = N * K elements).  dst must hold the sum of the planes. */
REF_API int ref_btla_compress_bits(int bits, const int8_t* src, uint8_t* dst, size_t elt) {
  int8_t* d = reinterpret_cast<int8_t*>(dst);
  switch (bits) {
    case 7: {
      auto b4 = reinterpret_cast<utils::bit4x2*>(d);
      auto b2 = reinterpret_cast<utils::bit2x4*>(b4 + elt / 2);
      auto b1 = reinterpret_cast<utils::bit1x8*>(b2 + elt / 4);
      return (int)kernel::ref::compress_7bit(src, b4, b2, b1, elt);
    }
    case 6:
      return (int)kernel::ref::compress_6bit(src, reinterpret_cast<utils::bit4x2*>(d), reinterpret_cast<utils::bit2x4*>(d + elt / 2), elt);
    case 5:
      return (int)kernel::ref::compress_5bit(src, reinterpret_cast<utils::bit4x2*>(d), reinterpret_cast<utils::bit1x8*>(d + elt / 2), elt);
    case 3:
      return (int)kernel::ref::compress_3bit(src, reinterpret_cast<utils::bit2x4*>(d), reinterpret_cast<utils::bit1x8*>(d + elt / 4), elt);
    case 2:
      return (int)kernel::ref::compress_2bit(src, reinterpret_cast<utils::bit2x4*>(d), elt);
    default:
      return -1;
  }
}

/* bestla_utils.h:116-153 bf16 RNE, :503-526 cast<> rounding */
REF_API uint16_t ref_btla_f32_to_bf16(float v) { return utils::bf16(v).x; }
REF_API float ref_btla_bf16_to_f32(uint16_t x) { return utils::bf16::from_bin(x).tofloat(); }
REF_API int8_t ref_btla_cast_f32_s8(float v) { return utils::cast<float, int8_t>(v); }
REF_API uint8_t ref_btla_cast_f32_u8(float v) { return utils::cast<float, uint8_t>(v); }
REF_API int ref_btla_cast_f32_s32(float v) { return utils::cast<float, int>(v); }

/* kernel_ref.h:2372/:2432/:2490 reference 4-bit GEMVs, NTILE=48 (AVX512/AMX cores, bestla_defs.h:36-54) and 24.
 * b4: packed nibbles in the GEMV layout of the respective core: fp32 -> [K][NTILE] (PackRow 1),
 * u8s8/s8s8 -> [K/4][NTILE][4] (PackRow 4).  scales/zp: [K/blocksize][ldzp]. */
template <int NTILE, int MTILE>
static int gemv_fp32(const float* A, int lda, const uint8_t* b4, const float* bs, const int8_t* bz, int ldzp, float* C,
                     int ldc, int k, int blocksize) {
  utils::GemvParamB<float> B;
  B.b4ptr = const_cast<uint8_t*>(b4);
  B.sptr = const_cast<float*>(bs);
  B.zpptr = const_cast<int8_t*>(bz);
  B.nbits = 4;
  B.ldzp = ldzp;
  B.kpad = k;
  return (int)kernel::ref::gemv_4bit_fp32_fp32<float, NTILE, MTILE>(A, lda, B, C, ldc, k, blocksize, nullptr, 0);
}
template <int NTILE, int MTILE>
static int gemv_u8s8(const uint8_t* a8, const float* as, const uint8_t* az, int lda, int ldazp, const uint8_t* b4,
                     const float* bs, const int8_t* bz, int ldzp, float* C, int ldc, int k, int blocksize) {
  utils::GemvParamA A{const_cast<uint8_t*>(a8), const_cast<float*>(as), const_cast<uint8_t*>(az), lda, ldazp};
  utils::GemvParamB<float> B;
  B.b4ptr = const_cast<uint8_t*>(b4);
  B.sptr = const_cast<float*>(bs);
  B.zpptr = const_cast<int8_t*>(bz);
  B.nbits = 4;
  B.ldzp = ldzp;
  B.kpad = k;
  return (int)kernel::ref::gemv_4bit_u8s8_fp32<float, NTILE, MTILE>(A, B, C, ldc, k, blocksize, nullptr, 0);
}
template <int NTILE, int MTILE>
static int gemv_s8s8(const int8_t* a8, const float* as, int lda, int ldazp, const uint8_t* b4, const float* bs,
                     const int8_t* bz, int ldzp, float* C, int ldc, int k, int blocksize) {
  utils::GemvParamA A{reinterpret_cast<uint8_t*>(const_cast<int8_t*>(a8)), const_cast<float*>(as), nullptr, lda, ldazp};
  utils::GemvParamB<float> B;
  B.b4ptr = const_cast<uint8_t*>(b4);
  B.sptr = const_cast<float*>(bs);
  B.zpptr = const_cast<int8_t*>(bz);
  B.nbits = 4;
  B.ldzp = ldzp;
  B.kpad = k;
  return (int)kernel::ref::gemv_4bit_s8s8_fp32<float, NTILE, MTILE>(A, B, C, ldc, k, blocksize, nullptr, 0);
}

#define DISPATCH_M(fn, ...)                  \
  switch (mtile) {                           \
    case 1: return fn<48, 1>(__VA_ARGS__);   \
    case 2: return fn<48, 2>(__VA_ARGS__);   \
    case 3: return fn<48, 3>(__VA_ARGS__);   \
    case 4: return fn<48, 4>(__VA_ARGS__);   \
    default: return -1;                      \
  }

REF_API int ref_btla_gemv_4bit_fp32_fp32_n48(int mtile, const float* A, int lda, const uint8_t* b4, const float* bs,
                                             const int8_t* bz, int ldzp, float* C, int ldc, int k, int blocksize) {
  DISPATCH_M(gemv_fp32, A, lda, b4, bs, bz, ldzp, C, ldc, k, blocksize)
}
REF_API int ref_btla_gemv_4bit_u8s8_fp32_n48(int mtile, const uint8_t* a8, const float* as, const uint8_t* az, int lda,
                                             int ldazp, const uint8_t* b4, const float* bs, const int8_t* bz, int ldzp,
                                             float* C, int ldc, int k, int blocksize) {
  DISPATCH_M(gemv_u8s8, a8, as, az, lda, ldazp, b4, bs, bz, ldzp, C, ldc, k, blocksize)
}
REF_API int ref_btla_gemv_4bit_s8s8_fp32_n48(int mtile, const int8_t* a8, const float* as, int lda, int ldazp,
                                             const uint8_t* b4, const float* bs, const int8_t* bz, int ldzp, float* C,
                                             int ldc, int k, int blocksize) {
  DISPATCH_M(gemv_s8s8, a8, as, lda, ldazp, b4, bs, bz, ldzp, C, ldc, k, blocksize)
}

/* kernel_ref.h:1113 decompress_kblock_s4_fp<PackRow,48,float>: one NTILE-wide strip [row][48] */
REF_API int ref_btla_decompress_kblock_s4_f32_n48(int packrow, const uint8_t* src, float* dst, int row, const void* scales,
                                                  uint32_t sdtype, const int8_t* zps, int k_offset, int n_offset,
                                                  int blocksize, int ldzp) {
  int8_t tmp[48 * 4];
  auto s = reinterpret_cast<utils::int4x2*>(const_cast<uint8_t*>(src));
  auto sc = const_cast<void*>(scales);
  auto zp = const_cast<int8_t*>(zps);
  switch (packrow) {
    case 1:
      return (int)kernel::ref::decompress_kblock_s4_fp<1, 48, float>(s, dst, row, 48, sc, (BTLA_DTYPE)sdtype, zp,
                                                                     k_offset, n_offset, blocksize, ldzp, tmp, sizeof(tmp));
    case 2:
      return (int)kernel::ref::decompress_kblock_s4_fp<2, 48, float>(s, dst, row, 48, sc, (BTLA_DTYPE)sdtype, zp,
                                                                     k_offset, n_offset, blocksize, ldzp, tmp, sizeof(tmp));
    case 4:
      return (int)kernel::ref::decompress_kblock_s4_fp<4, 48, float>(s, dst, row, 48, sc, (BTLA_DTYPE)sdtype, zp,
                                                                     k_offset, n_offset, blocksize, ldzp, tmp, sizeof(tmp));
  }
  return -1;
}
