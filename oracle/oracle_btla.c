/*
 * oracle/oracle_btla.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Plain-C restatement of the BesTLA low-bit arithmetic that defines results on the weight-only
 * matmul path (BASELINE.json configs 2-5).  All functions work on the *canonical unpacked* weight
 * container  { int8 q[K][N], float scale[K/g][N], int8 zp[K/g][N] | NULL }  -- the layout
 * quantize_f32_sign_int_rowblock produces and BTLAGemmPackB consumes -- so they are independent of
 * the ISA-specific packed blob layout.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may call this.  Parity is PINNED: tests/test_oracle_vs_ref.py
 * checks each function against oracle/_ref/libref_btla.so (the reference's kernel_ref.h compiled in
 * place) on seeded inputs and against fixtures in tests/golden/.
 *
 * Reference algorithm followed (paths relative to /root/reference/bestla/bestla):
 *   bestla_utils.h:116-153   bf16 from float: round-to-nearest-even on the upper 16 bits
 *   bestla_utils.h:503-526   cast<float,int8_t> = clamp(roundf(x)), cast<float,uint8_t> = clamp(trunc(x+0.5)),
 *                            cast<float,int> = (int)roundf(x)
 *   kernel_ref.h:1608-1720   quantize_f32_sign_int_rowblock (RTN; sym "sNauto" and asym branches)
 *   kernel_ref.h:1325-1414   nf4_unpack / nf4_quantize (codes 0b0000 and 0b0111 swapped vs bitsandbytes)
 *   kernel_ref.h:1802-1823   quantize_f32_f4_rowblock (scale = absmax, code = quantize(x * (1/absmax)))
 *   kernel_ref.h:1825-1928   quantize_fp_u8_colblock / quantize_fp_s8_colblock (dynamic activation quant)
 *   kernel_ref.h:2372-2531   gemv_4bit_{u8s8,s8s8,fp32}_fp32 (the M<=4 decode numerics)
 *   kernel_ref.h:1028-1056,1113-1128  decompress_kblock_s8_fp / _s4_fp: w = (float)(q - zp) * scale
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

ORC_API uint16_t orc_f32_to_bf16(float v) {
  uint32_t u;
  memcpy(&u, &v, 4);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
ORC_API float orc_bf16_to_f32(uint16_t x) {
  uint32_t u = (uint32_t)x << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

static int8_t cast_s8(float x) {
  x = roundf(x);
  if (x > 127.f) x = 127.f;
  if (x < -128.f) x = -128.f;
  return (int8_t)x;
}
static uint8_t cast_u8(float x) {
  x += 0.5f;
  if (x > 255.f) x = 255.f;
  if (x < 0.f) x = 0.f;
  return (uint8_t)x;
}
static int cast_s32(float x) { return (int)roundf(x); }
ORC_API int8_t orc_cast_f32_s8(float x) { return cast_s8(x); }
ORC_API uint8_t orc_cast_f32_u8(float x) { return cast_u8(x); }
ORC_API int orc_cast_f32_s32(float x) { return cast_s32(x); }

static int clipi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* RTN weight quantiser.  src [K][N] (ld_src), q [K][N] (ld_dst), scales/zps [ceil(K/g)][ld_dst].
 * nbits in 1..8 (S4_CLIP = 4, S8 = 8: kernel_ref.h:1694-1711 sends both through the same branches).
 * zps == NULL -> symmetric.  A trailing partial block (K % g) is quantised on its own. */
ORC_API void orc_btla_quantize_rowblock(const float* src, int8_t* q, int K, int N, int ld_src, int ld_dst, float* scales,
                                        int8_t* zps, int g, int nbits) {
  const int full = 1 << (nbits - 1);
  const int symv = full - 1;
  for (int n = 0; n < N; ++n) {
    for (int k0 = 0; k0 < K; k0 += g) {
      const int len = (k0 + g <= K) ? g : K - k0;
      const int sidx = k0 / g * ld_dst + n;
      if (!zps) {
        float vmax = FLT_MIN, vmin = FLT_MAX, amax = 0.f;
        for (int i = 0; i < len; ++i) {
          const float v = src[(size_t)(k0 + i) * ld_src + n];
          vmax = fmaxf(vmax, v);
          vmin = fminf(vmin, v);
          amax = fmaxf(amax, fabsf(v));
        }
        float nval = (float)symv + 0.5f;
        const float sum = vmax + vmin;
        /* the reference calls abs() on a float here; <cmath> is in scope so it resolves to the float overload */
        if (fabsf(sum) >= amax / (float)full) nval = sum > 0.f ? (float)-full : (float)full;
        const float scale = amax / nval;
        const float rscale = 1.f / scale;
        scales[sidx] = scale;
        for (int i = 0; i < len; ++i)
          q[(size_t)(k0 + i) * ld_dst + n] = (int8_t)clipi(cast_s8(src[(size_t)(k0 + i) * ld_src + n] * rscale), -full, symv);
      } else {
        float vmax = 0.f, vmin = 0.f;
        for (int i = 0; i < len; ++i) {
          const float v = src[(size_t)(k0 + i) * ld_src + n];
          vmax = fmaxf(vmax, v);
          vmin = fminf(vmin, v);
        }
        const float scale = (vmax - vmin) / (float)((1 << nbits) - 1);
        const float rscale = 1.f / scale;
        scales[sidx] = scale;
        const int zp = clipi(cast_s32((0.f - vmin) * rscale) - full, -full, symv);
        zps[sidx] = (int8_t)zp;
        for (int i = 0; i < len; ++i)
          q[(size_t)(k0 + i) * ld_dst + n] =
              (int8_t)clipi(cast_s32(src[(size_t)(k0 + i) * ld_src + n] * rscale) + zp, -full, symv);
      }
    }
  }
}

/* NF4 levels indexed by the reference's code (code 0 <-> 0.0, code 7 <-> -1.0: swapped vs bitsandbytes) */
static const float NF4_LUT[16] = {0.f,
                                  -0.6961928009986877f,
                                  -0.5250730514526367f,
                                  -0.39491748809814453f,
                                  -0.28444138169288635f,
                                  -0.18477343022823334f,
                                  -0.09105003625154495f,
                                  -1.f,
                                  0.07958029955625534f,
                                  0.16093020141124725f,
                                  0.24611230194568634f,
                                  0.33791524171829224f,
                                  0.44070982933044434f,
                                  0.5626170039176941f,
                                  0.7229568362236023f,
                                  1.0f};
ORC_API float orc_nf4_unpack(int code) { return NF4_LUT[code & 15]; }

/* decision thresholds = midpoints between adjacent levels, strict '>' as in nf4_quantize */
ORC_API int orc_nf4_quantize(float x) {
  static const float thr[15] = {-0.8480964004993439f, -0.6106329262256622f,  -0.4599952697753906f, -0.33967943489551544f,
                                -0.23460740596055984f, -0.13791173323988914f, -0.045525018125772476f, 0.03979014977812767f,
                                0.1202552504837513f,  0.2035212516784668f,   0.2920137718319893f,  0.3893125355243683f,
                                0.5016634166240692f,  0.6427869200706482f,   0.8614784181118011f};
  /* codes in increasing level order */
  static const int order[16] = {7, 1, 2, 3, 4, 5, 6, 0, 8, 9, 10, 11, 12, 13, 14, 15};
  int r = 0;
  while (r < 15 && x > thr[r]) ++r;
  return order[r];
}

ORC_API void orc_btla_quantize_nf4_rowblock(const float* src, int8_t* q, int K, int N, int ld_src, int ld_dst,
                                            float* scales, int g) {
  for (int n = 0; n < N; ++n)
    for (int k0 = 0; k0 < K; k0 += g) {
      const int len = (k0 + g <= K) ? g : K - k0;
      float amax = FLT_MIN;
      for (int i = 0; i < len; ++i) amax = fmaxf(amax, fabsf(src[(size_t)(k0 + i) * ld_src + n]));
      scales[k0 / g * ld_dst + n] = amax;
      const float r = 1.f / amax;
      for (int i = 0; i < len; ++i)
        q[(size_t)(k0 + i) * ld_dst + n] = (int8_t)orc_nf4_quantize(src[(size_t)(k0 + i) * ld_src + n] * r);
    }
}

/* dequantise the canonical container to fp32 [K][N]: int -> (q - zp) * scale; nf4 -> lut[q] * scale */
ORC_API void orc_btla_dequant(const int8_t* q, const float* scales, const int8_t* zps, float* w, int K, int N, int g,
                              int is_nf4) {
  for (int k = 0; k < K; ++k)
    for (int n = 0; n < N; ++n) {
      const int s = k / g * N + n;
      if (is_nf4)
        w[(size_t)k * N + n] = NF4_LUT[q[(size_t)k * N + n] & 15] * scales[s];
      else
        w[(size_t)k * N + n] = (float)(q[(size_t)k * N + n] - (zps ? zps[s] : 0)) * scales[s];
    }
}

/* dynamic activation quant, u8 asymmetric per (row, K-block). blkreduce may be NULL. */
ORC_API void orc_btla_quantize_act_u8(int M, int K, const float* src, int ld_src, uint8_t* dst, int ld_dst, float* scales,
                                      int ld_scale, uint8_t* zps, int g, float* blkreduce) {
  for (int m = 0; m < M; ++m)
    for (int k0 = 0; k0 < K; k0 += g) {
      const int full_blk = (k0 + g <= K);
      const int len = full_blk ? g : K - k0;
      float vmax = full_blk ? FLT_MIN : 0.f, vmin = 0.f; /* kernel_ref.h:1833 vs :1859 */
      for (int i = 0; i < len; ++i) {
        const float v = src[(size_t)m * ld_src + k0 + i];
        vmax = fmaxf(v, vmax);
        vmin = fminf(v, vmin);
      }
      const float scale = (vmax - vmin) / 255;
      const uint8_t zp = cast_u8((0 - vmin) / scale);
      const float rscale = 1.f / scale;
      scales[(size_t)m * ld_scale + k0 / g] = scale;
      zps[(size_t)m * ld_scale + k0 / g] = zp;
      int sum = 0;
      for (int i = 0; i < len; ++i) {
        const int qt = cast_s32(src[(size_t)m * ld_src + k0 + i] * rscale);
        sum += qt;
        dst[(size_t)m * ld_dst + k0 + i] = cast_u8((float)zp + (float)qt);
      }
      if (blkreduce) blkreduce[(size_t)m * ld_scale + k0 / g] = (float)sum * scale;
    }
}

/* dynamic activation quant, s8 symmetric per (row, K-block) */
ORC_API void orc_btla_quantize_act_s8(int M, int K, const float* src, int ld_src, int8_t* dst, int ld_dst, float* scales,
                                      int ld_scale, int g, float* reduce) {
  for (int m = 0; m < M; ++m)
    for (int k0 = 0; k0 < K; k0 += g) {
      const int len = (k0 + g <= K) ? g : K - k0;
      float amax = FLT_MIN;
      for (int i = 0; i < len; ++i) amax = fmaxf(fabsf(src[(size_t)m * ld_src + k0 + i]), amax);
      const float scale = amax / 127;
      const float rscale = 1.f / scale;
      scales[(size_t)m * ld_scale + k0 / g] = scale;
      int sum = 0;
      for (int i = 0; i < len; ++i) {
        const int8_t t = cast_s8(src[(size_t)m * ld_src + k0 + i] * rscale);
        dst[(size_t)m * ld_dst + k0 + i] = t;
        sum += t;
      }
      if (reduce) reduce[(size_t)m * ld_scale + k0 / g] = (float)sum * scale;
    }
}

/* ---- the three decode numerics (kernel_ref.h gemv_4bit_*), on the canonical container ----
 * q holds signed values (nibble - 8), as decompress_s4_s8 yields; zp as stored (signed). */

/* comp fp32: acc += a * (float)(q - zp) * scale, k ascending (kernel_ref.h:2490-2531) */
ORC_API void orc_btla_gemv_fp32(const float* A, int lda, const int8_t* q, const float* scales, const int8_t* zps, float* C,
                                int ldc, int M, int N, int K, int g) {
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float acc = 0.f;
      for (int k = 0; k < K; ++k) {
        const int s = k / g * N + n;
        const int v = q[(size_t)k * N + n] - (zps ? zps[s] : 0);
        acc += A[(size_t)m * lda + k] * (float)v * scales[s];
      }
      C[(size_t)m * ldc + n] = acc;
    }
}

/* comp int8, u8 activations: acc += (int)(a8 - azp) * (q - zp) * (ascale * bscale)  (kernel_ref.h:2372-2430).
 * The reference adds the 4 products of a k-quad one at a time in fp32; (a-azp)*(q-zp) is an exact int. */
ORC_API void orc_btla_gemv_u8s8(const uint8_t* a8, const float* as, const uint8_t* az, int lda, int ldas, const int8_t* q,
                                const float* scales, const int8_t* zps, float* C, int ldc, int M, int N, int K, int g) {
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float acc = 0.f;
      for (int k = 0; k < K; ++k) {
        const int b = k / g;
        const float vscale = as[(size_t)m * ldas + b] * scales[b * N + n];
        const int wv = q[(size_t)k * N + n] - (zps ? zps[b * N + n] : 0);
        const int av = (int)a8[(size_t)m * lda + k] - (int)az[(size_t)m * ldas + b];
        acc += (float)(av * wv) * vscale;
      }
      C[(size_t)m * ldc + n] = acc;
    }
}

/* comp int8, s8 activations (kernel_ref.h:2432-2488) */
ORC_API void orc_btla_gemv_s8s8(const int8_t* a8, const float* as, int lda, int ldas, const int8_t* q, const float* scales,
                                const int8_t* zps, float* C, int ldc, int M, int N, int K, int g) {
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float acc = 0.f;
      for (int k = 0; k < K; ++k) {
        const int b = k / g;
        const float vscale = as[(size_t)m * ldas + b] * scales[b * N + n];
        const int wv = q[(size_t)k * N + n] - (zps ? zps[b * N + n] : 0);
        acc += (float)((int)a8[(size_t)m * lda + k] * wv) * vscale;
      }
      C[(size_t)m * ldc + n] = acc;
    }
}

/* block-exact form of the int8 numerics: C = sum_b (ascale*bscale) * (float)isum_b, with
 * isum_b = sum_{k in b} (a - azp)(q - zp) as one exact integer per block.  This is what a dp4a kernel
 * computes; it differs from orc_btla_gemv_u8s8 only in fp32 summation granularity. */
ORC_API void orc_btla_gemv_u8s8_blocksum(const uint8_t* a8, const float* as, const uint8_t* az, int lda, int ldas,
                                         const int8_t* q, const float* scales, const int8_t* zps, float* C, int ldc, int M,
                                         int N, int K, int g) {
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float acc = 0.f;
      for (int k0 = 0; k0 < K; k0 += g) {
        const int b = k0 / g;
        const int len = (k0 + g <= K) ? g : K - k0;
        int isum = 0;
        for (int i = 0; i < len; ++i)
          isum += ((int)a8[(size_t)m * lda + k0 + i] - (int)az[(size_t)m * ldas + b]) *
                  (q[(size_t)(k0 + i) * N + n] - (zps ? zps[b * N + n] : 0));
        acc += (float)isum * (as[(size_t)m * ldas + b] * scales[b * N + n]);
      }
      C[(size_t)m * ldc + n] = acc;
    }
}

/* fp32 GEMM on already-dequantised weights W[K][N] (the UT_CompFp32 criterion, ut/bestla_prologue_b.cpp:471-511),
 * accumulated in double so it is an order-independent ground truth. */
ORC_API void orc_gemm_f64acc(const float* A, int lda, const float* W, float* C, int ldc, int M, int N, int K) {
  double* acc = (double*)malloc(sizeof(double) * (size_t)N);
  for (int m = 0; m < M; ++m) {
    for (int n = 0; n < N; ++n) acc[n] = 0.0;
    for (int k = 0; k < K; ++k) {
      const double a = A[(size_t)m * lda + k];
      const float* wr = W + (size_t)k * N;
      for (int n = 0; n < N; ++n) acc[n] += a * (double)wr[n];
    }
    for (int n = 0; n < N; ++n) C[(size_t)m * ldc + n] = (float)acc[n];
  }
  free(acc);
}

/* epilogue activations used by the fused FFN (kernel_ref.h:1574 swish alpha=-1: x / (1 + exp(-x))) */
ORC_API float orc_silu(float x) { return x / (1.f + expf(-x)); }
