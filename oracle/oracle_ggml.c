/*
 * oracle/oracle_ggml.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Plain-C restatement (no intrinsics, no reference headers) of the reference's ggml-style
 * Q4_0 x Q8_0 weight-only matmul: the path BASELINE.json config 0 / the headline metric runs.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * call this.  Parity is PINNED: tests/test_oracle_vs_ref.py checks every function here bit-for-bit
 * against oracle/_ref/libref_ggml.so (the reference's own headers compiled in place) and against
 * the committed fixtures in tests/golden/ generated from that library.
 *
 * Reference algorithm followed (paths relative to /root/reference/neural_speed):
 *   core/data_types.h:79-111        block_q4_0 {fp16 d; u8 qs[16]}, block_q8_0 {fp16 d; i8 qs[32]}
 *   core/data_types.h:148-230       fp16 <-> fp32 (F16C = IEEE round-to-nearest-even)
 *   vectors/cpu/quantize.h:243-279  quantize_row_q4_0_reference
 *   vectors/cpu/quantize.h:686-704  dequantize_row_q4_0
 *   vectors/cpu/quantize.h:447-560  quantize_row_q8_0 (the AVX/AVX2 body is what runs on x86:
 *                                   id = 127/amax, round-half-even) and :422-445 the *_reference
 *                                   variant (id = 1/d, roundf = half-away) -- they differ, both kept.
 *   core/layers/vec_dot.h:131-164   ne_vec_dot_q4_0_q8_0, AVX2 body: 8 fp32 lanes, lane l accumulates
 *                                   fma(dw*da, (float)sum_{j<4} w[4l+j]*a[4l+j], acc[l]); hsum order
 *                                   from quantize.h:46-52.  :318-333 scalar body (different rounding order).
 *   core/ne_layers.c:7085-7203      ne_compute_forward_mul_mat_q_f32 (INIT quantises src1 rows, COMPUTE
 *                                   = vec_dot per (src1 row, src0 row)).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))
#define QK 32

#pragma pack(push, 1)
typedef struct { uint16_t d; uint8_t qs[QK / 2]; } orc_q4_0;   /* 18 bytes */
typedef struct { uint16_t d; int8_t qs[QK]; } orc_q8_0;        /* 34 bytes */
#pragma pack(pop)

/* ---- IEEE binary16 <-> binary32, round-to-nearest-even (what _cvtss_sh(x,0)/_cvtsh_ss do) ---- */
ORC_API float orc_fp16_to_fp32(uint16_t h) {
  uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1f;
  uint32_t man = h & 0x3ffu;
  uint32_t bits;
  if (exp == 0) {
    if (man == 0) {
      bits = sign;
    } else { /* subnormal: normalise */
      int e = -1;
      do { man <<= 1; ++e; } while (!(man & 0x400u));
      bits = sign | (uint32_t)(127 - 15 - e) << 23 | (man & 0x3ffu) << 13;
    }
  } else if (exp == 31) {
    bits = sign | 0x7f800000u | man << 13;
  } else {
    bits = sign | (exp + 112) << 23 | man << 13;
  }
  float f;
  memcpy(&f, &bits, 4);
  return f;
}

ORC_API uint16_t orc_fp32_to_fp16(float f) {
  uint32_t x;
  memcpy(&x, &f, 4);
  uint32_t sign = (x >> 16) & 0x8000u;
  uint32_t ax = x & 0x7fffffffu;
  if (ax >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | (ax > 0x7f800000u ? 0x200u | ((ax >> 13) & 0x3ffu) : 0));
  if (ax >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u); /* >= 65520 rounds to inf */
  if (ax < 0x33000001u) return (uint16_t)sign;               /* < 2^-25 (or == 2^-25 tie->even 0) */
  int e = (int)(ax >> 23) - 127;
  uint32_t m = (ax & 0x7fffffu) | 0x800000u;
  int shift;
  uint32_t base;
  if (e < -14) { shift = 13 + (-14 - e); base = 0; }       /* result is subnormal half */
  else { shift = 13; base = (uint32_t)(e + 15) << 10; m &= 0x7fffffu; }
  uint32_t q = m >> shift;
  uint32_t rem = m & ((1u << shift) - 1);
  uint32_t half = 1u << (shift - 1);
  if (rem > half || (rem == half && (q & 1))) ++q;
  return (uint16_t)(sign | (base + q)); /* carry into exponent is the correct rounding */
}

/* ---- Q4_0 weights ---- */
ORC_API void orc_quantize_row_q4_0(const float* x, void* vy, int k) {
  orc_q4_0* y = (orc_q4_0*)vy;
  for (int b = 0; b < k / QK; ++b) {
    const float* xb = x + b * QK;
    float amax = 0.f, vmax = 0.f; /* the signed value with the largest magnitude (first wins ties) */
    for (int j = 0; j < QK; ++j)
      if (fabsf(xb[j]) > amax) { amax = fabsf(xb[j]); vmax = xb[j]; }
    const float d = vmax / -8.f;
    const float inv = d != 0.f ? 1.0f / d : 0.0f;
    y[b].d = orc_fp32_to_fp16(d);
    for (int j = 0; j < QK / 2; ++j) {
      /* x*id + 8.5f: gcc -O3 -mfma (the reference's default x86 flags, -ffp-contract=fast) contracts this
       * into one fma; oracle/_ref is built with those flags, so the fused form is the pinned one. */
      int lo = (int8_t)fmaf(xb[j], inv, 8.5f);        /* truncation, as the reference's (int8_t) cast */
      int hi = (int8_t)fmaf(xb[j + QK / 2], inv, 8.5f);
      if (lo > 15) lo = 15;
      if (hi > 15) hi = 15;
      y[b].qs[j] = (uint8_t)(lo | hi << 4);          /* element j -> low nibble, j+16 -> high nibble */
    }
  }
}

ORC_API void orc_dequantize_row_q4_0(const void* vx, float* y, int k) {
  const orc_q4_0* x = (const orc_q4_0*)vx;
  for (int b = 0; b < k / QK; ++b) {
    const float d = orc_fp16_to_fp32(x[b].d);
    for (int j = 0; j < QK / 2; ++j) {
      y[b * QK + j] = (float)((x[b].qs[j] & 0xf) - 8) * d;
      y[b * QK + j + QK / 2] = (float)((x[b].qs[j] >> 4) - 8) * d;
    }
  }
}

/* ---- Q8_0 activations ---- */
/* x86 runtime body: d = amax/127 (stored fp16), q = rint(x * (127/amax)) */
ORC_API void orc_quantize_row_q8_0(const float* x, void* vy, int k) {
  orc_q8_0* y = (orc_q8_0*)vy;
  for (int b = 0; b < k / QK; ++b) {
    const float* xb = x + b * QK;
    float amax = 0.f;
    for (int j = 0; j < QK; ++j) amax = fmaxf(amax, fabsf(xb[j]));
    y[b].d = orc_fp32_to_fp16(amax / 127.f);
    const float inv = amax != 0.f ? 127.f / amax : 0.f;
    for (int j = 0; j < QK; ++j) y[b].qs[j] = (int8_t)(int)nearbyintf(xb[j] * inv); /* FE_TONEAREST = half-even */
  }
}

/* *_reference variant: id = 1/d, roundf (half away from zero) */
ORC_API void orc_quantize_row_q8_0_reference(const float* x, void* vy, int k) {
  orc_q8_0* y = (orc_q8_0*)vy;
  for (int b = 0; b < k / QK; ++b) {
    const float* xb = x + b * QK;
    float amax = 0.f;
    for (int j = 0; j < QK; ++j) amax = fmaxf(amax, fabsf(xb[j]));
    const float d = amax / 127.f;
    const float inv = d != 0.f ? 1.0f / d : 0.f;
    y[b].d = orc_fp32_to_fp16(d);
    for (int j = 0; j < QK; ++j) y[b].qs[j] = (int8_t)roundf(xb[j] * inv);
  }
}

ORC_API void orc_dequantize_row_q8_0(const void* vx, float* y, int k) {
  const orc_q8_0* x = (const orc_q8_0*)vx;
  for (int b = 0; b < k / QK; ++b) {
    const float d = orc_fp16_to_fp32(x[b].d);
    for (int j = 0; j < QK; ++j) y[b * QK + j] = (float)x[b].qs[j] * d;
  }
}

/* hsum_float_8 order (quantize.h:46-52): (a0+a4)+(a2+a6) then + ((a1+a5)+(a3+a7)) */
static float lanes8_hsum(const float* a) {
  const float r0 = a[4] + a[0], r1 = a[5] + a[1], r2 = a[6] + a[2], r3 = a[7] + a[3];
  const float s0 = r0 + r2, s1 = r1 + r3;
  return s0 + s1;
}

/* AVX2-structured dot: bit-exact with the reference built with -mavx2 -mfma.
 * bytes_from_nibbles_32 puts elements 0..15 (low nibbles) in the low 128 bits and 16..31 in the high
 * ones, i.e. natural element order; each fp32 lane owns 4 consecutive elements. */
ORC_API void orc_vec_dot_q4_0_q8_0(int n, float* s, const void* vx, const void* vy) {
  const orc_q4_0* x = (const orc_q4_0*)vx;
  const orc_q8_0* y = (const orc_q8_0*)vy;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int b = 0; b < n / QK; ++b) {
    const float d = orc_fp16_to_fp32(x[b].d) * orc_fp16_to_fp32(y[b].d);
    int w[QK];
    for (int j = 0; j < QK / 2; ++j) {
      w[j] = (x[b].qs[j] & 0xf) - 8;
      w[j + QK / 2] = (x[b].qs[j] >> 4) - 8;
    }
    for (int l = 0; l < 8; ++l) {
      int si = 0;
      for (int j = 0; j < 4; ++j) si += w[4 * l + j] * y[b].qs[4 * l + j];
      acc[l] = fmaf(d, (float)si, acc[l]);
    }
  }
  *s = lanes8_hsum(acc);
}

/* scalar body of the reference (vec_dot.h:318-333): sumf += sumi*dw*da, left to right */
ORC_API void orc_vec_dot_q4_0_q8_0_scalar(int n, float* s, const void* vx, const void* vy) {
  const orc_q4_0* x = (const orc_q4_0*)vx;
  const orc_q8_0* y = (const orc_q8_0*)vy;
  float sumf = 0.f;
  for (int b = 0; b < n / QK; ++b) {
    int si = 0;
    for (int j = 0; j < QK / 2; ++j)
      si += ((x[b].qs[j] & 0xf) - 8) * y[b].qs[j] + ((x[b].qs[j] >> 4) - 8) * y[b].qs[j + QK / 2];
    sumf += (float)si * orc_fp16_to_fp32(x[b].d) * orc_fp16_to_fp32(y[b].d);
  }
  *s = sumf;
}

/* integer block sums only (exact; what any correct kernel must reproduce bit-for-bit) */
ORC_API void orc_block_isum_q4_0_q8_0(int n, int32_t* out, const void* vx, const void* vy) {
  const orc_q4_0* x = (const orc_q4_0*)vx;
  const orc_q8_0* y = (const orc_q8_0*)vy;
  for (int b = 0; b < n / QK; ++b) {
    int si = 0;
    for (int j = 0; j < QK / 2; ++j)
      si += ((x[b].qs[j] & 0xf) - 8) * y[b].qs[j] + ((x[b].qs[j] >> 4) - 8) * y[b].qs[j + QK / 2];
    out[b] = si;
  }
}

/* dst[m][n] = vec_dot(W row n, q8(A row m)); W: [N][K/32] q4_0, A: [M][K] f32, dst: [M][N] f32.
 * wdata: M*K/32*34 bytes of scratch (the reference's params->wdata).  nth<=0 -> all threads. */
ORC_API int orc_mul_mat_q4_0_f32(const void* w, const float* a, float* dst, int N, int K, int M, void* wdata, int nth) {
  const size_t arow = (size_t)K / QK * sizeof(orc_q8_0);
  const size_t wrow = (size_t)K / QK * sizeof(orc_q4_0);
  for (int m = 0; m < M; ++m) orc_quantize_row_q8_0(a + (size_t)m * K, (char*)wdata + m * arow, K);
#ifdef _OPENMP
  if (nth <= 0) nth = omp_get_max_threads();
#else
  nth = 1;
#endif
#pragma omp parallel for num_threads(nth) schedule(static)
  for (int n = 0; n < N; ++n)
    for (int m = 0; m < M; ++m)
      orc_vec_dot_q4_0_q8_0(K, &dst[(size_t)m * N + n], (const char*)w + n * wrow, (const char*)wdata + m * arow);
  return nth;
}

/* greedy argmax, lowest index wins ties (models/model_utils/model_utils.cpp:2963-2985) */
ORC_API int orc_argmax_f32(const float* x, int n) {
  int best = 0;
  for (int i = 1; i < n; ++i)
    if (x[i] > x[best]) best = i;
  return best;
}
