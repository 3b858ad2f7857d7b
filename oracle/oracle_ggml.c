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

/* ================================================================================================================
 * Q6_K weights x Q8_K activations (llama.cpp "Q4_0" GGUF files keep output.weight in Q6_K).
 *   block_q6_K / block_q8_K          core/data_types.h:133-145
 *   nearest_int, make_qx_quants      vectors/cpu/quantize.h:801-875
 *   quantize_row_q6_K_reference      :877-954      dequantize_row_q6_K :956-999
 *   quantize_row_q8_K_reference      :1020-1055
 *   ggml_vec_dot_q6_K_q8_K (AVX2)    core/layers/vec_dot.h:907-983
 * ================================================================================================================ */
#define QKK 256
#pragma pack(push, 1)
typedef struct { uint8_t ql[QKK / 2]; uint8_t qh[QKK / 4]; int8_t scales[QKK / 16]; uint16_t d; } orc_q6_K; /* 210 bytes */
#pragma pack(pop)
typedef struct { float d; int8_t qs[QKK]; int16_t bsums[QKK / 16]; } orc_q8_K;                            /* 292 bytes */

static int orc_nearest_int(float fval) { /* quantize.h:801-807: round half to even through the 1.5*2^23 trick */
  float val = fval + 12582912.f;
  int i;
  memcpy(&i, &val, sizeof(int));
  return (i & 0x007fffff) - 0x00400000;
}
#define ORC_MAX(a, b) ((a) > (b) ? (a) : (b))
#define ORC_MIN(a, b) ((a) < (b) ? (a) : (b))

/* make_qx_quants with rmse_type 1 (the only mode Q6_K uses, quantize.h:889).  The reference's default x86 build contracts
 * a*b+c into fma (-ffp-contract=fast); the explicit fmaf calls below reproduce what gcc emits for these expressions. */
static float orc_make_qx_quants_rmse1(int n, int nmax, const float* x, int8_t* L) {
  float max = 0, amax = 0;
  for (int i = 0; i < n; ++i) {
    float ax = fabsf(x[i]);
    if (ax > amax) { amax = ax; max = x[i]; }
  }
  if (amax < 1e-30f) {
    for (int i = 0; i < n; ++i) L[i] = 0;
    return 0.f;
  }
  float iscale = -nmax / max;
  float sumlx = 0, suml2 = 0;
  for (int i = 0; i < n; ++i) {
    int l = orc_nearest_int(iscale * x[i]);
    l = ORC_MAX(-nmax, ORC_MIN(nmax - 1, l));
    L[i] = (int8_t)(l + nmax);
    float w = x[i] * x[i];
    sumlx = fmaf(w * x[i], (float)l, sumlx);
    suml2 = fmaf(w * (float)l, (float)l, suml2);
  }
  float scale = sumlx / suml2;
  float best = scale * sumlx;
  for (int is = -9; is <= 9; ++is) {
    if (is == 0) continue;
    iscale = -(nmax + 0.1f * is) / max;
    sumlx = suml2 = 0;
    for (int i = 0; i < n; ++i) {
      int l = orc_nearest_int(iscale * x[i]);
      l = ORC_MAX(-nmax, ORC_MIN(nmax - 1, l));
      float w = x[i] * x[i];
      sumlx = fmaf(w * x[i], (float)l, sumlx);
      suml2 = fmaf(w * (float)l, (float)l, suml2);
    }
    if (suml2 > 0 && sumlx * sumlx > best * suml2) {
      for (int i = 0; i < n; ++i) {
        int l = orc_nearest_int(iscale * x[i]);
        L[i] = (int8_t)(nmax + ORC_MAX(-nmax, ORC_MIN(nmax - 1, l)));
      }
      scale = sumlx / suml2;
      best = scale * sumlx;
    }
  }
  return scale;
}

ORC_API void orc_quantize_row_q6_K(const float* x, void* vy, int k) {
  orc_q6_K* y = (orc_q6_K*)vy;
  int8_t L[QKK];
  float scales[QKK / 16];
  for (int i = 0; i < k / QKK; ++i, x += QKK) {
    float max_scale = 0, max_abs_scale = 0;
    for (int ib = 0; ib < QKK / 16; ++ib) {
      const float scale = orc_make_qx_quants_rmse1(16, 32, x + 16 * ib, L + 16 * ib);
      scales[ib] = scale;
      const float a = fabsf(scale);
      if (a > max_abs_scale) { max_abs_scale = a; max_scale = scale; }
    }
    if (!max_abs_scale) {
      memset(&y[i], 0, sizeof(orc_q6_K));
      continue;
    }
    float iscale = -128.f / max_scale;
    y[i].d = orc_fp32_to_fp16(1 / iscale);
    for (int ib = 0; ib < QKK / 16; ++ib) y[i].scales[ib] = (int8_t)ORC_MIN(127, orc_nearest_int(iscale * scales[ib]));
    for (int j = 0; j < QKK / 16; ++j) {
      float d = orc_fp16_to_fp32(y[i].d) * y[i].scales[j];
      if (!d) continue;
      for (int ii = 0; ii < 16; ++ii) {
        int l = orc_nearest_int(x[16 * j + ii] / d);
        l = ORC_MAX(-32, ORC_MIN(31, l));
        L[16 * j + ii] = (int8_t)(l + 32);
      }
    }
    uint8_t* ql = y[i].ql;
    uint8_t* qh = y[i].qh;
    for (int j = 0; j < QKK; j += 128, ql += 64, qh += 32)
      for (int l = 0; l < 32; ++l) {
        ql[l] = (uint8_t)((L[j + l] & 0xF) | ((L[j + l + 64] & 0xF) << 4));
        ql[l + 32] = (uint8_t)((L[j + l + 32] & 0xF) | ((L[j + l + 96] & 0xF) << 4));
        qh[l] = (uint8_t)((L[j + l] >> 4) | ((L[j + l + 32] >> 4) << 2) | ((L[j + l + 64] >> 4) << 4) | ((L[j + l + 96] >> 4) << 6));
      }
  }
}

/* element e (0..255) of a block as the signed 6-bit value minus 32, and its scale index (dequantize_row_q6_K :956-999) */
static int orc_q6_K_elem(const orc_q6_K* b, int e, int* sc_idx) {
  const int half = e >> 7, r = e & 127, c = r >> 5, l = r & 31;
  const uint8_t lo = b->ql[64 * half + 32 * (c & 1) + l];
  const int nib = (c >> 1) ? (lo >> 4) : (lo & 0xF);
  const int hi = (b->qh[32 * half + l] >> (2 * c)) & 3;
  *sc_idx = 8 * half + 2 * c + (l >> 4);
  return (nib | (hi << 4)) - 32;
}

ORC_API void orc_dequantize_row_q6_K(const void* vx, float* y, int k) {
  const orc_q6_K* x = (const orc_q6_K*)vx;
  for (int i = 0; i < k / QKK; ++i) {
    const float d = orc_fp16_to_fp32(x[i].d);
    for (int e = 0; e < QKK; ++e) {
      int si;
      const int q = orc_q6_K_elem(&x[i], e, &si);
      y[i * QKK + e] = d * x[i].scales[si] * q; /* (d * sc) * q, left to right as the reference */
    }
  }
}

ORC_API void orc_quantize_row_q8_K(const float* x, void* vy, int k) {
  orc_q8_K* y = (orc_q8_K*)vy;
  for (int i = 0; i < k / QKK; ++i, x += QKK) {
    float max = 0, amax = 0;
    for (int j = 0; j < QKK; ++j) {
      float ax = fabsf(x[j]);
      if (ax > amax) { amax = ax; max = x[j]; }
    }
    if (!amax) {
      y[i].d = 0;
      memset(y[i].qs, 0, QKK);
      continue; /* bsums are left untouched by the reference too */
    }
    const float iscale = -128.f / max;
    for (int j = 0; j < QKK; ++j) y[i].qs[j] = (int8_t)ORC_MIN(127, orc_nearest_int(iscale * x[j]));
    for (int j = 0; j < QKK / 16; ++j) {
      int sum = 0;
      for (int ii = 0; ii < 16; ++ii) sum += y[i].qs[j * 16 + ii];
      y[i].bsums[j] = (int16_t)sum;
    }
    y[i].d = 1 / iscale;
  }
}

/* AVX2 structure (vec_dot.h:907-983): per super-block an int32x8 `sumi`; lane L gathers, from each of the eight
 * 32-element chunks, elements 4L..4L+3 times their 16-group scale; then acc[L] = fma(d, (float)sumi[L], acc[L]) and the
 * final hsum_float_8.  (maddubs never saturates here: 2*63*127 < 32767.) */
ORC_API void orc_vec_dot_q6_K_q8_K(int n, float* s, const void* vx, const void* vy) {
  const orc_q6_K* x = (const orc_q6_K*)vx;
  const orc_q8_K* y = (const orc_q8_K*)vy;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < n / QKK; ++i) {
    const float d = y[i].d * orc_fp16_to_fp32(x[i].d);
    int32_t sumi[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int chunk = 0; chunk < 8; ++chunk)
      for (int L = 0; L < 8; ++L)
        for (int j = 0; j < 4; ++j) {
          const int e = 32 * chunk + 4 * L + j;
          int si;
          const int q = orc_q6_K_elem(&x[i], e, &si);
          sumi[L] += (int32_t)x[i].scales[si] * q * y[i].qs[e];
        }
    for (int L = 0; L < 8; ++L) acc[L] = fmaf(d, (float)sumi[L], acc[L]);
  }
  *s = lanes8_hsum(acc);
}

ORC_API int orc_mul_mat_q6_K_f32(const void* w, const float* a, float* dst, int N, int K, int M, void* wdata, int nth) {
  const size_t arow = (size_t)K / QKK * sizeof(orc_q8_K);
  const size_t wrow = (size_t)K / QKK * sizeof(orc_q6_K);
  for (int m = 0; m < M; ++m) orc_quantize_row_q8_K(a + (size_t)m * K, (char*)wdata + m * arow, K);
#ifdef _OPENMP
  if (nth <= 0) nth = omp_get_max_threads();
#else
  nth = 1;
#endif
#pragma omp parallel for num_threads(nth) schedule(static)
  for (int n = 0; n < N; ++n)
    for (int m = 0; m < M; ++m)
      orc_vec_dot_q6_K_q8_K(K, &dst[(size_t)m * N + n], (const char*)w + n * wrow, (const char*)wdata + m * arow);
  return nth;
}
