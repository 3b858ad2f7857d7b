/*
 * oracle/ref_ggml.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin harness that compiles the *reference's own* ggml-style Q4_0/Q8_0 code straight from
 * /root/reference (headers are #included where they lie; nothing is copied into this repo) and
 * exposes it with a C ABI so tests can pin oracle/oracle_ggml.c against the real thing and so
 * bench.py --impl reference can time the reference's CPU path.
 *
 * Reference sources compiled (see oracle/Makefile for the -I path):
 *   neural_speed/core/data_types.h          block_q4_0 / block_q8_0, fp16 helpers
 *   neural_speed/vectors/cpu/quantize.h     quantize_row_q4_0[_reference], quantize_row_q8_0[_reference],
 *                                           dequantize_row_q4_0
 *   neural_speed/core/layers/vec_dot.h      ne_vec_dot_q4_0_q8_0
 * The mul_mat driver below restates ne_compute_forward_mul_mat_q_f32
 * (neural_speed/core/ne_layers.c:7085-7203): INIT quantises every src1 row to Q8_0 into wdata
 * (:7146-7157), COMPUTE splits src0 rows evenly over threads (:7166-7170) and calls vec_dot_q
 * per (src1 row, src0 row) (:7178-7203).  The reference's thread pool is replaced by OpenMP.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "core/layers/vec_dot.h"

#define REF_API __attribute__((visibility("default")))

/* ne_init() fills this table (neural_speed/core/ne_layers.c:741-746); the headers only declare it. */
REF_API void ref_ggml_init(void) {
  static int done = 0;
  if (done) return;
  for (int i = 0; i < (1 << 16); ++i) {
    uint16_t ui = (uint16_t)i;
    ne_fp16_t h;
    memcpy(&h, &ui, sizeof(h));
    table_f32_f16[i] = NE_COMPUTE_FP16_TO_FP32(h);
  }
  done = 1;
}

REF_API int ref_ggml_simd_level(void) {
#if defined(__AVX2__)
  return 2;
#elif defined(__AVX__)
  return 1;
#else
  return 0;
#endif
}

REF_API void ref_quantize_row_q4_0(const float* x, void* y, int k) { quantize_row_q4_0(x, y, k); }
REF_API void ref_dequantize_row_q4_0(const void* x, float* y, int k) { dequantize_row_q4_0((const block_q4_0*)x, y, k); }
REF_API void ref_quantize_row_q8_0(const float* x, void* y, int k) { quantize_row_q8_0(x, y, k); }
REF_API void ref_quantize_row_q8_0_reference(const float* x, void* y, int k) {
  quantize_row_q8_0_reference(x, (block_q8_0*)y, k);
}
REF_API void ref_dequantize_row_q8_0(const void* x, float* y, int k) { dequantize_row_q8_0(x, y, k); }
REF_API void ref_vec_dot_q4_0_q8_0(int n, float* s, const void* vx, const void* vy) {
  ne_vec_dot_q4_0_q8_0(n, s, vx, vy);
}
REF_API void ref_vec_dot_q8_0_q8_0(int n, float* s, const void* vx, const void* vy) {
  ne_vec_dot_q8_0_q8_0(n, s, vx, vy);
}
REF_API float ref_fp16_to_fp32(uint16_t h) {
  ne_fp16_t v;
  memcpy(&v, &h, 2);
  return NE_COMPUTE_FP16_TO_FP32(v);
}
REF_API uint16_t ref_fp32_to_fp16(float f) {
  ne_fp16_t v = NE_COMPUTE_FP32_TO_FP16(f);
  uint16_t h;
  memcpy(&h, &v, 2);
  return h;
}

/* dst[m][n] = sum_k dequant(W[n][k]) * q8(A[m][k]);  W: N rows of K/32 block_q4_0, A: [M][K] f32, dst: [M][N] f32.
 * wdata must hold M*K/32*sizeof(block_q8_0) bytes.  nth<=0 -> all OpenMP threads. Returns threads used. */
REF_API int ref_mul_mat_q4_0_f32(const void* w, const float* a, float* dst, int N, int K, int M, void* wdata, int nth) {
  ref_ggml_init();
  const size_t row_size = (size_t)K / QK8_0 * sizeof(block_q8_0);
  const size_t nb01 = (size_t)K / QK4_0 * sizeof(block_q4_0);
  /* NE_TASK_INIT runs on one thread in the reference (core/layers/ne_bestla.cpp:56-59) */
  for (int m = 0; m < M; ++m) quantize_row_q8_0(a + (size_t)m * K, (char*)wdata + m * row_size, K);
#ifdef _OPENMP
  if (nth <= 0) nth = omp_get_max_threads();
#else
  nth = 1;
#endif
#pragma omp parallel num_threads(nth)
  {
#ifdef _OPENMP
    const int ith = omp_get_thread_num();
    const int nthr = omp_get_num_threads();
#else
    const int ith = 0, nthr = 1;
#endif
    const int64_t dr = (N + nthr - 1) / nthr;
    const int64_t ir10 = dr * ith;
    const int64_t ir11 = MIN(ir10 + dr, N);
    for (int m = 0; m < M; ++m) {
      const char* src1_col = (const char*)wdata + m * row_size;
      float* dst_col = dst + (size_t)m * N;
      for (int64_t ir = ir10; ir < ir11; ++ir) {
        ne_vec_dot_q4_0_q8_0(K, &dst_col[ir], (const char*)w + ir * nb01, src1_col);
      }
    }
  }
  return nth;
}

/* ---- Q6_K weights x Q8_K activations (quantize.h:877-1060, vec_dot.h:744-990): the type llama.cpp "Q4_0" GGUF files use
 * for output.weight.  Same driver shape as above (ne_layers.c:7085-7203 is type-generic through quantize_fns). */
REF_API void ref_quantize_row_q6_K(const float* x, void* y, int k) { quantize_row_q6_K(x, y, k); }
REF_API void ref_dequantize_row_q6_K(const void* x, float* y, int k) { dequantize_row_q6_K((const block_q6_K*)x, y, k); }
REF_API void ref_quantize_row_q8_K(const float* x, void* y, int k) { quantize_row_q8_K(x, y, k); }
REF_API void ref_vec_dot_q6_K_q8_K(int n, float* s, const void* vx, const void* vy) { ggml_vec_dot_q6_K_q8_K(n, s, vx, vy); }
REF_API int ref_sizeof_block_q6_K(void) { return (int)sizeof(block_q6_K); }
REF_API int ref_sizeof_block_q8_K(void) { return (int)sizeof(block_q8_K); }
REF_API int ref_mul_mat_q6_K_f32(const void* w, const float* a, float* dst, int N, int K, int M, void* wdata, int nth) {
  ref_ggml_init();
  const size_t row_size = (size_t)K / QK_K * sizeof(block_q8_K);
  const size_t nb01 = (size_t)K / QK_K * sizeof(block_q6_K);
  for (int m = 0; m < M; ++m) quantize_row_q8_K(a + (size_t)m * K, (char*)wdata + m * row_size, K);
#ifdef _OPENMP
  if (nth <= 0) nth = omp_get_max_threads();
#else
  nth = 1;
#endif
#pragma omp parallel num_threads(nth)
  {
#ifdef _OPENMP
    const int ith = omp_get_thread_num();
    const int nthr = omp_get_num_threads();
#else
    const int ith = 0, nthr = 1;
#endif
    const int64_t dr = (N + nthr - 1) / nthr;
    const int64_t ir10 = dr * ith;
    const int64_t ir11 = MIN(ir10 + dr, N);
    for (int m = 0; m < M; ++m) {
      const char* src1_col = (const char*)wdata + m * row_size;
      float* dst_col = dst + (size_t)m * N;
      for (int64_t ir = ir10; ir < ir11; ++ir) ggml_vec_dot_q6_K_q8_K(K, &dst_col[ir], (const char*)w + ir * nb01, src1_col);
    }
  }
  return nth;
}
