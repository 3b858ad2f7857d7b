/*
 * oracle/ref_ne.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Compiles the reference's own graph engine, neural_speed/core/ne_layers.c, where it lies under /root/reference (nothing is
 * copied) and drives single ops of the Llama eval graph through its PUBLIC API (ne_init / ne_new_tensor_* / ne_rope_inplace /
 * ne_soft_max_inplace / ne_mul_mat / ne_rms_norm / ne_graph_compute), so oracle/llama_model.py's element-wise restatements
 * can be pinned against the real implementation (tests/test_oracle_vs_ref.py).
 *
 * ne_layers.c calls into BesTLA through ne_bestla.h; those entry points are stubbed here:
 *   - bestla_support() answers "not supported" so every node takes the ggml path of ne_layers.c itself,
 *   - bestla_parallel_for() restates the nth == 1 branch of core/layers/ne_bestla.cpp:42-52 (INIT, COMPUTE, FINALIZE),
 *   - bestla_layernormalization() (called unconditionally by ne_compute_forward_rms_norm_f32, ne_layers.c:6624) is provided
 *     by ref_ne_stubs.cpp from the reference's bestla/kernel_ref.h (its portable body; kernel_avx2.h / kernel_avx512f.h pull
 *     in xbyak and cannot be compiled here -- they vectorise the same sums, i.e. differ in summation order only),
 *   - the matmul / attention fusions abort: the harness never builds such nodes.
 */
#include "core/ne_layers.c"

#define REF_API __attribute__((visibility("default")))

static void ref_ne_die(const char* what) {
  fprintf(stderr, "oracle/ref_ne: unexpected call into stub %s\n", what);
  abort();
}
void bestla_init(void) {}
int bestla_set_threads(int n) { return n; }
void* bestla_get_thread_handle(void) { return NULL; }
void bestla_timer(bool b) { (void)b; }
bool bestla_support(struct ne_tensor* node, int n_threads, size_t* workspace, size_t* dev_workspace) {
  (void)node;
  (void)n_threads;
  *workspace = 0;
  *dev_workspace = 0;
  return false;
}
enum ne_backend bestla_backend_support(struct ne_tensor* src0, struct ne_tensor* src1, enum ne_op op) {
  (void)src0;
  (void)src1;
  (void)op;
  return NE_BACKEND_CPU;
}
void bestla_parallel_for(forward_compute_fptr fcomp, struct ne_compute_params* mainparams, struct ne_tensor* node) {
  struct ne_compute_params params = *mainparams; /* harness graphs run with n_threads = 1 */
  params.ith = 0;
  params.nth = 1;
  params.type = NE_TASK_INIT;
  fcomp(&params, node);
  params.type = NE_TASK_COMPUTE;
  fcomp(&params, node);
  params.type = NE_TASK_FINALIZE;
  fcomp(&params, node);
}
void bestla_mul(int batch, int vsize, const float* tensor, const float* vector, int vstep, float* out) {
  for (int b = 0; b < batch; ++b)
    for (int i = 0; i < vsize; ++i) out[(size_t)b * vsize + i] = tensor[(size_t)b * vsize + i] * vector[(size_t)b * vstep + i];
}
void bestla_add(int batch, int vsize, const float* tensor, const float* vector, int vstep, float* out) {
  for (int b = 0; b < batch; ++b)
    for (int i = 0; i < vsize; ++i) out[(size_t)b * vsize + i] = tensor[(size_t)b * vsize + i] + vector[(size_t)b * vstep + i];
}
#define DIE_STUB(ret, name, args) \
  ret name args {                 \
    ref_ne_die(#name);            \
  }
DIE_STUB(unsigned long long, bestla_f32f32_get_workspace_size, (int m, int n, int k, void* w))
DIE_STUB(void, bestla_f32f32_forward, (float* a, void* w, float* o, int m, int n, int k, int lda, int ldo, void* ws))
DIE_STUB(bool, bestla_fusion_add_f32f32_support, (void* w, int m, int n, int k))
DIE_STUB(void, bestla_fusion_add_f32f32_forward,
         (float* a, void* w, float* b, float* o, int m, int n, int k, int lda, int ldo, bool bc, void* ws))
DIE_STUB(unsigned long long, bestla_fusion_QKV_f32f32_get_workspace_size, (int m, int n, int k, void* w))
DIE_STUB(bool, bestla_fusion_QKV_f32f32_support, (void* q, void* k_, void* v, int m, int n, int k))
DIE_STUB(void, bestla_fusion_QKV_f32f32_forward,
         (float* a, void* q, void* k_, void* v, float* o, int m, int n, int k, int lda, int ldo, void* ws))
DIE_STUB(unsigned long long, bestla_fusion_FFN_f32f32_get_workspace_size, (int s, int fi, int fm, int fo, void* w1, void* w2))
DIE_STUB(bool, bestla_fusion_FFN_SiLu_f32f32_support, (void* w1, void* w2, void* w3, int s, int fi, int fm, int fo))
DIE_STUB(void, bestla_fusion_FFN_SiLu_f32f32_forward,
         (float* a, void* w1, void* w2, void* w3, float* t1, float* t2, float* o, int s, int fi, int fm, int fo, void* ws))
DIE_STUB(bool, bestla_fusion_FFN_Gelu_Mul_f32f32_support, (void* w1, void* w2, void* w3, int s, int fi, int fm, int fo))
DIE_STUB(void, bestla_fusion_FFN_Gelu_Mul_f32f32_forward,
         (float* a, void* w1, void* w2, void* w3, float* t1, float* t2, float* o, int s, int fi, int fm, int fo, void* ws))
DIE_STUB(bool, bestla_fusion_FFN_GeLu_f32f32_support, (void* w1, void* w2, int s, int fi, int fm, int fo))
DIE_STUB(void, bestla_fusion_FFN_GeLu_f32f32_forward,
         (float* a, void* w1, void* w2, float* t1, float* o, int s, int fi, int fm, int fo, void* ws))
DIE_STUB(bool, bestla_fusion_FFN_Add_GeLu_f32f32_support, (void* w1, void* w2, int s, int fi, int fm, int fo))
DIE_STUB(void, bestla_fusion_FFN_Add_GeLu_f32f32_forward,
         (float* a, void* w1, void* w2, float* b1, float* b2, float* t1, float* o, int s, int fi, int fm, int fo, bool bc, void* ws))
DIE_STUB(void, bestla_unpackweight_fp32, (void* w, int n, int k, float* f, int ld))
DIE_STUB(void, bestla_packweight_copyattr, (const float* f, void* d, int n, int k, int ld, void* s))

#include "ref_ne_harness.h"
