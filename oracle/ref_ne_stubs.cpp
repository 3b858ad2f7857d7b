// oracle/ref_ne_stubs.cpp -- TEST INFRASTRUCTURE ONLY.  The one BesTLA entry point ne_layers.c needs for the ops the harness
// drives: bestla_layernormalization (core/layers/ne_bestla.cpp:114-117 -> BTLALayerNorm -> kernel::wrapper::LayerNormalization,
// bestla/kernel_wrapper.h:1372-1404).  The reference picks avx512f:: / avx2:: / ref::layernorm at run time through xbyak's CPUID
// wrapper; only the portable ref:: body (bestla/kernel_ref.h:2199-2250) compiles without xbyak, so that is what is pinned.
#include <cstddef>

#include "bestla/kernel_ref.h"

extern "C" void bestla_layernormalization(int norm_count, int norm_size, bool isrms, float epsilon, const float* FpIn, float* FpOut) {
  for (int i = 0; i < norm_count; ++i)
    bestla::kernel::ref::layernorm<float>(FpIn + (size_t)i * norm_size, nullptr, nullptr, epsilon, norm_size,
                                          FpOut + (size_t)i * norm_size, nullptr, nullptr, isrms);
}

// Symbols of other reference translation units (mha_dense.cpp, conv.cpp, argsort.cpp, memory.cpp) that ne_layers.c references
// but the harness graphs never reach.  Defined here (no reference header in scope) so the names resolve at load time.
#include <cstdio>
#include <cstdlib>
#define REF_NE_DIE(name)                                                          \
  extern "C" void name(void) {                                                    \
    std::fprintf(stderr, "oracle/ref_ne: unexpected call into stub " #name "\n"); \
    std::abort();                                                                 \
  }
REF_NE_DIE(bestla_fusion_attn_fp32_fp16_fp16_fp32_forward)
REF_NE_DIE(bestla_fusion_attn_workspace_size)
REF_NE_DIE(bestla_reordered_attn_fp32_forward)
REF_NE_DIE(bestla_reordered_attn_fp32_shift_rope_k)
REF_NE_DIE(bestla_reordered_attn_fp32_update_k)
REF_NE_DIE(bestla_reordered_attn_fp32_update_v)
REF_NE_DIE(ne_attention_padding_mask_f32_forward)
REF_NE_DIE(ne_compute_forward_argsort)
REF_NE_DIE(ne_compute_forward_conv_1d)
REF_NE_DIE(ne_compute_forward_conv_1d_1s)
REF_NE_DIE(ne_compute_forward_conv_1d_2s)
