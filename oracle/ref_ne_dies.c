/* oracle/ref_ne_dies.c -- TEST INFRASTRUCTURE ONLY. */
/* Symbols of other reference translation units (mha_dense.cpp, conv.cpp, argsort.cpp, memory.cpp) that ne_layers.c references
 * but the harness graphs never reach.  Defined here (no reference header in scope) so the names resolve at load time. */
#include <stdio.h>
#include <stdlib.h>
#define REF_NE_DIE(name)                                                          \
  void name(void) {                                                    \
    fprintf(stderr, "oracle/ref_ne: unexpected call into stub " #name "\n"); \
    abort();                                                                 \
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
