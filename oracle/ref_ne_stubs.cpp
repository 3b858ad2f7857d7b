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
