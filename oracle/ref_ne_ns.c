/*
 * oracle/ref_ne_ns.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * The reference's own graph engine, neural_speed/core/ne_layers.c, compiled where it lies under /root/reference and linked
 * against libns_b200.so INSTEAD of the reference's core/layers/*.cpp: every bestla_* entry point ne_layers.c calls
 * (bestla_support, bestla_parallel_for, bestla_f32f32_forward, bestla_fusion_QKV/FFN_*, bestla_mul/add/layernormalization ...)
 * resolves to the CUDA drop-ins -- INTEGRATION.md A ("no source change") exercised for real: ne_graph_compute sizes its work
 * buffer through bestla_support, enters BesTLA nodes once (n_tasks = 1) and the kernels run on the GPU.
 * tests/test_gpu_ne_dropin.py drives a tiny BesTLA-blob Llama through it and compares with the CPU oracle.
 */
#include "core/ne_layers.c"

#define REF_API __attribute__((visibility("default")))

#include "ref_ne_harness.h"
