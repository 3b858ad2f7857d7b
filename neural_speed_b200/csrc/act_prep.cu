// act_prep.cu -- activation preparation kernels: the device counterpart of the reference's activation prologues
//   ggml  : quantize_row_q8_0                 (neural_speed/vectors/cpu/quantize.h:447-560, x86 body: id = 127/amax, RNE)
//   BesTLA: ActivationKBlockQuantize          (bestla/bestla/bestla_prologue_a.h:105-214) =
//           quantize_fp_u8_colblock / _s8_    (bestla/bestla/kernel_ref.h:1825 / :1886)
//           ShuffleActivationKBlock*          (bestla_prologue_a.h:299-424): column gather by g_idx before quantisation
// Output goes to a small device workspace in exactly the byte image the matmul kernels copy into shared memory.
#include "nsb.cuh"

namespace {

// byte position inside an 8-group so that dp4a operands line up with the NSB4 nibble order (see nsb.cuh)
__device__ __forceinline__ int perm8_pos(int e) { return ((e & 3) << 1) | (e >> 2); }

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_isum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// utils::cast<float,uint8_t> (bestla_utils.h:515-521)
__device__ __forceinline__ int cast_u8(float x) {
  if (x != x) return 0;  // NaN -> 0 as on x86 (see cast_s8)
  x += 0.5f;
  x = fminf(x, 255.f);
  x = fmaxf(x, 0.f);
  return (int)x;
}
// utils::cast<float,int8_t> (bestla_utils.h:507-513)
__device__ __forceinline__ int cast_s8(float x) {
  // all-zero block: scale is denormal, 1/scale = inf, 0*inf = NaN.  On the reference's x86 build the NaN survives
  // std::min/std::max and converts to 0 (cvttss2si -> INT_MIN -> int8 0); CUDA's fminf would return 127 instead.
  if (x != x) return 0;
  x = roundf(x);
  x = fminf(x, 127.f);
  x = fmaxf(x, -128.f);
  return (int)x;
}

// One warp per (row m, quantisation block b).  COMP: NS_COMP_Q8_0 | NS_COMP_INT8 (u8 asym) | NS_COMP_INT8_S8.
template <int COMP>
__global__ void __launch_bounds__(256) act_quant_kernel(const float* __restrict__ A, int lda, int M, int K, int kpad,
                                                        int group, int ngroups, const int* __restrict__ shuffle,
                                                        int perm8, uint8_t* __restrict__ aq, int2* __restrict__ meta,
                                                        int meta_stride, int act_row, int ring_layout) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (gw >= M * ngroups) return;
  const int m = gw / ngroups, b = gw - m * ngroups;
  const int k0 = b * group;
  const int kend = min(k0 + group, K);
  const int kend_pad = (b == ngroups - 1) ? kpad : kend;  // last block also owns the zero padding up to kpad
  const float* row = A + (size_t)m * lda;

  // pass 1: range
  const bool full = (k0 + group <= K);
  float vmax = (COMP == NS_COMP_INT8 && !full) ? 0.f : ((COMP == NS_COMP_Q8_0) ? 0.f : 1.17549435e-38f);
  float vmin = 0.f;
  for (int k = k0 + lane; k < kend; k += 32) {
    const float v = row[shuffle ? shuffle[k] : k];
    if (COMP == NS_COMP_INT8) {
      vmax = fmaxf(v, vmax);
      vmin = fminf(v, vmin);
    } else {
      vmax = fmaxf(vmax, fabsf(v));
    }
  }
  vmax = warp_max(vmax);
  if (COMP == NS_COMP_INT8) vmin = warp_min(vmin);

  float scale, rscale;
  int za = 0;
  if (COMP == NS_COMP_Q8_0) {
    // d = amax/127 rounded to fp16 (block_q8_0.d); the multiplier is 127/amax (x86 body of quantize_row_q8_0)
    scale = __half2float(__float2half_rn(vmax / 127.f));
    rscale = vmax != 0.f ? 127.f / vmax : 0.f;
  } else if (COMP == NS_COMP_INT8) {
    scale = (vmax - vmin) / 255;
    za = cast_u8((0 - vmin) / scale);
    rscale = 1.f / scale;
  } else {
    scale = vmax / 127;
    rscale = 1.f / scale;
  }

  // pass 2: quantise 32 elements (one chunk) per iteration, emit permuted bytes + chunk meta
  uint8_t* qrow = aq + (size_t)m * act_row;
  for (int kc = k0; kc < kend_pad; kc += 32) {
    const int k = kc + lane;
    int q;
    if (k < kend) {
      const float v = row[shuffle ? shuffle[k] : k];
      if (COMP == NS_COMP_Q8_0) {
        q = __float2int_rn(v * rscale);  // round-half-even like _mm256_round_ps(NEAREST)
      } else if (COMP == NS_COMP_INT8) {
        const int qt = (int)roundf(v * rscale);  // cast<float,int>
        q = cast_u8((float)za + (float)qt);
      } else {
        q = cast_s8(v * rscale);
      }
    } else {
      q = za;  // padding contributes (a - za) == 0
    }
    const int sa = warp_isum(q);
    const int pos = perm8 ? ((lane & ~7) | perm8_pos(lane & 7)) : lane;
    if (ring_layout) {
      // gemv_ring.cu image: per super-block of 32 chunks, the first 16 bytes of every chunk, then the second 16 bytes
      const int c = kc >> 5;
      qrow[(c >> 5) * 1024 + (pos >> 4) * 512 + (c & 31) * 16 + (pos & 15)] = (uint8_t)q;
    } else {
      qrow[kc + pos] = (uint8_t)q;
    }
    if (lane == 0) meta[(size_t)m * meta_stride + (kc >> 5)] = make_int2(__float_as_int(scale), (sa & 0xffff) | (za << 16));
  }
}

// fp32 / bf16-rounded activations, natural order, zero padded to kpad
__global__ void __launch_bounds__(256) act_copy_kernel(const float* __restrict__ A, int lda, int M, int K, int kpad,
                                                       const int* __restrict__ shuffle, int round_bf16,
                                                       float* __restrict__ af) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)M * kpad) return;
  const int m = (int)(idx / kpad), k = (int)(idx - (size_t)m * kpad);
  float v = 0.f;
  if (k < K) v = A[(size_t)m * lda + (shuffle ? shuffle[k] : k)];
  if (round_bf16) v = __bfloat162float(__float2bfloat16_rn(v));
  af[idx] = v;
}

// plain quantiser with un-permuted, un-fused outputs: the parity-test view of the same arithmetic
template <int COMP>
__global__ void act_quant_plain_kernel(const float* __restrict__ A, int lda, int M, int K, int group, int ngroups,
                                       uint8_t* __restrict__ q, float* __restrict__ scales, int* __restrict__ zps) {
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (gw >= M * ngroups) return;
  const int m = gw / ngroups, b = gw - m * ngroups;
  const int k0 = b * group, kend = min(k0 + group, K);
  const float* row = A + (size_t)m * lda;
  const bool full = (k0 + group <= K);
  float vmax = (COMP == NS_COMP_INT8 && !full) ? 0.f : ((COMP == NS_COMP_Q8_0) ? 0.f : 1.17549435e-38f);
  float vmin = 0.f;
  for (int k = k0 + lane; k < kend; k += 32) {
    const float v = row[k];
    if (COMP == NS_COMP_INT8) {
      vmax = fmaxf(v, vmax);
      vmin = fminf(v, vmin);
    } else {
      vmax = fmaxf(vmax, fabsf(v));
    }
  }
  vmax = warp_max(vmax);
  if (COMP == NS_COMP_INT8) vmin = warp_min(vmin);
  float scale, rscale;
  int za = 0;
  if (COMP == NS_COMP_Q8_0) {
    scale = __half2float(__float2half_rn(vmax / 127.f));
    rscale = vmax != 0.f ? 127.f / vmax : 0.f;
  } else if (COMP == NS_COMP_INT8) {
    scale = (vmax - vmin) / 255;
    za = cast_u8((0 - vmin) / scale);
    rscale = 1.f / scale;
  } else {
    scale = vmax / 127;
    rscale = 1.f / scale;
  }
  for (int k = k0 + lane; k < kend; k += 32) {
    const float v = row[k];
    int qv;
    if (COMP == NS_COMP_Q8_0) qv = __float2int_rn(v * rscale);
    else if (COMP == NS_COMP_INT8) qv = cast_u8((float)za + (float)(int)roundf(v * rscale));
    else qv = cast_s8(v * rscale);
    q[(size_t)m * K + k] = (uint8_t)qv;
  }
  if (lane == 0) {
    scales[(size_t)m * ngroups + b] = scale;
    if (COMP == NS_COMP_INT8 && zps) zps[(size_t)m * ngroups + b] = za;
  }
}

}  // namespace

static size_t meta_stride_of(int kpad) { return ns_round_up((size_t)(kpad >> 5), 2); }  // int2 units, 16-B multiple

size_t ns_act_workspace_bytes(int m, int kpad) {
  const size_t i8 = (size_t)m * ns_round_up((size_t)kpad, 1024) + (size_t)m * meta_stride_of(kpad) * sizeof(int2);
  const size_t f32 = (size_t)m * kpad * sizeof(float);
  return ns_round_up(i8 > f32 ? i8 : f32, 256);
}

template <typename... Args>
static cudaError_t launch_pdl(void (*kern)(Args...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

int ns_launch_act_prep(const float* act, int lda, int m, const ns_weight* w, void* ws, cudaStream_t st) {
  const int kpad = w->kpad;
  if (w->comp == NS_COMP_F32 || w->comp == NS_COMP_BF16) {
    const size_t total = (size_t)m * kpad;
    const int blocks = (int)((total + 255) / 256);
    NS_CUDA_TRY(launch_pdl(act_copy_kernel, dim3(blocks), dim3(256), 0, st, act, lda, m, w->k, kpad,
                           (const int*)w->shuffle, (int)(w->comp == NS_COMP_BF16), (float*)ws));
    ns_count_launch();
    return NS_OK;
  }
  uint8_t* aq = (uint8_t*)ws;
  const int ring_layout = (w->wfmt == NS_W_S4) ? 1 : 0;  // consumed by gemv_ring.cu; other formats use natural rows
  const int act_row = ring_layout ? (int)ns_round_up((size_t)kpad, 1024) : kpad;
  int2* meta = (int2*)((char*)ws + ns_round_up((size_t)m * act_row, 16));
  const int ms = (int)meta_stride_of(kpad);
  // activation blocks: ggml Q8_0 is always 32; BesTLA uses the weight's K-block (bestla_prologue_a.h:133)
  const int group = (w->comp == NS_COMP_Q8_0) ? 32 : w->group;
  const int ngroups = (w->k + group - 1) / group;
  const int warps = m * ngroups;
  const int blocks = (warps + 7) / 8;
  const int perm8 = (w->wfmt == NS_W_S8) ? 0 : 1;
  cudaError_t e;
  if (w->comp == NS_COMP_Q8_0)
    e = launch_pdl(act_quant_kernel<NS_COMP_Q8_0>, dim3(blocks), dim3(256), 0, st, act, lda, m, w->k, kpad, group,
                   ngroups, (const int*)w->shuffle, perm8, aq, meta, ms, act_row, ring_layout);
  else if (w->comp == NS_COMP_INT8)
    e = launch_pdl(act_quant_kernel<NS_COMP_INT8>, dim3(blocks), dim3(256), 0, st, act, lda, m, w->k, kpad, group,
                   ngroups, (const int*)w->shuffle, perm8, aq, meta, ms, act_row, ring_layout);
  else
    e = launch_pdl(act_quant_kernel<NS_COMP_INT8_S8>, dim3(blocks), dim3(256), 0, st, act, lda, m, w->k, kpad, group,
                   ngroups, (const int*)w->shuffle, perm8, aq, meta, ms, act_row, ring_layout);
  NS_CUDA_TRY(e);
  ns_count_launch();
  return NS_OK;
}

extern "C" int ns_device_quantize_act(const float* act_dev, int lda, int m, int k, int group, int comp, void* q_dev,
                                      float* scale_dev, int* zp_dev, void* queue) {
  if (int rc = ns_ensure_device()) return rc;
  cudaStream_t st = (cudaStream_t)queue;
  if (comp == NS_COMP_Q8_0) group = 32;
  if (group <= 0) group = k;
  const int ngroups = (k + group - 1) / group;
  const int blocks = (m * ngroups + 7) / 8;
  if (comp == NS_COMP_Q8_0)
    act_quant_plain_kernel<NS_COMP_Q8_0><<<blocks, 256, 0, st>>>(act_dev, lda, m, k, group, ngroups, (uint8_t*)q_dev, scale_dev, zp_dev);
  else if (comp == NS_COMP_INT8)
    act_quant_plain_kernel<NS_COMP_INT8><<<blocks, 256, 0, st>>>(act_dev, lda, m, k, group, ngroups, (uint8_t*)q_dev, scale_dev, zp_dev);
  else if (comp == NS_COMP_INT8_S8)
    act_quant_plain_kernel<NS_COMP_INT8_S8><<<blocks, 256, 0, st>>>(act_dev, lda, m, k, group, ngroups, (uint8_t*)q_dev, scale_dev, zp_dev);
  else {
    ns_set_error("ns_device_quantize_act: unsupported comp %d", comp);
    return NS_E_INVALID;
  }
  NS_CUDA_TRY(cudaGetLastError());
  ns_count_launch();
  return NS_OK;
}
