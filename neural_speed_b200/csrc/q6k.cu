// q6k.cu -- ggml Q6_K weights x Q8_K activations (the type llama.cpp "Q4_0" GGUF files keep output.weight in).
//
// Reference path: ne_compute_forward_mul_mat_q_f32 (core/ne_layers.c:7085-7203) with quantize_fns[NE_TYPE_Q6_K]
// (ne_layers.c:320-327): INIT quantises each activation row with quantize_row_q8_K (vectors/cpu/quantize.h:1020-1060),
// COMPUTE calls ggml_vec_dot_q6_K_q8_K (core/layers/vec_dot.h:907-983, the AVX2 body of the reference's default build).
// The kernel reproduces that arithmetic bit for bit: the integer part is exact and the float part keeps the reference's
// shape -- eight fp32 lanes, lane L owning elements 4L..4L+3 of every 32-element chunk, one fma per super-block per lane,
// then the hsum_float_8 order.
//
// Device layout of a Q6_K weight (wfmt NS_W_Q6K), nb = K/256 super-blocks, one row = `pitch` bytes:
//   [ ql : nb x 128 ][ qh : nb x 64 ][ scales : nb x 16 int8  (at sc_off) ][ d : nb x f32 (at zp_off) ]
// i.e. the four members of block_q6_K (core/data_types.h:133-138) split into planes so every load is aligned (the
// 210-byte source block is only 2-byte aligned) and a warp reads each plane with full 32-byte sectors.
#include "nsb.cuh"

namespace {

constexpr int QKK = 256;
constexpr int kSrcBlock = 210;  // sizeof(block_q6_K)
constexpr int kWarps = 8;       // weight rows per CTA

__global__ void __launch_bounds__(256) repack_q6k_kernel(const uint8_t* __restrict__ src, size_t nb01, uint8_t* __restrict__ rows,
                                                         int n, int nb, int pitch, int sc_off, int d_off) {
  const size_t total = (size_t)n * nb;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total * 53; idx += (size_t)gridDim.x * blockDim.x) {
    // 53 work items per super-block: 32 ql words, 16 qh words, 4 scale words, 1 d
    const size_t blk = idx / 53;
    const int item = (int)(idx % 53);
    const int r = (int)(blk / nb), i = (int)(blk % nb);
    const uint8_t* s = src + (size_t)r * nb01 + (size_t)i * kSrcBlock;
    uint8_t* row = rows + (size_t)r * pitch;
    if (item < 32) {
      const uint8_t* p = s + item * 4;
      *(uint32_t*)(row + i * 128 + item * 4) = p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24);
    } else if (item < 48) {
      const int wi = item - 32;
      const uint8_t* p = s + 128 + wi * 4;
      *(uint32_t*)(row + nb * 128 + i * 64 + wi * 4) = p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24);
    } else if (item < 52) {
      const int wi = item - 48;
      const uint8_t* p = s + 192 + wi * 4;
      *(uint32_t*)(row + sc_off + i * 16 + wi * 4) = p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24);
    } else {
      const __half h = __ushort_as_half((unsigned short)(s[208] | (s[209] << 8)));
      *(float*)(row + d_off + i * 4) = __half2float(h);
    }
  }
}

// quantize_row_q8_K_reference (quantize.h:1020-1055): one warp per (activation row, super-block).
// ws layout: qs [m][K] int8, then d [m][nb] f32 at byte offset roundup(m*K, 16).
__global__ void __launch_bounds__(256) act_quant_q8k_kernel(const float* __restrict__ act, int lda, int m, int k, int8_t* __restrict__ qs,
                                                            float* __restrict__ ds) {
  pdl_launch_dependents();
  pdl_wait();
  const int nb = k / QKK;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= m * nb) return;
  const int row = warp / nb, i = warp % nb;
  const float* x = act + (size_t)row * lda + (size_t)i * QKK + lane * 8;
  float v[8];
  const float4 a = *(const float4*)x, b = *(const float4*)(x + 4);
  v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
  // first element (lowest index) with the strictly largest |x| decides `max` (sign included)
  float amax = 0.f, mx = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float ax = fabsf(v[j]);
    if (ax > amax) {
      amax = ax;
      mx = v[j];
    }
  }
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float oa = __shfl_xor_sync(0xffffffffu, amax, o), om = __shfl_xor_sync(0xffffffffu, mx, o);
    const bool other_is_lower = (lane ^ o) < lane;
    if (oa > amax || (oa == amax && other_is_lower)) {
      amax = oa;
      mx = om;
    }
  }
  int8_t* q = qs + (size_t)row * k + (size_t)i * QKK + lane * 8;
  if (amax == 0.f) {
    *(uint2*)q = make_uint2(0u, 0u);
    if (lane == 0) ds[row * nb + i] = 0.f;
    return;
  }
  const float iscale = __fdiv_rn(-128.f, mx);
  uint32_t w[2] = {0u, 0u};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int qv = __float2int_rn(__fmul_rn(iscale, v[j]));  // nearest_int: round half to even (quantize.h:801-807)
    qv = qv > 127 ? 127 : qv;
    w[j >> 2] |= (uint32_t)(qv & 0xff) << (8 * (j & 3));
  }
  *(uint2*)q = make_uint2(w[0], w[1]);
  if (lane == 0) ds[row * nb + i] = __fdiv_rn(1.f, iscale);
}

__device__ __forceinline__ int sext8(uint32_t w, int byte) { return (int)(int8_t)((w >> (8 * byte)) & 0xffu); }

// one warp per weight row; thread t owns fp32 lane L = t & 7 of super-blocks i = (t >> 3) + 4 s
template <int M>
__global__ void __launch_bounds__(kWarps * 32) gemv_q6k_kernel(const uint8_t* __restrict__ rows, int pitch, int sc_off, int d_off, int n,
                                                               int k, const int8_t* __restrict__ aq, const float* __restrict__ ad,
                                                               float* __restrict__ dst, int ldo, int m, const float* __restrict__ bias,
                                                               int bias_bcast, const float* __restrict__ residual) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int nb = k / QKK;
  int8_t* s_aq = (int8_t*)smem;                                             // [M][k]
  float* s_ad = (float*)(smem + (size_t)M * k);                             // [M][nb]
  int* s_sum = (int*)(smem + (size_t)M * k + (size_t)M * nb * 4);           // [kWarps][M][nb][8]
  pdl_launch_dependents();
  pdl_wait();
  for (int idx = threadIdx.x; idx < M * k / 16; idx += blockDim.x) {
    const int r = idx / (k / 16), c = idx % (k / 16);
    ((uint4*)s_aq)[idx] = r < m ? ((const uint4*)(aq + (size_t)r * k))[c] : make_uint4(0, 0, 0, 0);
  }
  for (int idx = threadIdx.x; idx < M * nb; idx += blockDim.x) s_ad[idx] = (idx / nb) < m ? ad[idx] : 0.f;
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * kWarps + warp;
  if (row >= n) return;
  const uint8_t* wr = rows + (size_t)row * pitch;
  const int L = lane & 7, ig = lane >> 3;
  int* my_sum = s_sum + (size_t)warp * M * nb * 8;

  for (int i = ig; i < nb; i += 4) {
    uint32_t qlw[2][2], qhw[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      qlw[j][0] = __ldg((const uint32_t*)(wr + i * 128 + 64 * j + 4 * L));
      qlw[j][1] = __ldg((const uint32_t*)(wr + i * 128 + 64 * j + 32 + 4 * L));
      qhw[j] = __ldg((const uint32_t*)(wr + nb * 128 + i * 64 + 32 * j + 4 * L));
    }
    const uint4 scw = __ldg((const uint4*)(wr + sc_off + i * 16));
    const uint32_t scv[4] = {scw.x, scw.y, scw.z, scw.w};
    const int hi16 = L >> 2;  // elements 4L.. fall in the second 16-group of the chunk when L >= 4
#pragma unroll
    for (int mm = 0; mm < M; ++mm) {
      int sumi = 0;
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t lo = (c >> 1) ? ((qlw[j][c & 1] >> 4) & 0x0F0F0F0Fu) : (qlw[j][c & 1] & 0x0F0F0F0Fu);
          const uint32_t hi = ((qhw[j] >> (2 * c)) & 0x03030303u) << 4;
          const int a = *(const int*)(s_aq + (size_t)mm * k + i * QKK + 128 * j + 32 * c + 4 * L);
          const int dot = dp4a_us(lo | hi, a, 0) - 32 * dp4a_us(0x01010101u, a, 0);
          const int sidx = 8 * j + 2 * c + hi16;
          sumi += sext8(scv[sidx >> 2], sidx & 3) * dot;
        }
      my_sum[(mm * nb + i) * 8 + L] = sumi;
    }
  }
  __syncwarp();
  // lanes 0..7: acc[L] = fma(d_i, (float)sumi[i][L], acc[L]) over the super-blocks in order (vec_dot.h:979)
#pragma unroll
  for (int mm = 0; mm < M; ++mm) {
    float acc = 0.f;
    if (lane < 8)
      for (int i = 0; i < nb; ++i) {
        const float d = __fmul_rn(s_ad[mm * nb + i], *(const float*)(wr + d_off + i * 4));
        acc = __fmaf_rn(d, (float)my_sum[(mm * nb + i) * 8 + lane], acc);
      }
    // hsum_float_8 (quantize.h:46-52): (a0+a4)+(a2+a6) + ((a1+a5)+(a3+a7))
    const float r = __fadd_rn(acc, __shfl_down_sync(0xffffffffu, acc, 4));   // lanes 0..3: a_l + a_{l+4}
    const float s2 = __fadd_rn(r, __shfl_down_sync(0xffffffffu, r, 2));      // lanes 0,1: r_l + r_{l+2}
    const float tot = __fadd_rn(s2, __shfl_down_sync(0xffffffffu, s2, 1));   // lane 0: s0 + s1
    if (lane == 0 && mm < m) {
      const size_t o = (size_t)mm * ldo + row;
      float v = tot;
      if (bias) v += bias_bcast ? bias[row] : bias[o];
      if (residual) v += residual[o];
      dst[o] = v;
    }
  }
}

__global__ void __launch_bounds__(256) dequant_q6k_kernel(const uint8_t* __restrict__ rows, int pitch, int sc_off, int d_off, int n, int k,
                                                          float* __restrict__ dst, int ld) {
  const int nb = k / QKK;
  const size_t total = (size_t)n * k;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(idx / k), e0 = (int)(idx % k);
    const int i = e0 / QKK, e = e0 % QKK;
    const uint8_t* wr = rows + (size_t)r * pitch;
    const int half = e >> 7, rr = e & 127, c = rr >> 5, l = rr & 31;
    const uint8_t lo = wr[i * 128 + 64 * half + 32 * (c & 1) + l];
    const int nib = (c >> 1) ? (lo >> 4) : (lo & 0xF);
    const int hi = (wr[nb * 128 + i * 64 + 32 * half + l] >> (2 * c)) & 3;
    const int q = (nib | (hi << 4)) - 32;
    const int sc = (int)(int8_t)wr[sc_off + i * 16 + 8 * half + 2 * c + (l >> 4)];
    const float d = *(const float*)(wr + d_off + i * 4);
    dst[(size_t)r * ld + e0] = __fmul_rn(__fmul_rn(d, (float)sc), (float)q);  // d * sc * q, left to right (quantize.h:973-976)
  }
}

}  // namespace

void ns_q6k_layout(ns_weight* w) {
  const int nb = w->k / QKK;
  w->kpad = w->k;
  w->group = 16;
  w->ngroups = w->k / 16;
  w->q_bytes = nb * 192;
  w->sc_off = nb * 192;
  w->zp_off = nb * 208;  // the f32 super-block scales live where other formats keep zero points
  w->pitch = (int)ns_round_up((size_t)nb * 212, 16);
}

int ns_launch_repack_q6k(const void* rows_dev, size_t nb01, ns_weight* w, cudaStream_t st) {
  const int nb = w->k / QKK;
  const size_t items = (size_t)w->n * nb * 53;
  const unsigned grid = (unsigned)((items + 255) / 256 > 65535u * 16 ? 65535u * 16 : (items + 255) / 256);
  repack_q6k_kernel<<<grid, 256, 0, st>>>((const uint8_t*)rows_dev, nb01, w->rows, w->n, nb, w->pitch, w->sc_off, w->zp_off);
  NS_CUDA_TRY(cudaGetLastError());
  return NS_OK;
}

size_t ns_q6k_workspace_bytes(int m, int k) { return ns_round_up((size_t)m * k, 16) + (size_t)m * (k / QKK) * 4; }

int ns_launch_dequant_q6k(const ns_weight* w, float* dst, int ld, cudaStream_t st) {
  const size_t total = (size_t)w->n * w->k;
  const unsigned grid = (unsigned)((total + 255) / 256 > 1u << 20 ? 1u << 20 : (total + 255) / 256);
  dequant_q6k_kernel<<<grid, 256, 0, st>>>(w->rows, w->pitch, w->sc_off, w->zp_off, w->n, w->k, dst, ld);
  NS_CUDA_TRY(cudaGetLastError());
  return NS_OK;
}

template <int M>
static int launch_q6k(const ns_weight* w, const int8_t* aq, const float* ad, float* dst, int ldo, int m, const float* bias,
                      int bias_bcast, const float* residual, cudaStream_t st) {
  auto kern = gemv_q6k_kernel<M>;
  const int nb = w->k / QKK;
  const size_t smem = (size_t)M * w->k + (size_t)M * nb * 4 + (size_t)kWarps * M * nb * 8 * 4;
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    if (smem > 200 * 1024) {
      ns_set_error("Q6_K GEMV: K=%d needs %zu bytes of shared memory", w->k, smem);
      return NS_E_UNSUPPORTED;
    }
    NS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  const unsigned grid = (unsigned)((w->n + kWarps - 1) / kWarps);
  NS_CUDA_TRY(ns_launch_pdl(kern, dim3(grid), dim3(kWarps * 32), smem, st, (const uint8_t*)w->rows, w->pitch, w->sc_off, w->zp_off,
                            w->n, w->k, aq, ad, dst, ldo, m, bias, bias_bcast, residual));
  ns_count_launch();
  return NS_OK;
}

// dst[m][n] (+bias, +residual) for up to 4 activation rows; ws = ns_q6k_workspace_bytes(m, k) bytes of device scratch
int ns_launch_mul_mat_q6k(const ns_weight* w, const float* act, int lda, float* dst, int ldo, int m, const float* bias,
                          int bias_bcast, const float* residual, void* ws, cudaStream_t st) {
  if (m < 1 || m > 4 || (lda & 3) || ((uintptr_t)act & 15)) {
    ns_set_error("Q6_K matmul: need 1..4 rows, lda %% 4 == 0 and 16-byte aligned activations");
    return NS_E_INVALID;
  }
  int8_t* aq = (int8_t*)ws;
  float* ad = (float*)((uint8_t*)ws + ns_round_up((size_t)m * w->k, 16));
  const int nb = w->k / QKK;
  const int warps = m * nb;
  NS_CUDA_TRY(ns_launch_pdl(act_quant_q8k_kernel, dim3((unsigned)((warps + 7) / 8)), dim3(256), 0, st, act, lda, m, w->k, aq, ad));
  ns_count_launch();
  if (m == 1) return launch_q6k<1>(w, aq, ad, dst, ldo, m, bias, bias_bcast, residual, st);
  if (m == 2) return launch_q6k<2>(w, aq, ad, dst, ldo, m, bias, bias_bcast, residual, st);
  return launch_q6k<4>(w, aq, ad, dst, ldo, m, bias, bias_bcast, residual, st);
}
