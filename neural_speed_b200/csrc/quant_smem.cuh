// quant_smem.cuh -- in-kernel activation quantisation into the shared-memory image the decode GEMV kernels consume.
// Same arithmetic as act_quant_kernel<COMP> (act_prep.cu), i.e. quantize_row_q8_0 (vectors/cpu/quantize.h:447, x86 body),
// quantize_fp_u8_colblock / quantize_fp_s8_colblock (bestla/bestla/kernel_ref.h:1825 / :1886): bit-exact codes, scales and
// zero points.  Called by all consumer threads of a CTA (NT threads, whole warps).
#pragma once
#include "nsb.cuh"

struct QuantIn {
  const float* in;  // [M][lda] fp32, global (read through L2: another SM may have just written it)
  int lda, k, kpad, group;
  int act_row, meta_off, meta_stride;  // image geometry (bytes, bytes, int2 units)
};

namespace nsq {
__device__ __forceinline__ void sts64(uint32_t a, uint32_t x, uint32_t y) {
  asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ float4 ldcg4(const float* p) {
  float4 r;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float ldcg1(const float* p) {
  float r;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
// utils::cast<float,uint8_t> / <float,int8_t> (bestla_utils.h:507-521) with the x86 NaN->0 behaviour (see act_prep.cu)
__device__ __forceinline__ int cast_u8(float x) {
  if (x != x) return 0;
  x += 0.5f;
  x = fminf(x, 255.f);
  x = fmaxf(x, 0.f);
  return (int)x;
}
__device__ __forceinline__ int cast_s8(float x) {
  if (x != x) return 0;
  x = roundf(x);
  x = fminf(x, 127.f);
  x = fmaxf(x, -128.f);
  return (int)x;
}

// Quantise the op's activations [M][K] (fp32, global, read through L2) into the shared-memory image gemv_ring.cu expects.
// One thread owns one 8-group (8 consecutive k); TPB = group/8 consecutive threads own one quantisation block.
// Arithmetic identical to act_quant_kernel<COMP> (act_prep.cu): bit-exact codes, scales and zero points.
template <int COMP, int NT>
__device__ __forceinline__ void quantise_to_smem(const QuantIn& P, int M, uint32_t smem_base) {
  const int tpb = (COMP == NS_COMP_Q8_0 ? 32 : P.group) >> 3;  // threads per quantisation block (4..32, power of two)
  const int ngroups8 = P.kpad >> 3;
  const int tid = threadIdx.x;
  for (int m = 0; m < M; ++m) {
    const float* row = P.in + (size_t)m * P.lda;
    const uint32_t img = smem_base + (uint32_t)m * P.act_row;
    const uint32_t meta = smem_base + P.meta_off + 8u * (uint32_t)(m * P.meta_stride);
    // loads of NI consecutive passes are issued back to back (L2 latency ~0.5 us each would otherwise serialise)
    constexpr int NI = 3;
    for (int eb = 0; eb < ngroups8; eb += NI * NT) {
     float vv[NI][8];
#pragma unroll
     for (int it = 0; it < NI; ++it) {
      const int e = eb + it * NT + tid;
      const int k0 = e * 8;
      if (e < ngroups8 && k0 + 8 <= P.k) {
        const float4 x0 = ldcg4(row + k0), x1 = ldcg4(row + k0 + 4);
        vv[it][0] = x0.x; vv[it][1] = x0.y; vv[it][2] = x0.z; vv[it][3] = x0.w;
        vv[it][4] = x1.x; vv[it][5] = x1.y; vv[it][6] = x1.z; vv[it][7] = x1.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) vv[it][i] = (e < ngroups8 && k0 + i < P.k) ? ldcg1(row + k0 + i) : 0.f;
      }
     }
#pragma unroll
     for (int it = 0; it < NI; ++it) {
      const int e0 = eb + it * NT;
      if (e0 >= ngroups8) break;  // uniform across the CTA
      const int e = e0 + tid;
      const bool live = e < ngroups8;
      const int k0 = e * 8;
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = vv[it][i];
      // block range (all lanes of the warp take part in the shuffles)
      float vmax = (COMP == NS_COMP_Q8_0) ? 0.f : 1.17549435e-38f, vmin = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (COMP == NS_COMP_INT8) {
          vmax = fmaxf(v[i], vmax);
          vmin = fminf(v[i], vmin);
        } else {
          vmax = fmaxf(vmax, fabsf(v[i]));
        }
      }
      for (int o = 1; o < tpb; o <<= 1) {
        vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        if (COMP == NS_COMP_INT8) vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
      }
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
      int q[8], sa = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (k0 + i < P.k) {
          if (COMP == NS_COMP_Q8_0) q[i] = __float2int_rn(v[i] * rscale);
          else if (COMP == NS_COMP_INT8) q[i] = cast_u8((float)za + (float)(int)roundf(v[i] * rscale));
          else q[i] = cast_s8(v[i] * rscale);
        } else {
          q[i] = za;  // padding contributes (a - za) == 0
        }
        sa += q[i];
      }
      // chunk sum over the 4 threads of a 32-element chunk
      sa += __shfl_xor_sync(0xffffffffu, sa, 1);
      sa += __shfl_xor_sync(0xffffffffu, sa, 2);
      if (live) {
        // bytes in dp4a order: Alo = (a0,a4,a1,a5), Ahi = (a2,a6,a3,a7)
        const uint32_t alo = (q[0] & 0xff) | ((q[4] & 0xff) << 8) | ((q[1] & 0xff) << 16) | ((uint32_t)(q[5] & 0xff) << 24);
        const uint32_t ahi = (q[2] & 0xff) | ((q[6] & 0xff) << 8) | ((q[3] & 0xff) << 16) | ((uint32_t)(q[7] & 0xff) << 24);
        const int c = e >> 2, i = e & 3;
        sts64(img + (uint32_t)(c >> 5) * 1024u + (uint32_t)(i >> 1) * 512u + (uint32_t)(c & 31) * 16u + (uint32_t)(i & 1) * 8u, alo, ahi);
        if (i == 0) sts64(meta + 8u * (uint32_t)c, __float_as_uint(scale), (uint32_t)((sa & 0xffff) | (za << 16)));
      }
     }  // it
    }
  }
}

}  // namespace nsq
