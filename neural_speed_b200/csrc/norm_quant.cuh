// norm_quant.cuh -- activation prologue of the decode GEMV with a fused RMSNorm (gemv_ring.cu): optional RMSNorm, then the reference's
// activation quantisation, written straight into the shared-memory image the dp4a loops read.
//   rms_norm * weight : models/llama/llama.cpp:205-210 (ne_rms_norm + ne_mul), arithmetic of kernel_ref.h:2199-2225 as rmsnorm_kernel
//                       (llama.cu): y = x * (1 / sqrt(sum(x^2)/n + eps)) * w
//   quantisation      : quantize_row_q8_0 (vectors/cpu/quantize.h:447, x86 body), quantize_fp_u8_colblock / _s8_colblock
//                       (bestla/bestla/kernel_ref.h:1825 / :1886) -- bit-exact codes, scales and zero points (same code as quant_smem.cuh)
// Called by NT threads (whole warps) that share named barrier `BAR`.  A thread owns 8 consecutive k; group/8 threads own a block.
#pragma once
#include "nsb.cuh"
#include "quant_smem.cuh"

namespace nsq {

struct NormQuantIn {
  const float* in;      // [M][lda] fp32 (global, read through L2: another SM may have just written it)
  const float* norm_w;  // [k] RMSNorm weight, or NULL: no normalisation
  float eps;
  int lda, k, kpad, group;
  int act_row, meta_off, meta_stride;
};

// The RMSNorm weights of the 8-groups thread `tid` owns in a single-pass row (kpad / 8 <= 3 * NT).  They are constants of the model:
// the caller fetches them BEFORE griddepcontrol.wait, so the L2 broadcast of the weight vector to every CTA overlaps the previous
// launch's tail instead of sitting between the wait and the first dp4a.
template <int NT>
__device__ __forceinline__ void prefetch_norm_w(const float* __restrict__ norm_w, int k, int kpad, int tid, float (&gw)[3][8]) {
  const int ngroups8 = kpad >> 3;
#pragma unroll
  for (int it = 0; it < 3; ++it) {
    const int e = it * NT + tid, k0 = e * 8;
    if (e < ngroups8 && k0 + 8 <= k) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(norm_w + k0)), w1 = __ldg(reinterpret_cast<const float4*>(norm_w + k0 + 4));
      gw[it][0] = w0.x; gw[it][1] = w0.y; gw[it][2] = w0.z; gw[it][3] = w0.w;
      gw[it][4] = w1.x; gw[it][5] = w1.y; gw[it][6] = w1.z; gw[it][7] = w1.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) gw[it][i] = (e < ngroups8 && k0 + i < k) ? __ldg(norm_w + k0 + i) : 0.f;
    }
  }
}

template <int BAR, int NT>
__device__ __forceinline__ void bar_sync() {
  asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(NT) : "memory");
}

// red: shared float[NT/32]
template <int COMP, int NT, int BAR>
__device__ __forceinline__ void norm_quantise_to_smem(const NormQuantIn& P, int M, uint32_t smem_base, float* red, int tid,
                                                      const float (*gw)[8] = nullptr) {  // gw: prefetch_norm_w's registers (single-pass rows)
  constexpr int NI = 3;  // load passes kept in registers (3 x 512 threads x 8 = 12288 elements)
  const int tpb = (COMP == NS_COMP_Q8_0 ? 32 : P.group) >> 3;
  const int ngroups8 = P.kpad >> 3;
  const bool single = ngroups8 <= NI * NT;
  const bool norm = P.norm_w != nullptr;
  for (int m = 0; m < M; ++m) {
    const float* row = P.in + (size_t)m * P.lda;
    const uint32_t img = smem_base + (uint32_t)m * P.act_row;
    const uint32_t meta = smem_base + P.meta_off + 8u * (uint32_t)(m * P.meta_stride);
    float vv[NI][8];
    float inv = 1.f;
    auto load8 = [&](int e, float* v) {
      const int k0 = e * 8;
      if (e < ngroups8 && k0 + 8 <= P.k) {
        const float4 x0 = ldcg4(row + k0), x1 = ldcg4(row + k0 + 4);
        v[0] = x0.x; v[1] = x0.y; v[2] = x0.z; v[3] = x0.w;
        v[4] = x1.x; v[5] = x1.y; v[6] = x1.z; v[7] = x1.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (e < ngroups8 && k0 + i < P.k) ? ldcg1(row + k0 + i) : 0.f;
      }
    };
    if (single) {
#pragma unroll
      for (int it = 0; it < NI; ++it) load8(it * NT + tid, vv[it]);
    }
    if (norm) {
      float ss = 0.f;
      if (single) {
#pragma unroll
        for (int it = 0; it < NI; ++it)
#pragma unroll
          for (int i = 0; i < 8; ++i) ss = fmaf(vv[it][i], vv[it][i], ss);
      } else {
        for (int eb = 0; eb < ngroups8; eb += NI * NT) {
#pragma unroll
          for (int it = 0; it < NI; ++it) load8(eb + it * NT + tid, vv[it]);
#pragma unroll
          for (int it = 0; it < NI; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i) ss = fmaf(vv[it][i], vv[it][i], ss);
        }
      }
      ss = warp_sum(ss);
      bar_sync<BAR, NT>();  // red[] free (previous row / previous use)
      if ((tid & 31) == 0) red[tid >> 5] = ss;
      bar_sync<BAR, NT>();
      float tot = 0.f;
#pragma unroll
      for (int i = 0; i < NT / 32; ++i) tot += red[i];
      inv = 1.f / sqrtf(tot / (float)P.k + P.eps);
    }
    for (int eb = 0; eb < ngroups8; eb += NI * NT) {
      if (!single) {
#pragma unroll
        for (int it = 0; it < NI; ++it) load8(eb + it * NT + tid, vv[it]);
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
        if (norm && gw != nullptr && single) {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = v[i] * inv * gw[it][i];  // (padding lanes: gw = 0, v = 0)
        } else if (norm) {
          if (live && k0 + 8 <= P.k) {
            const float4 w0 = __ldg((const float4*)(P.norm_w + k0)), w1 = __ldg((const float4*)(P.norm_w + k0 + 4));
            v[0] = v[0] * inv * w0.x; v[1] = v[1] * inv * w0.y; v[2] = v[2] * inv * w0.z; v[3] = v[3] * inv * w0.w;
            v[4] = v[4] * inv * w1.x; v[5] = v[5] * inv * w1.y; v[6] = v[6] * inv * w1.z; v[7] = v[7] * inv * w1.w;
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = (live && k0 + i < P.k) ? v[i] * inv * __ldg(P.norm_w + k0 + i) : 0.f;
          }
        }
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
        sa += __shfl_xor_sync(0xffffffffu, sa, 1);
        sa += __shfl_xor_sync(0xffffffffu, sa, 2);
        if (live) {
          const uint32_t alo = (q[0] & 0xff) | ((q[4] & 0xff) << 8) | ((q[1] & 0xff) << 16) | ((uint32_t)(q[5] & 0xff) << 24);
          const uint32_t ahi = (q[2] & 0xff) | ((q[6] & 0xff) << 8) | ((q[3] & 0xff) << 16) | ((uint32_t)(q[7] & 0xff) << 24);
          const int c = e >> 2, i = e & 3;
          sts64(img + (uint32_t)(c >> 5) * 1024u + (uint32_t)(i >> 1) * 512u + (uint32_t)(c & 31) * 16u + (uint32_t)(i & 1) * 8u, alo, ahi);
          if (i == 0) sts64(meta + 8u * (uint32_t)c, __float_as_uint(scale), (uint32_t)((sa & 0xffff) | (za << 16)));
        }
      }
    }
  }
}

}  // namespace nsq
