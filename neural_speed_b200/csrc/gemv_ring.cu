// gemv_ring.cu -- the decode GEMV for 4-bit integer weights with 8-bit integer activations (ggml Q4_0 x Q8_0 and BesTLA
// int4 CompInt8): HBM -> shared-memory ring via TMA bulk copies, dp4a out of shared memory.
//
// Replaces the same reference functions as gemv.cu (ne_vec_dot_q4_0_q8_0, core/layers/vec_dot.h:131; gemv_4bit_u8s8_fp32 /
// gemv_4bit_s8s8_fp32, bestla/bestla/kernel_ref.h:2372/2432) for M <= 4.
//
// Why a ring: at 6.6 TB/s each of the 148 SMs must keep >= ~45-90 KB of loads in flight (Little's law, ~1-2 us loaded
// HBM latency); register-staged loads cap that at occupancy x 128 B per thread.  Here one elected producer thread per CTA
// issues cp.async.bulk (UBLKCP) copies of whole weight-row PAIRS -- a row of the NSB layout is one contiguous
// [nibbles | scales | zero-points] byte range -- into a ring of `stages` slots guarded by full/empty mbarriers; ~100 KB
// per CTA, 2 CTAs per SM, stay in flight regardless of what the consumer warps are doing.  Each of the 8 consumer warps
// owns a whole stage at a time (two rows), so there is no cross-warp reduction and rows are dealt round-robin over
// CTAs (perfect balance at any N).  The producer starts streaming BEFORE griddepcontrol.wait: under programmatic
// dependent launch the next kernel's ring fills while the previous kernel drains.
// Roofline: HBM.  Algorithmic bytes per launch = sum over weights of N*K/2 + N*ceil(K/g)*(scale_bytes [+1 if asym]).
#include "nsb.cuh"

namespace {

constexpr int kConsumers = 8;
constexpr int kThreads = (kConsumers + 1) * 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
// TMA 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct PairSrc {
  const uint8_t* r0;
  const uint8_t* r1;
  long long out0, out1;
  bool valid1;
};

__device__ __forceinline__ PairSrc resolve_pair(const GemvParams& P, int p) {
  PairSrc s;
  if (P.mode == NS_GEMV_GATE_UP_SILU) {
    s.r0 = P.rows[0] + (size_t)p * P.pitch;
    s.r1 = P.rows[1] + (size_t)p * P.pitch;
    s.out0 = s.out1 = p;
    s.valid1 = true;
    return s;
  }
  int row = 2 * p, wi = 0;
  if (P.nw > 1 && row >= P.n[0]) {
    row -= P.n[0];
    wi = 1;
    if (P.nw > 2 && row >= P.n[1]) {
      row -= P.n[1];
      wi = 2;
    }
  }
  // every weight but the last has an even n (checked by the launcher), so a pair never straddles two weights
  s.valid1 = row + 1 < P.n[wi];
  s.r0 = P.rows[wi] + (size_t)row * P.pitch;
  s.r1 = s.valid1 ? s.r0 + P.pitch : s.r0;
  s.out0 = P.dst_off[wi] + row;
  s.out1 = s.out0 + 1;
  return s;
}

template <int AMODE, int M, bool ASYM>
__global__ void __launch_bounds__(kThreads, 2) gemv_ring_kernel(const GemvParams P, int ring_off, int stages) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int stage_bytes = 2 * P.pitch;
  unsigned char* ring = smem + ring_off;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)stages * stage_bytes);
  uint64_t* empty = full + stages;

  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int first = blockIdx.x;
  const int my_units = first < P.npairs ? (P.npairs - first + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == kConsumers) {
    // ===================== producer: stream whole row pairs, never touches activations =====================
    if (lane == 0) {
      int s = 0;
      uint32_t phase = 0;
      for (int j = 0; j < my_units; ++j) {
        if (j >= stages) mbar_wait(&empty[s], phase ^ 1);
        const PairSrc ps = resolve_pair(P, first + j * (int)gridDim.x);
        unsigned char* dst = ring + (size_t)s * stage_bytes;
        mbar_expect_tx(&full[s], (uint32_t)stage_bytes);
        bulk_g2s(dst, ps.r0, (uint32_t)P.pitch, &full[s]);
        bulk_g2s(dst + P.pitch, ps.r1, (uint32_t)P.pitch, &full[s]);
        if (++s == stages) {
          s = 0;
          phase ^= 1;
        }
      }
    }
    return;
  }

  // ===================== consumers =====================
  pdl_wait();  // activations (and residual) come from earlier kernels
  {
    const uint4* src = reinterpret_cast<const uint4*>(P.act);
    uint4* dstv = reinterpret_cast<uint4*>(smem);
    const int nvec = P.act_bytes >> 4;
    for (int i = threadIdx.x; i < nvec; i += kConsumers * 32) dstv[i] = src[i];
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kConsumers * 32) : "memory");
  const int2* meta_s = reinterpret_cast<const int2*>(smem + P.meta_off);
  const int nchunks = P.kpad >> 5;

  for (int j = warp; j < my_units; j += kConsumers) {
    const int s = j % stages;
    const uint32_t phase = (uint32_t)(j / stages) & 1u;
    const PairSrc ps = resolve_pair(P, first + j * (int)gridDim.x);
    mbar_wait(&full[s], phase);
    const unsigned char* r0 = ring + (size_t)s * stage_bytes;
    const unsigned char* r1 = r0 + P.pitch;

    float acc[2][M];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int m = 0; m < M; ++m) acc[r][m] = 0.f;

#pragma unroll 2
    for (int c = lane; c < nchunks; c += 32) {
      const uint4 wv[2] = {reinterpret_cast<const uint4*>(r0)[c], reinterpret_cast<const uint4*>(r1)[c]};
      const int gi = (P.cpg == 1) ? c : c / P.cpg;
      const float ws[2] = {ns_scale_at(r0 + P.sc_off, P.stype, gi), ns_scale_at(r1 + P.sc_off, P.stype, gi)};
      int off[2] = {8, 8};
      if (ASYM) {
        off[0] += (int)(signed char)r0[P.zp_off + gi];
        off[1] += (int)(signed char)r1[P.zp_off + gi];
      }
      uint32_t lo[2][4], hi[2][4];
      int su[2] = {0, 0};
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const uint32_t ww[4] = {wv[r].x, wv[r].y, wv[r].z, wv[r].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          lo[r][i] = ww[i] & 0x0F0F0F0Fu;
          hi[r][i] = (ww[i] >> 4) & 0x0F0F0F0Fu;
          if (AMODE == A_U8) {  // sum of the weight codes, needed for the activation zero point
            su[r] = dp4a_ss(0x01010101, (int)lo[r][i], su[r]);
            su[r] = dp4a_ss(0x01010101, (int)hi[r][i], su[r]);
          }
        }
      }
#pragma unroll
      for (int m = 0; m < M; ++m) {
        const uint4* ap = reinterpret_cast<const uint4*>(smem + (size_t)m * P.kpad) + 2 * c;
        const uint4 a0 = ap[0], a1 = ap[1];
        const int2 mt = meta_s[m * P.meta_stride + c];
        const float a_scale = __int_as_float(mt.x);
        const int sa = (int)(short)(mt.y & 0xffff);
        const int za = (mt.y >> 16) & 0xff;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          int ps_ = 0;
          // NSB4: word i pairs with activation words (Alo_i, Ahi_i) = ((a0,a4,a1,a5),(a2,a6,a3,a7)) of 8-group i
          if (AMODE == A_U8) {
            ps_ = dp4a_uu(a0.x, lo[r][0], ps_); ps_ = dp4a_uu(a0.y, hi[r][0], ps_);
            ps_ = dp4a_uu(a0.z, lo[r][1], ps_); ps_ = dp4a_uu(a0.w, hi[r][1], ps_);
            ps_ = dp4a_uu(a1.x, lo[r][2], ps_); ps_ = dp4a_uu(a1.y, hi[r][2], ps_);
            ps_ = dp4a_uu(a1.z, lo[r][3], ps_); ps_ = dp4a_uu(a1.w, hi[r][3], ps_);
          } else {
            ps_ = dp4a_ss((int)a0.x, (int)lo[r][0], ps_); ps_ = dp4a_ss((int)a0.y, (int)hi[r][0], ps_);
            ps_ = dp4a_ss((int)a0.z, (int)lo[r][1], ps_); ps_ = dp4a_ss((int)a0.w, (int)hi[r][1], ps_);
            ps_ = dp4a_ss((int)a1.x, (int)lo[r][2], ps_); ps_ = dp4a_ss((int)a1.y, (int)hi[r][2], ps_);
            ps_ = dp4a_ss((int)a1.z, (int)lo[r][3], ps_); ps_ = dp4a_ss((int)a1.w, (int)hi[r][3], ps_);
          }
          // sum (a - za)(u - off) = sum a*u - off*Sa - za*(Su - 32*off): one exact integer per 32-element chunk
          int isum = ps_ - off[r] * sa;
          if (AMODE == A_U8) isum -= za * (su[r] - 32 * off[r]);
          acc[r][m] = fmaf((float)isum, a_scale * ws[r], acc[r][m]);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);  // slot may be refilled

#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int m = 0; m < M; ++m) acc[r][m] = warp_sum(acc[r][m]);
    if (lane == 0) {
      if (P.mode == NS_GEMV_GATE_UP_SILU) {
#pragma unroll
        for (int m = 0; m < M; ++m) {
          if (m < P.m) {
            const float g = acc[0][m], up = acc[1][m];
            const float sg = g / (1.f + expf(-g));  // swish alpha=-1 (kernel_ref.h:1574)
            if (P.aux) P.aux[(size_t)m * P.ldo + ps.out0] = sg;
            P.dst[(size_t)m * P.ldo + ps.out0] = sg * up;
          }
        }
      } else {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          if (r == 1 && !ps.valid1) continue;
          const long long out = r ? ps.out1 : ps.out0;
#pragma unroll
          for (int m = 0; m < M; ++m) {
            if (m < P.m) {
              const size_t o = (size_t)m * P.ldo + out;
              float v = acc[r][m];
              if (P.bias) v += P.bias_bcast ? P.bias[out] : P.bias[o];
              if (P.residual) v += P.residual[o];
              P.dst[o] = v;
            }
          }
        }
      }
    }
  }
}

template <int AMODE, int M, bool ASYM>
int launch_one(const GemvParams& P, int mt, cudaStream_t st) {
  auto kern = gemv_ring_kernel<AMODE, M, ASYM>;
  static bool attr_set = false;
  if (!attr_set) {
    NS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  const int stage_bytes = 2 * P.pitch;
  const size_t act_region = ns_round_up(ns_round_up((size_t)mt * P.kpad, 16) + (size_t)mt * P.meta_stride * 8, 128);
  const size_t budget = 110 * 1024;  // two CTAs per SM
  int stages = 2;
  int ctas_per_sm = 2;
  if (act_region + 2 * (size_t)stage_bytes + 64 <= budget) {
    stages = (int)((budget - act_region - 64) / (stage_bytes + 16));
    if (stages > 64) stages = 64;
  } else {
    // very long rows: one CTA per SM with whatever ring fits
    ctas_per_sm = 1;
    stages = (int)((200 * 1024 - act_region - 64) / (stage_bytes + 16));
    if (stages < 1) {
      ns_set_error("gemv_ring: row pitch %d too large for shared memory", P.pitch);
      return NS_E_UNSUPPORTED;
    }
    if (stages > 16) stages = 16;
  }
  const size_t smem = act_region + (size_t)stages * stage_bytes + (size_t)stages * 16;
  int grid = ns_num_sms() * ctas_per_sm;
  if (grid > P.npairs) grid = P.npairs;
  if (grid < 1) grid = 1;
  NS_CUDA_TRY(ns_launch_pdl(kern, dim3(grid), dim3(kThreads), smem, st, P, (int)act_region, stages));
  ns_count_launch();
  return NS_OK;
}

template <int AMODE, bool ASYM>
int launch_m(const GemvParams& P, int mt, cudaStream_t st) {
  switch (mt) {
    case 1: return launch_one<AMODE, 1, ASYM>(P, mt, st);
    case 2: return launch_one<AMODE, 2, ASYM>(P, mt, st);
    default: return launch_one<AMODE, 4, ASYM>(P, mt, st);
  }
}

}  // namespace

int ns_launch_gemv_ring(const GemvParams& P, int amode, bool asym, int mt, cudaStream_t st) {
  if (amode == A_U8) return asym ? launch_m<A_U8, true>(P, mt, st) : launch_m<A_U8, false>(P, mt, st);
  return asym ? launch_m<A_S8, true>(P, mt, st) : launch_m<A_S8, false>(P, mt, st);
}
