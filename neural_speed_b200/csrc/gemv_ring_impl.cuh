// gemv_ring_impl.cuh -- the ring GEMV kernel template, its shared-memory planner and launcher.  Included by gemv_ring.cu (two CTAs
// per SM, 7 consumer warps) and gemv_ring_wide.cu (one CTA per SM, 14 consumer warps): two translation units so that the ~170
// instantiations compile in parallel.  See gemv_ring.cu for the design notes.
#pragma once
#include "nsb.cuh"
#include "quant_smem.cuh"
#include "norm_quant.cuh"

namespace {

constexpr int kConsumers = 7;
constexpr int kThreads = (kConsumers + 1) * 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
// TMA 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
  return r;
}
__device__ __forceinline__ uint2 lds64(uint32_t a) {
  uint2 r;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(a));
  return r;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
  uint32_t r;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(a));
  return r;
}
__device__ __forceinline__ uint32_t lds16(uint32_t a) {
  unsigned short r;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(r) : "r"(a));
  return r;
}
__device__ __forceinline__ int lds8s(uint32_t a) {
  int r;
  asm volatile("ld.shared.s8 %0, [%1];" : "=r"(r) : "r"(a));
  return r;
}
template <int STYPE>
__device__ __forceinline__ float lds_scale(uint32_t base, int idx) {
  if (STYPE == NS_S_F32) return __uint_as_float(lds32(base + 4 * idx));
  if (STYPE == NS_S_F16) return __half2float(__ushort_as_half((unsigned short)lds16(base + 2 * idx)));
  return __uint_as_float(lds16(base + 2 * idx) << 16);
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

struct RingCfg {
  int ring_off;     // byte offset of the ring inside dynamic shared memory
  int stages;
  int units;        // row pairs (ROWS == 2) or rows (ROWS == 1) of this launch
  int active;       // consumer warps that own ring stages (<= kConsumers; `stages` is a multiple of it)
  int act_row;      // bytes per activation row in the staged image
  int red_off;      // byte offset of the RMSNorm reduction scratch (8 floats) inside the activation region, fused-norm launches only
  uint32_t cpg_magic;  // ceil(2^32 / cpg): gi = umulhi(c, magic)
};

// ROWS == 1: unit u is row u of the concatenated weights (long rows: a pair would leave one ring stage per consumer warp)
__device__ __forceinline__ PairSrc resolve_single(const GemvParams& P, int u) {
  PairSrc s;
  int row = u, wi = 0;
  if (P.nw > 1 && row >= P.n[0]) {
    row -= P.n[0];
    wi = 1;
    if (P.nw > 2 && row >= P.n[1]) {
      row -= P.n[1];
      wi = 2;
    }
  }
  s.r0 = s.r1 = P.rows[wi] + (size_t)row * P.pitch;
  s.out0 = s.out1 = P.dst_off[wi] + row;
  s.valid1 = false;
  return s;
}

// NORM: the fused-RMSNorm prologue is a separate instantiation, so the plain kernels (the headline path) keep the code and the
// register allocation they had without it
// NC consumer warps: 7 (+1 producer warp, two CTAs per SM) or 14 (+2 producer warps, ONE CTA per SM: the activation row is pulled
// through L2 and quantised once per SM instead of twice -- the 16-44 KB broadcast to every CTA is what separates the fused launch
// list from the pre-quantised one; the two producer warps take alternate ring stages)
template <int AMODE, int M, bool ASYM, int STYPE, int ROWS, bool NORM, int NC>
__global__ void __launch_bounds__((NC + (NC > kConsumers ? 2 : 1)) * 32, NC > kConsumers ? 1 : 2)
    gemv_ring_kernel(const GemvParams P, const RingCfg R) {
  constexpr int NP = NC > kConsumers ? 2 : 1;  // producer warps
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int stage_bytes = ROWS * P.pitch;
  const int stages = R.stages;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t ring = smem_base + R.ring_off;
  const uint32_t full0 = ring + (uint32_t)stages * stage_bytes;
  const uint32_t empty0 = full0 + 8u * stages;

  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int first = blockIdx.x;
  const int gstride = (int)gridDim.x;
  const int my_units = first < R.units ? (R.units - first + gstride - 1) / gstride : 0;

  if (warp >= NC) {
    // ===================== producer(s): stream whole row pairs, never touch activations =====================
    // producer warp pw takes units pw, pw + NP, ...; unit j always lands in stage j % stages (`stages` is a multiple of NC: even)
    if (lane == 0) {
      int s = warp - NC;
      uint32_t phase = 0;
      for (int j = warp - NC; j < my_units; j += NP) {
        if (j >= stages) mbar_wait(empty0 + 8 * s, phase ^ 1);
        const PairSrc ps = ROWS == 2 ? resolve_pair(P, first + j * gstride) : resolve_single(P, first + j * gstride);
        const uint32_t dst = ring + (uint32_t)s * stage_bytes;
        mbar_expect_tx(full0 + 8 * s, (uint32_t)stage_bytes);
        bulk_g2s(dst, ps.r0, (uint32_t)P.pitch, full0 + 8 * s);
        if (ROWS == 2) bulk_g2s(dst + P.pitch, ps.r1, (uint32_t)P.pitch, full0 + 8 * s);
        s += NP;
        if (s >= stages) {
          s -= stages;
          phase ^= 1;
        }
      }
    }
    return;
  }

  // ===================== consumers =====================
  // fused RMSNorm: the norm weights are model constants -- fetch this thread's share before the wait (single-pass rows)
  float gw[NORM ? 3 : 1][8];
  bool gw_pref = false;
  if constexpr (NORM) {
    gw_pref = P.norm_w != nullptr && (P.kpad >> 3) <= 3 * NC * 32;
    if (gw_pref) nsq::prefetch_norm_w<NC * 32>(P.norm_w, P.k, P.kpad, (int)threadIdx.x, gw);
  }
  pdl_wait();  // activations (and residual) come from earlier kernels
  bool normed = false;
  if constexpr (NORM) {
    if (P.norm_w) {  // (a norm-capable image also serves plain nodes: see GemvParams::one_image)
      // fused ne_rms_norm + ne_mul + NE_TASK_INIT: every CTA already reads the whole fp32 row, the sum of squares costs one more
      // block reduction instead of a kernel boundary (llama.cpp:205-210; arithmetic of rmsnorm_kernel, llama.cu)
      const nsq::NormQuantIn ni{P.act_f32, P.norm_w, P.norm_eps, P.lda, P.k, P.kpad, P.comp == NS_COMP_Q8_0 ? 32 : P.group,
                                R.act_row, P.meta_off, P.meta_stride};
      float* red = reinterpret_cast<float*>(smem + R.red_off);
      if (AMODE == A_U8)
        nsq::norm_quantise_to_smem<NS_COMP_INT8, NC * 32, 1>(ni, P.m, smem_base, red, (int)threadIdx.x, gw_pref ? gw : nullptr);
      else if (P.comp == NS_COMP_Q8_0)
        nsq::norm_quantise_to_smem<NS_COMP_Q8_0, NC * 32, 1>(ni, P.m, smem_base, red, (int)threadIdx.x, gw_pref ? gw : nullptr);
      else
        nsq::norm_quantise_to_smem<NS_COMP_INT8_S8, NC * 32, 1>(ni, P.m, smem_base, red, (int)threadIdx.x, gw_pref ? gw : nullptr);
      normed = true;
    }
  }
  if (normed) {
  } else if (P.act_f32) {
    // fused NE_TASK_INIT: quantise the fp32 rows straight into the shared-memory image (no separate kernel, no round trip)
    const QuantIn qi{P.act_f32, P.lda, P.k, P.kpad, P.comp == NS_COMP_Q8_0 ? 32 : P.group, R.act_row, P.meta_off, P.meta_stride};
    if (P.comp == NS_COMP_Q8_0) nsq::quantise_to_smem<NS_COMP_Q8_0, NC * 32>(qi, P.m, smem_base);
    else if (P.comp == NS_COMP_INT8) nsq::quantise_to_smem<NS_COMP_INT8, NC * 32>(qi, P.m, smem_base);
    else nsq::quantise_to_smem<NS_COMP_INT8_S8, NC * 32>(qi, P.m, smem_base);
  } else {
    const uint4* src = reinterpret_cast<const uint4*>(P.act);
    uint4* dstv = reinterpret_cast<uint4*>(smem);
    const int nvec = P.act_bytes >> 4;
    for (int i = threadIdx.x; i < nvec; i += NC * 32) dstv[i] = src[i];
  }
  asm volatile("bar.sync 1, %0;" ::"n"(NC * 32) : "memory");
  const uint32_t meta_s = smem_base + P.meta_off;
  const int nchunks = P.kpad >> 5;

  // This warp visits units warp, warp+active, ... .  `stages` is a multiple of `active` (launcher), so stage s is only ever
  // consumed by warp s % active: every mbarrier is waited on by ONE warp that observes all of its phases in order.
  // (A parity wait issued a whole phase early returns true immediately -- waiters must never run ahead of a barrier.)
  int s = warp;
  uint32_t phase = 0;

  const int active = R.active;  // long rows leave room for fewer stages than consumer warps: the surplus warps only helped quantise
  if (warp >= active) return;
  for (int j = warp; j < my_units; j += active) {
    const PairSrc ps = ROWS == 2 ? resolve_pair(P, first + j * gstride) : resolve_single(P, first + j * gstride);
    mbar_wait(full0 + 8 * s, phase);
    const uint32_t r0 = ring + (uint32_t)s * stage_bytes;
    const uint32_t r1 = ROWS == 2 ? r0 + P.pitch : r0;
    const uint32_t rb[2] = {r0, r1};

    float acc[ROWS][M];
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
      for (int m = 0; m < M; ++m) acc[r][m] = 0.f;

#pragma unroll 2
    for (int c = lane; c < nchunks; c += 32) {
      const int gi = (P.cpg == 1) ? c : (int)__umulhi((uint32_t)c, R.cpg_magic);
      uint4 wv[ROWS];
      float ws[ROWS];
      int off[ROWS];
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        wv[r] = lds128(rb[r] + 16 * c);
        ws[r] = lds_scale<STYPE>(rb[r] + P.sc_off, gi);
        off[r] = 8;
        if (ASYM) off[r] += lds8s(rb[r] + P.zp_off + gi);
      }
      // low nibbles as bytes, high nibbles as bytes * 16 (no shift): exact, divided out after the dot
      uint32_t lo[ROWS][4], hi[ROWS][4];
      int su[ROWS];
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        su[r] = 0;
        const uint32_t ww[4] = {wv[r].x, wv[r].y, wv[r].z, wv[r].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          lo[r][i] = ww[i] & 0x0F0F0F0Fu;
          hi[r][i] = ww[i] & 0xF0F0F0F0u;
        }
        if (AMODE == A_U8) {  // sum of the weight codes, needed for the activation zero point
          int sl = 0, sh = 0;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            sl = dp4a_uu(lo[r][i], 0x01010101u, sl);
            sh = dp4a_uu(hi[r][i], 0x01010101u, sh);
          }
          su[r] = sl + (sh >> 4);
        }
      }
      // activation image: per 32-chunk super-block the first 16 B of every chunk, then the second 16 B (conflict-free)
      const uint32_t a_off = (uint32_t)(c >> 5) * 1024u + (uint32_t)(c & 31) * 16u;
#pragma unroll
      for (int m = 0; m < M; ++m) {
        const uint32_t ab = smem_base + (uint32_t)m * R.act_row + a_off;
        const uint4 a0 = lds128(ab), a1 = lds128(ab + 512);
        const uint2 mt = lds64(meta_s + 8u * (uint32_t)(m * P.meta_stride + c));
        const float a_scale = __uint_as_float(mt.x);
        const int sa = (int)(short)(mt.y & 0xffff);
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
          int pl = 0, ph = 0;
          // NSB4: word i pairs with activation words (Alo_i, Ahi_i) = ((a0,a4,a1,a5),(a2,a6,a3,a7)) of 8-group i
          if (AMODE == A_U8) {
            pl = dp4a_uu(a0.x, lo[r][0], pl); ph = dp4a_uu(a0.y, hi[r][0], ph);
            pl = dp4a_uu(a0.z, lo[r][1], pl); ph = dp4a_uu(a0.w, hi[r][1], ph);
            pl = dp4a_uu(a1.x, lo[r][2], pl); ph = dp4a_uu(a1.y, hi[r][2], ph);
            pl = dp4a_uu(a1.z, lo[r][3], pl); ph = dp4a_uu(a1.w, hi[r][3], ph);
          } else {  // signed activations x unsigned weight bytes
            pl = dp4a_us(lo[r][0], (int)a0.x, pl); ph = dp4a_us(hi[r][0], (int)a0.y, ph);
            pl = dp4a_us(lo[r][1], (int)a0.z, pl); ph = dp4a_us(hi[r][1], (int)a0.w, ph);
            pl = dp4a_us(lo[r][2], (int)a1.x, pl); ph = dp4a_us(hi[r][2], (int)a1.y, ph);
            pl = dp4a_us(lo[r][3], (int)a1.z, pl); ph = dp4a_us(hi[r][3], (int)a1.w, ph);
          }
          // sum (a - za)(u - off) = sum a*u - off*Sa - za*(Su - 32*off): one exact integer per 32-element chunk
          int isum = pl + (ph >> 4) - off[r] * sa;  // ph is an exact multiple of 16
          if (AMODE == A_U8) {
            const int za = (int)((mt.y >> 16) & 0xff);
            isum -= za * (su[r] - 32 * off[r]);
          }
          acc[r][m] = fmaf((float)isum, a_scale * ws[r], acc[r][m]);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty0 + 8 * s);  // slot may be refilled
    s += active;
    if (s >= stages) {
      s -= stages;
      phase ^= 1u;
    }

#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
      for (int m = 0; m < M; ++m) acc[r][m] = warp_sum(acc[r][m]);
    if (lane == 0) {
      if (ROWS == 2 && P.mode == NS_GEMV_GATE_UP_SILU) {
#pragma unroll
        for (int m = 0; m < M; ++m) {
          if (m < P.m) {
            const float g = acc[0][m], up = acc[ROWS - 1][m];
            const float sg = P.eltop == NS_ELT_GELU ? ns_gelu(g) : ns_silu(g);  // kernel_ref.h:1569-1576
            if (P.aux) P.aux[(size_t)m * P.ldo + ps.out0] = sg;
            P.dst[(size_t)m * P.ldo + ps.out0] = sg * up;
          }
        }
      } else {
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
          if (r == 1 && !ps.valid1) continue;
          const long long out = r ? ps.out1 : ps.out0;
#pragma unroll
          for (int m = 0; m < M; ++m) {
            if (m < P.m) {
              const size_t o = (size_t)m * P.ldo + out;
              float v = acc[r][m];
              if (P.bias) v += P.bias_bcast ? P.bias[out] : P.bias[o];
              if (P.eltop == NS_ELT_GELU) v = ns_gelu(v);
              if (P.residual) v += P.residual[o];
              P.dst[o] = v;
            }
          }
        }
      }
    }
  }
}

struct RingPlan {
  int rows, stages, active, ctas;
  size_t budget;
  double score;
};

// Shared-memory plan.  Candidates: row pairs or single rows per stage, half an SM (two CTAs per SM) or a whole SM.  The score
// is the number of consumer warps per SM that own a stage; pairs share the activation loads between two rows (measured:
// K = 11008 with 7 pair stages beats 14 single-row stages, 950 vs 937 tok/s), so single rows are only taken when pairs would
// leave consumer warps without a stage (K >= ~14000).  Ties go to the deeper ring.
static RingPlan plan_ring(const GemvParams& P, size_t act_region, bool wide) {
  static const int env_budget = getenv("NS_RING_BUDGET_KB") ? atoi(getenv("NS_RING_BUDGET_KB")) : 0;  // tuning aids
  static const int env_rows = getenv("NS_RING_ROWS") ? atoi(getenv("NS_RING_ROWS")) : 0;
  const size_t budgets[2] = {(size_t)(env_budget > 0 ? env_budget : 113) * 1024, 200 * 1024};
  RingPlan best = {0, 0, 0, 0, 0, -1.0};
  for (int rows = 2; rows >= 1; --rows) {
    if (rows == 1 && P.mode == NS_GEMV_GATE_UP_SILU) continue;  // the gate/up epilogue needs both rows in one warp
    if (env_rows && rows != env_rows && !(env_rows == 1 && P.mode == NS_GEMV_GATE_UP_SILU)) continue;
    const int stage_bytes = rows * P.pitch;
    for (int i = 0; i < 2; ++i) {
      if (wide && i == 0) continue;  // the 14-consumer-warp kernel owns the SM
      const int kc = (wide && i == 1) ? 2 * kConsumers : kConsumers;
      int raw = 0;
      if (budgets[i] > act_region + 64) raw = (int)((budgets[i] - act_region - 64) / (stage_bytes + 16));
      if (raw > (wide ? 56 : 32)) raw = wide ? 56 : 32;
      const int ac = raw < kc ? raw : kc;
      if (ac < 1) continue;
      const int st = raw - raw % ac;  // one consumer warp per stage residue class (see kernel)
      const int ctas = i == 0 ? 2 : 1;
      const double score = ctas * ac * (rows == 2 ? 1.1 : 1.0) + 0.001 * st;
      if (score > best.score) best = RingPlan{rows, st, ac, ctas, budgets[i], score};
    }
  }
  return best;
}

template <int AMODE, int M, bool ASYM, int STYPE, int ROWS, bool NORM, int NC = kConsumers>
int launch_rows(const GemvParams& P, const RingPlan& plan, size_t act_region, int act_row, int red_off, cudaStream_t st) {
  auto kern = gemv_ring_kernel<AMODE, M, ASYM, STYPE, ROWS, NORM, NC>;
  constexpr int threads = (NC + (NC > kConsumers ? 2 : 1)) * 32;
  static bool attr_set = false;
  if (!attr_set) {
    NS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  const int stage_bytes = ROWS * P.pitch;
  const size_t smem = act_region + (size_t)plan.stages * stage_bytes + (size_t)plan.stages * 16;
  static const int env_cps = getenv("NS_RING_CPS") ? atoi(getenv("NS_RING_CPS")) : 0;  // tuning aid
  const int ctas_per_sm = plan.ctas == 1 ? 1 : (env_cps > 0 ? env_cps : 2);
  RingCfg R;
  R.ring_off = (int)act_region;
  R.stages = plan.stages;
  R.active = plan.active;
  R.act_row = act_row;
  R.red_off = red_off;
  if (ROWS == 2) {
    R.units = P.npairs;
  } else {
    long long rows = 0;
    for (int i = 0; i < P.nw; ++i) rows += P.n[i];
    R.units = (int)rows;
  }
  R.cpg_magic = P.cpg > 1 ? (uint32_t)((0x100000000ull + (uint64_t)P.cpg - 1) / (uint64_t)P.cpg) : 0u;
  int grid = ns_num_sms() * ctas_per_sm;
  if (grid > R.units) grid = R.units;
  if (grid < 1) grid = 1;
  static const bool dbg = getenv("NS_RING_DEBUG") != nullptr;
  if (dbg)
    fprintf(stderr, "gemv_ring: k=%d pitch=%d rows/stage=%d stages=%d active=%d smem=%zu ctas/sm=%d\n", P.k, P.pitch, ROWS, plan.stages,
            plan.active, smem, ctas_per_sm);
  NS_CUDA_TRY(ns_launch_pdl(kern, dim3(grid), dim3(threads), smem, st, P, R));
  ns_count_launch();
  return NS_OK;
}

}  // namespace
