// program.cu -- persistent multi-op decode kernel: a whole token's worth of weight-only matmuls in ONE launch.
//
// What it replaces: the reference rebuilds an ne graph per token and walks it node by node (ne_graph_compute,
// neural_speed/core/ne_layers.c:11915; llama graph, models/llama/llama.cpp:136-143,217-231,586,612-618,718); every
// matmul node first quantises its activations (NE_TASK_INIT, ne_layers.c:7143-7157) and then runs the dots.  On B200 a
// decode GEMV lasts 1.5-8 us, so a kernel boundary (drain + launch + refill of the load pipeline, ~2 us of idle HBM)
// costs as much as the work.  Here an "ns_program" is the list of matmul nodes of one token; one cooperative launch
// executes all of them:
//   * producer warp (1 elected thread per CTA): walks the op list and streams weight-row pairs with cp.async.bulk into the
//     shared-memory ring, never waiting for activations -- it runs ahead across op boundaries, so HBM stays busy while the
//     consumers synchronise;
//   * consumer warps: per op: grid barrier (release/acquire counter in global memory) -> quantise the op's fp32 input
//     vector(s) into the shared-memory activation image (same arithmetic as act_prep.cu: Q8_0 / BesTLA u8 / s8, bit-exact)
//     -> dp4a over their ring stages (same arithmetic as gemv_ring.cu) -> epilogue (bias / residual / SiLU*mul).
// Activations are read with ld.global.cg (L2) because another SM rewrites them between ops within the same launch.
// Roofline: HBM; algorithmic bytes per launch = sum over ops of N*K/2 + N*ceil(K/g)*(scale_bytes [+1 if asym]).
#include <vector>

#include "nsb.cuh"
#include "quant_smem.cuh"

namespace {

constexpr int kConsumers = 8;
constexpr int kConsumerThreads = kConsumers * 32;
constexpr int kThreads = kConsumerThreads + 32;

struct ProgOp {
  const uint8_t* rows[3];
  int n[3];
  long long dst_off[3];
  int nw, mode;
  int k, kpad, pitch, sc_off, zp_off, cpg, group;
  uint32_t cpg_magic;
  const float* in;
  int lda;
  float* dst;
  int ldo;
  const float* bias;
  int bias_bcast;
  const float* residual;
  float* aux;
  int npairs;
  int barrier_before;
  int act_row, meta_off, meta_stride;
  int pps;     // row pairs per ring slot (small rows are packed so a slot stays full)
  int nunits;  // ceil(npairs / pps)
};

struct ProgCfg {
  int ring_off;
  int stages;
  int slot_bytes;
  int m;
};

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
__device__ __forceinline__ void sts64(uint32_t a, uint32_t x, uint32_t y) {
  asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
template <int STYPE>
__device__ __forceinline__ float lds_scale(uint32_t base, int idx) {
  if (STYPE == NS_S_F32) return __uint_as_float(lds32(base + 4 * idx));
  if (STYPE == NS_S_F16) return __half2float(__ushort_as_half((unsigned short)lds16(base + 2 * idx)));
  return __uint_as_float(lds16(base + 2 * idx) << 16);
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
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

struct PairSrc {
  const uint8_t* r0;
  const uint8_t* r1;
  long long out0, out1;
  bool valid1;
};
__device__ __forceinline__ PairSrc resolve_pair(const ProgOp& P, int p) {
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
  s.valid1 = row + 1 < P.n[wi];
  s.r0 = P.rows[wi] + (size_t)row * P.pitch;
  s.r1 = s.valid1 ? s.r0 + P.pitch : s.r0;
  s.out0 = P.dst_off[wi] + row;
  s.out1 = s.out0 + 1;
  return s;
}

template <int COMP, int M, bool ASYM, int STYPE>
__global__ void __launch_bounds__(kThreads, 2)
    program_kernel(const ProgOp* __restrict__ ops, int nops, const ProgCfg R, unsigned* __restrict__ counters,
                   unsigned* __restrict__ epoch_ptr) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ ProgOp op_s;  // the consumers' current op
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int stages = R.stages;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t ring = smem_base + R.ring_off;
  const uint32_t full0 = ring + (uint32_t)stages * R.slot_bytes;
  const uint32_t empty0 = full0 + 8u * stages;
  const int first = blockIdx.x, gstride = (int)gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const unsigned epoch = ld_acquire(epoch_ptr);            // launches completed so far
  const unsigned target = (epoch + 1u) * (unsigned)gstride;  // every CTA arrives once per op per launch

  if (warp == kConsumers) {
    // ===================== producer: streams the weights of ALL ops, never waits for activations =====================
    if (lane == 0) {
      int s = 0;
      uint32_t phase = 0;
      long long g = 0;
      for (int oi = 0; oi < nops; ++oi) {
        const ProgOp& P = ops[oi];
        const int my_units = first < P.nunits ? (P.nunits - first + gstride - 1) / gstride : 0;
        for (int j = 0; j < my_units; ++j, ++g) {
          if (g >= stages) mbar_wait(empty0 + 8 * s, phase ^ 1);
          const int p0 = (first + j * gstride) * P.pps;
          const int np = min(P.pps, P.npairs - p0);
          const uint32_t dst = ring + (uint32_t)s * R.slot_bytes;
          mbar_expect_tx(full0 + 8 * s, 2u * (uint32_t)np * (uint32_t)P.pitch);
          if (P.mode == NS_GEMV_GATE_UP_SILU) {
            // slot = [np gate rows][np up rows]: two contiguous ranges
            bulk_g2s(dst, P.rows[0] + (size_t)p0 * P.pitch, (uint32_t)(np * P.pitch), full0 + 8 * s);
            bulk_g2s(dst + np * P.pitch, P.rows[1] + (size_t)p0 * P.pitch, (uint32_t)(np * P.pitch), full0 + 8 * s);
          } else {
            for (int t = 0; t < np; ++t) {  // slot = [pair 0: row, row+1][pair 1: ...]
              const PairSrc ps = resolve_pair(P, p0 + t);
              const uint32_t d = dst + (uint32_t)t * 2u * (uint32_t)P.pitch;
              if (ps.r1 == ps.r0 + P.pitch) {
                bulk_g2s(d, ps.r0, 2u * (uint32_t)P.pitch, full0 + 8 * s);
              } else {
                bulk_g2s(d, ps.r0, (uint32_t)P.pitch, full0 + 8 * s);
                bulk_g2s(d + P.pitch, ps.r1, (uint32_t)P.pitch, full0 + 8 * s);
              }
            }
          }
          if (++s == stages) {
            s = 0;
            phase ^= 1;
          }
        }
      }
    }
    return;
  }

  // ===================== consumers =====================
  constexpr int AMODE = (COMP == NS_COMP_INT8) ? A_U8 : A_S8;
  int s = warp;  // stages is a multiple of kConsumers: stage class == warp (see gemv_ring.cu)
  uint32_t phase = 0;
  int g_mod = 0;  // (global unit index of this CTA's next op start) mod kConsumers
  for (int oi = 0; oi < nops; ++oi) {
    // ---- op boundary: wait for the producers of this op's input, load the op descriptor ----
    {
      // descriptor (immutable) is fetched by many threads while thread 0 waits for the previous op to finish everywhere
      constexpr int kWords = (int)(sizeof(ProgOp) / 4);
      if (threadIdx.x >= 32 && threadIdx.x < 32 + kWords)
        reinterpret_cast<uint32_t*>(&op_s)[threadIdx.x - 32] = reinterpret_cast<const uint32_t*>(ops + oi)[threadIdx.x - 32];
      if (threadIdx.x == 0 && oi > 0 && ops[oi].barrier_before) {
        while ((int)(ld_acquire(counters + (oi - 1)) - target) < 0) {
        }
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory");
    const ProgOp& P = op_s;
    {
      QuantIn qi{P.in, P.lda, P.k, P.kpad, P.group, P.act_row, P.meta_off, P.meta_stride};
      nsq::quantise_to_smem<COMP, kConsumerThreads>(qi, R.m, smem_base);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory");

    const int my_units = first < P.nunits ? (P.nunits - first + gstride - 1) / gstride : 0;
    const uint32_t meta_s = smem_base + P.meta_off;
    const int nchunks = P.kpad >> 5;
    int u0 = warp - g_mod;
    if (u0 < 0) u0 += kConsumers;
    for (int j = u0; j < my_units; j += kConsumers) {
      const int p0 = (first + j * gstride) * P.pps;
      const int np = min(P.pps, P.npairs - p0);
      mbar_wait(full0 + 8 * s, phase);
      const uint32_t slot = ring + (uint32_t)s * R.slot_bytes;
     for (int t = 0; t < np; ++t) {
      const PairSrc ps = resolve_pair(P, p0 + t);
      const uint32_t r0 = (P.mode == NS_GEMV_GATE_UP_SILU) ? slot + (uint32_t)(t * P.pitch) : slot + (uint32_t)(t * 2 * P.pitch);
      const uint32_t r1 = (P.mode == NS_GEMV_GATE_UP_SILU) ? slot + (uint32_t)((np + t) * P.pitch) : r0 + P.pitch;
      float acc[2][M];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int m = 0; m < M; ++m) acc[r][m] = 0.f;
#pragma unroll 2
      for (int c = lane; c < nchunks; c += 32) {
        const uint4 wv[2] = {lds128(r0 + 16 * c), lds128(r1 + 16 * c)};
        const int gi = (P.cpg == 1) ? c : (int)__umulhi((uint32_t)c, P.cpg_magic);
        const float ws[2] = {lds_scale<STYPE>(r0 + P.sc_off, gi), lds_scale<STYPE>(r1 + P.sc_off, gi)};
        int off[2] = {8, 8};
        if (ASYM) {
          off[0] += lds8s(r0 + P.zp_off + gi);
          off[1] += lds8s(r1 + P.zp_off + gi);
        }
        uint32_t lo[2][4], hi[2][4];
        int su[2] = {0, 0};
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const uint32_t ww[4] = {wv[r].x, wv[r].y, wv[r].z, wv[r].w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            lo[r][i] = ww[i] & 0x0F0F0F0Fu;
            hi[r][i] = ww[i] & 0xF0F0F0F0u;
          }
          if (AMODE == A_U8) {
            int sl = 0, sh = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              sl = dp4a_uu(lo[r][i], 0x01010101u, sl);
              sh = dp4a_uu(hi[r][i], 0x01010101u, sh);
            }
            su[r] = sl + (sh >> 4);
          }
        }
        const uint32_t a_off = (uint32_t)(c >> 5) * 1024u + (uint32_t)(c & 31) * 16u;
#pragma unroll
        for (int m = 0; m < M; ++m) {
          const uint32_t ab = smem_base + (uint32_t)m * P.act_row + a_off;
          const uint4 a0 = lds128(ab), a1 = lds128(ab + 512);
          const uint2 mt = lds64(meta_s + 8u * (uint32_t)(m * P.meta_stride + c));
          const float a_scale = __uint_as_float(mt.x);
          const int sa = (int)(short)(mt.y & 0xffff);
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            int pl = 0, ph = 0;
            if (AMODE == A_U8) {
              pl = dp4a_uu(a0.x, lo[r][0], pl); ph = dp4a_uu(a0.y, hi[r][0], ph);
              pl = dp4a_uu(a0.z, lo[r][1], pl); ph = dp4a_uu(a0.w, hi[r][1], ph);
              pl = dp4a_uu(a1.x, lo[r][2], pl); ph = dp4a_uu(a1.y, hi[r][2], ph);
              pl = dp4a_uu(a1.z, lo[r][3], pl); ph = dp4a_uu(a1.w, hi[r][3], ph);
            } else {
              pl = dp4a_us(lo[r][0], (int)a0.x, pl); ph = dp4a_us(hi[r][0], (int)a0.y, ph);
              pl = dp4a_us(lo[r][1], (int)a0.z, pl); ph = dp4a_us(hi[r][1], (int)a0.w, ph);
              pl = dp4a_us(lo[r][2], (int)a1.x, pl); ph = dp4a_us(hi[r][2], (int)a1.y, ph);
              pl = dp4a_us(lo[r][3], (int)a1.z, pl); ph = dp4a_us(hi[r][3], (int)a1.w, ph);
            }
            int isum = pl + (ph >> 4) - off[r] * sa;
            if (AMODE == A_U8) {
              const int za = (int)((mt.y >> 16) & 0xff);
              isum -= za * (su[r] - 32 * off[r]);
            }
            acc[r][m] = fmaf((float)isum, a_scale * ws[r], acc[r][m]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int m = 0; m < M; ++m) acc[r][m] = warp_sum(acc[r][m]);
      if (lane == 0) {
        if (P.mode == NS_GEMV_GATE_UP_SILU) {
#pragma unroll
          for (int m = 0; m < M; ++m) {
            if (m < R.m) {
              const float gt = acc[0][m], up = acc[1][m];
              const float sg = gt / (1.f + expf(-gt));
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
              if (m < R.m) {
                const size_t o = (size_t)m * P.ldo + out;
                float v = acc[r][m];
                if (P.bias) v += P.bias_bcast ? ldcg1(P.bias + out) : ldcg1(P.bias + o);
                if (P.residual) v += ldcg1(P.residual + o);
                P.dst[o] = v;
              }
            }
          }
        }
      }
     }  // pairs of this slot
      __syncwarp();
      if (lane == 0) mbar_arrive(empty0 + 8 * s);  // slot may be refilled
      s += kConsumers;
      if (s >= stages) {
        s -= stages;
        phase ^= 1u;
      }
    }
    g_mod = (g_mod + my_units) % kConsumers;
    // ---- op done in this CTA: publish (release) ----
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory");
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(counters + oi, 1u);
    }
  }
  // last op finished everywhere -> advance the epoch exactly once (block 0), so the next launch sees fresh targets
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    while ((int)(ld_acquire(counters + (nops - 1)) - target) < 0) {
    }
    __threadfence();
    atomicAdd(epoch_ptr, 1u);
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------- host side
struct ns_program {
  int m;
  int comp, stype, asym;
  bool finalized;
  std::vector<ProgOp> ops;
  ProgOp* d_ops;
  unsigned* d_counters;  // [nops] + epoch at [nops]
  ProgCfg cfg;
  size_t smem;
  int grid;
  size_t alg_bytes;
};

extern "C" ns_program* ns_program_create(int m) {
  if (ns_ensure_device()) return nullptr;
  if (m < 1 || m > 4) {
    ns_set_error("ns_program_create: m must be 1..4 (decode batches; larger M goes through the tensor-core GEMM)");
    return nullptr;
  }
  ns_program* p = new ns_program();
  p->m = m;
  p->comp = -1;
  p->finalized = false;
  p->d_ops = nullptr;
  p->d_counters = nullptr;
  p->alg_bytes = 0;
  return p;
}

extern "C" int ns_program_add_matmul(ns_program* p, const ns_weight* const* weights, int nw, int mode, const float* in, int lda,
                                     float* dst, int ldo, const float* bias, int bias_bcast, const float* residual,
                                     float* aux, int barrier_before) {
  if (!p || p->finalized || !weights || nw < 1 || nw > 3 || mode < 0 || mode > 2 || !in || !dst) {
    ns_set_error("ns_program_add_matmul: invalid arguments");
    return NS_E_INVALID;
  }
  const ns_weight* w0 = weights[0];
  const bool imode = (w0->comp == NS_COMP_Q8_0 || w0->comp == NS_COMP_INT8 || w0->comp == NS_COMP_INT8_S8);
  const int qgroup = w0->comp == NS_COMP_Q8_0 ? 32 : w0->group;
  if (w0->wfmt != NS_W_S4 || !imode || w0->shuffle || !(qgroup == 32 || qgroup == 64 || qgroup == 128 || qgroup == 256) ||
      (w0->group % 32 != 0) || (w0->k % qgroup != 0)) {
    ns_set_error("ns_program: only 4-bit integer weights with integer activations and groups of 32..256 are supported");
    return NS_E_UNSUPPORTED;
  }
  if (p->comp < 0) {
    p->comp = w0->comp;
    p->stype = w0->stype;
    p->asym = w0->asym;
  }
  long long ntot = 0;
  ProgOp op;
  memset(&op, 0, sizeof(op));
  for (int i = 0; i < nw; ++i) {
    const ns_weight* wi = weights[i];
    if (wi->comp != p->comp || wi->stype != p->stype || wi->asym != p->asym || wi->wfmt != NS_W_S4 || wi->k != w0->k ||
        wi->group != w0->group || wi->shuffle) {
      ns_set_error("ns_program: all weights of a program must share format, scale type and compute type");
      return NS_E_UNSUPPORTED;
    }
    if (mode == NS_GEMV_CONCAT && i + 1 < nw && (wi->n & 1)) {
      ns_set_error("ns_program: every weight but the last of a fused matmul needs an even n");
      return NS_E_UNSUPPORTED;
    }
    op.rows[i] = wi->rows;
    op.n[i] = wi->n;
    op.dst_off[i] = (mode == NS_GEMV_CONCAT) ? ntot : 0;  // concatenated along n: [m][n0+n1+n2] with ldo
    ntot += wi->n;
    p->alg_bytes += ns_weight_algorithmic_bytes(wi);
  }
  if (mode == NS_GEMV_GATE_UP_SILU && (nw != 2 || weights[0]->n != weights[1]->n)) {
    ns_set_error("ns_program: gate/up fusion needs two weights with equal n");
    return NS_E_INVALID;
  }
  op.nw = nw;
  op.mode = mode;
  op.k = w0->k;
  op.kpad = w0->kpad;
  op.pitch = w0->pitch;
  op.sc_off = w0->sc_off;
  op.zp_off = w0->zp_off;
  op.group = w0->group;
  op.cpg = (w0->group + 31) / 32;
  op.cpg_magic = op.cpg > 1 ? (uint32_t)((0x100000000ull + (uint64_t)op.cpg - 1) / (uint64_t)op.cpg) : 0u;
  op.in = in;
  op.lda = lda;
  op.dst = dst;
  op.ldo = ldo;
  op.bias = bias;
  op.bias_bcast = bias_bcast;
  op.residual = residual;
  op.aux = aux;
  op.npairs = (mode == NS_GEMV_GATE_UP_SILU) ? w0->n : (int)((ntot + 1) / 2);
  op.barrier_before = barrier_before;
  op.act_row = (int)ns_round_up((size_t)w0->kpad, 1024);
  op.meta_stride = ns_meta_stride(w0->kpad);
  op.meta_off = p->m * op.act_row;
  p->ops.push_back(op);
  return NS_OK;
}

extern "C" size_t ns_program_algorithmic_bytes(const ns_program* p) { return p ? p->alg_bytes : 0; }

extern "C" int ns_program_finalize(ns_program* p, void* queue) {
  if (!p || p->ops.empty()) return NS_E_INVALID;
  if (p->finalized) return NS_OK;
  cudaStream_t st = ns_stream_of(queue);
  const int mt = p->m >= 3 ? 4 : p->m;
  size_t act_region = 0;
  int slot = 0;
  for (const ProgOp& o : p->ops) {
    act_region = std::max(act_region, ns_round_up((size_t)mt * o.act_row + (size_t)mt * o.meta_stride * 8, 128));
    slot = std::max(slot, 2 * o.pitch);
  }
  const size_t budgets[2] = {113 * 1024, 200 * 1024};  // 2 x (113 KB + 1 KB reserved) = 228 KB = one SM
  int stages = 0;
  size_t budget = 0;
  for (int i = 0; i < 2; ++i) {
    budget = budgets[i];
    if (budget > act_region + 64) stages = (int)((budget - act_region - 64) / ((size_t)slot + 16));
    stages -= stages % kConsumers;
    if (stages >= kConsumers) break;
    stages = 0;
  }
  if (stages < kConsumers) {
    ns_set_error("ns_program: rows too long for the shared-memory ring");
    return NS_E_UNSUPPORTED;
  }
  if (stages > 48) stages = 48;
  p->cfg.ring_off = (int)act_region;
  p->cfg.stages = stages;
  p->cfg.slot_bytes = slot;
  p->cfg.m = p->m;
  p->smem = act_region + (size_t)stages * slot + (size_t)stages * 16;
  p->grid = ns_num_sms() * (budget > 113 * 1024 ? 1 : 2);
  for (ProgOp& o : p->ops) {
    o.pps = std::max(1, std::min(4, slot / (2 * o.pitch)));
    o.nunits = (o.npairs + o.pps - 1) / o.pps;
  }
  const size_t nops = p->ops.size();
  NS_CUDA_TRY(cudaMalloc((void**)&p->d_ops, nops * sizeof(ProgOp)));
  NS_CUDA_TRY(cudaMalloc((void**)&p->d_counters, (nops + 1) * sizeof(unsigned)));
  NS_CUDA_TRY(cudaMemcpyAsync(p->d_ops, p->ops.data(), nops * sizeof(ProgOp), cudaMemcpyHostToDevice, st));
  NS_CUDA_TRY(cudaMemsetAsync(p->d_counters, 0, (nops + 1) * sizeof(unsigned), st));
  NS_CUDA_TRY(cudaStreamSynchronize(st));
  p->finalized = true;
  return NS_OK;
}

template <int COMP, int M, bool ASYM, int STYPE>
static int run_one(ns_program* p, cudaStream_t st) {
  auto kern = program_kernel<COMP, M, ASYM, STYPE>;
  static bool attr_set = false;
  if (!attr_set) {
    NS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p->grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = p->smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs must be co-resident: they synchronise through global memory
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const ProgOp* ops = p->d_ops;
  int nops = (int)p->ops.size();
  unsigned* counters = p->d_counters;
  unsigned* epoch = p->d_counters + nops;
  NS_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, ops, nops, p->cfg, counters, epoch));
  ns_count_launch();
  return NS_OK;
}
template <int COMP, bool ASYM, int STYPE>
static int run_m(ns_program* p, cudaStream_t st) {
  switch (p->m) {
    case 1: return run_one<COMP, 1, ASYM, STYPE>(p, st);
    case 2: return run_one<COMP, 2, ASYM, STYPE>(p, st);
    default: return run_one<COMP, 4, ASYM, STYPE>(p, st);
  }
}
template <int COMP, bool ASYM>
static int run_s(ns_program* p, cudaStream_t st) {
  switch (p->stype) {
    case NS_S_F32: return run_m<COMP, ASYM, NS_S_F32>(p, st);
    case NS_S_F16: return run_m<COMP, ASYM, NS_S_F16>(p, st);
    default: return run_m<COMP, ASYM, NS_S_BF16>(p, st);
  }
}
template <int COMP>
static int run_a(ns_program* p, cudaStream_t st) {
  return p->asym ? run_s<COMP, true>(p, st) : run_s<COMP, false>(p, st);
}

extern "C" int ns_program_run(ns_program* p, void* queue) {
  if (int rc = ns_ensure_device()) return rc;
  if (!p || !p->finalized) {
    ns_set_error("ns_program_run: program not finalized");
    return NS_E_INVALID;
  }
  cudaStream_t st = ns_stream_of(queue);
  switch (p->comp) {
    case NS_COMP_Q8_0: return run_a<NS_COMP_Q8_0>(p, st);
    case NS_COMP_INT8: return run_a<NS_COMP_INT8>(p, st);
    default: return run_a<NS_COMP_INT8_S8>(p, st);
  }
}

extern "C" void ns_program_free(ns_program* p) {
  if (!p) return;
  if (p->d_ops) cudaFree(p->d_ops);
  if (p->d_counters) cudaFree(p->d_counters);
  delete p;
}
