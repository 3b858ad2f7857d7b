// program.cu -- the decode matmul engine for 4-bit integer weights x 8-bit integer activations, M <= 4 rows:
// a list of matmul nodes ("program") executed by ONE kernel.  A program of one node is the per-op GEMV
// (ns_mul_mat / ns_mul_qkv / ns_ffn_silu); a program of a whole token's nodes is one cooperative launch per token.
//
// What it replaces: the reference walks an ne graph node by node (ne_graph_compute, neural_speed/core/ne_layers.c:11915;
// llama graph, models/llama/llama.cpp:217-231,586,612-618,718); every matmul node quantises its activations
// (NE_TASK_INIT, ne_layers.c:7143-7157; quantize_row_q8_0, vectors/cpu/quantize.h:447; ActivationKBlockQuantize,
// bestla/bestla/bestla_prologue_a.h:105) and then runs the dots (ne_vec_dot_q4_0_q8_0, core/layers/vec_dot.h:131;
// gemv_4bit_u8s8_fp32 / _s8s8_, bestla/bestla/kernel_ref.h:2372/2432).
//
// B200 mapping.  HBM -> shared memory -> dp4a, with the two halves decoupled:
//   * producer warp (one elected thread per CTA) walks the op list and streams weights with cp.async.bulk (TMA 1-D,
//     SASS UBLKCP).  A unit is a SEGMENT (64 chunks = 2048 k) of a PAIR of weight rows: 2 x 1 KB of nibbles + their
//     scales (+ zero points) -- ~2.3 KB.  Each consumer warp owns a private sub-ring of SPW slots (>= 4), so a warp always
//     has several units landing while it computes one; the producer never waits for activations and runs ahead across
//     op boundaries, which keeps HBM busy while consumers synchronise.
//   * 8 consumer warps per CTA, 2 CTAs per SM.  Per op: [grid barrier when the input comes from the previous op] ->
//     quantise the op's fp32 input rows into the shared-memory activation image (bit-exact with act_prep.cu / the
//     reference quantisers) -> per unit: nibbles & 0x0F0F0F0F / & 0xF0F0F0F0 against permuted int8 activations with dp4a
//     (every 32-element chunk dot is an exact integer), fp32 fma(isum, a_scale * w_scale) -> after a pair's last
//     segment: warp reduce, epilogue (bias / residual / SiLU*mul), store.
//   * activations written by other SMs inside the same launch are read with ld.global.cg; the grid barrier is a
//     release/acquire counter per op in global memory, made launch-invariant by an epoch word.
// Roofline: HBM.  Algorithmic bytes per launch = sum over ops of N*K/2 + N*ceil(K/g)*(scale_bytes [+1 if asym]).
#include <algorithm>
#include <vector>

#include "nsb.cuh"

namespace {

constexpr int kConsumers = 8;
constexpr int kConsumerThreads = kConsumers * 32;
constexpr int kThreads = kConsumerThreads + 32;
constexpr int SEGC = 64;  // 32-element chunks per segment

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
  int nseg;       // segments per row
  int sc_seg;     // scale bytes per full segment (16-B multiple)
  int zp_seg;     // zero-point bytes per full segment (16-B multiple, 0 if symmetric)
};

struct ProgCfg {
  int ring_off;
  int spw;         // slots per consumer warp
  int slot_bytes;  // 2 * (SEGC*16 + sc_seg + zp_seg), max over ops
  int m;
  int pdl;         // 1: single-op launch under programmatic dependent launch (no grid barrier, no epoch)
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

// Quantise the op's activations [M][K] (fp32, global, read through L2) into the shared-memory image:
// per row: bytes in dp4a order, per 32-chunk super-block the first 16 B of every chunk then the second 16 B
// (conflict-free LDS.128), followed by per-chunk meta {a_scale, (Sa & 0xffff) | za << 16}.
// One thread owns one 8-group (8 consecutive k); tpb = group/8 consecutive threads own one quantisation block.
// Arithmetic identical to act_quant_kernel<COMP> (act_prep.cu): bit-exact codes, scales and zero points.
template <int COMP>
__device__ __forceinline__ void quantise_to_smem(const ProgOp& P, int M, uint32_t smem_base) {
  const int tpb = (COMP == NS_COMP_Q8_0 ? 32 : P.group) >> 3;  // threads per quantisation block (4..32, power of two)
  const int ngroups8 = P.kpad >> 3;
  const int tid = threadIdx.x;
  constexpr int NI = 3;  // loads of NI passes are issued back to back (each is an L2 round trip)
  for (int m = 0; m < M; ++m) {
    const float* row = P.in + (size_t)m * P.lda;
    const uint32_t img = smem_base + (uint32_t)m * P.act_row;
    const uint32_t meta = smem_base + P.meta_off + 8u * (uint32_t)(m * P.meta_stride);
    for (int eb = 0; eb < ngroups8; eb += NI * kConsumerThreads) {
      float vv[NI][8];
#pragma unroll
      for (int it = 0; it < NI; ++it) {
        const int e = eb + it * kConsumerThreads + tid;
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
        const int e0 = eb + it * kConsumerThreads;
        if (e0 >= ngroups8) break;  // uniform across the CTA
        const int e = e0 + tid;
        const bool live = e < ngroups8;
        const int k0 = e * 8;
        float vmax = (COMP == NS_COMP_Q8_0) ? 0.f : 1.17549435e-38f, vmin = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (COMP == NS_COMP_INT8) {
            vmax = fmaxf(vv[it][i], vmax);
            vmin = fminf(vv[it][i], vmin);
          } else {
            vmax = fmaxf(vmax, fabsf(vv[it][i]));
          }
        }
        for (int o = 1; o < tpb; o <<= 1) {  // all 32 lanes take part
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
            if (COMP == NS_COMP_Q8_0) q[i] = __float2int_rn(vv[it][i] * rscale);
            else if (COMP == NS_COMP_INT8) q[i] = cast_u8((float)za + (float)(int)roundf(vv[it][i] * rscale));
            else q[i] = cast_s8(vv[it][i] * rscale);
          } else {
            q[i] = za;  // padding contributes (a - za) == 0
          }
          sa += q[i];
        }
        sa += __shfl_xor_sync(0xffffffffu, sa, 1);  // sum over the 4 threads of a 32-element chunk
        sa += __shfl_xor_sync(0xffffffffu, sa, 2);
        if (live) {
          // bytes in dp4a order: Alo = (a0,a4,a1,a5), Ahi = (a2,a6,a3,a7)
          const uint32_t alo = (q[0] & 0xff) | ((q[4] & 0xff) << 8) | ((q[1] & 0xff) << 16) | ((uint32_t)(q[5] & 0xff) << 24);
          const uint32_t ahi = (q[2] & 0xff) | ((q[6] & 0xff) << 8) | ((q[3] & 0xff) << 16) | ((uint32_t)(q[7] & 0xff) << 24);
          const int c = e >> 2, i = e & 3;
          sts64(img + (uint32_t)(c >> 5) * 1024u + (uint32_t)(i >> 1) * 512u + (uint32_t)(c & 31) * 16u + (uint32_t)(i & 1) * 8u,
                alo, ahi);
          if (i == 0) sts64(meta + 8u * (uint32_t)c, __float_as_uint(scale), (uint32_t)((sa & 0xffff) | (za << 16)));
        }
      }
    }
  }
}

// slot layout: [row0 q : SEGC*16][row1 q : SEGC*16][row0 scales : sc_seg][row1 scales : sc_seg][row0 zp : zp_seg][row1 zp]
struct SegInfo {
  int nch;          // chunks in this segment
  uint32_t qb, sb, zb;  // bytes actually copied per row for q / scales / zp
};
__device__ __forceinline__ SegInfo seg_info(const ProgOp& P, int seg, int ssize) {
  SegInfo si;
  const int nchunks = P.kpad >> 5;
  si.nch = min(SEGC, nchunks - seg * SEGC);
  si.qb = (uint32_t)si.nch * 16u;
  const int ng = (si.nch + P.cpg - 1) / P.cpg;
  si.sb = ((uint32_t)(ng * ssize) + 15u) & ~15u;
  si.zb = P.zp_seg ? (((uint32_t)ng + 15u) & ~15u) : 0u;
  return si;
}

template <int COMP, int M, bool ASYM, int STYPE>
__global__ void __launch_bounds__(kThreads, 2)
    program_kernel(const ProgOp op0, const ProgOp* __restrict__ ops, int nops, const ProgCfg R, unsigned* __restrict__ counters,
                   unsigned* __restrict__ epoch_ptr) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ ProgOp op_s;  // the consumers' current op
  constexpr int SSIZE = (STYPE == NS_S_F32) ? 4 : 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int spw = R.spw;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t ring = smem_base + R.ring_off;
  const int nslots = spw * kConsumers;
  const uint32_t full0 = ring + (uint32_t)nslots * R.slot_bytes;
  const uint32_t empty0 = full0 + 8u * nslots;
  const int first = blockIdx.x, gstride = (int)gridDim.x;

  if (R.pdl) pdl_launch_dependents();
  if (threadIdx.x == 0) {
    for (int s = 0; s < nslots; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  unsigned target = 0;
  if (!R.pdl) {
    const unsigned epoch = ld_acquire(epoch_ptr);     // launches completed so far
    target = (epoch + 1u) * (unsigned)gstride;        // every CTA arrives once per op per launch
  }

  if (warp == kConsumers) {
    // ===================== producer: streams the weights of ALL ops, never waits for activations =====================
    if (lane == 0) {
      // sub-ring w = slots [w*spw, (w+1)*spw); cursor / phase / use count per consumer warp
      int cur[kConsumers], used[kConsumers];
      uint32_t ph[kConsumers];
#pragma unroll
      for (int w = 0; w < kConsumers; ++w) {
        cur[w] = 0;
        used[w] = 0;
        ph[w] = 0;
      }
      for (int oi = 0; oi < nops; ++oi) {
        const ProgOp& P = ops ? ops[oi] : op0;
        const int my_pairs = first < P.npairs ? (P.npairs - first + gstride - 1) / gstride : 0;
        for (int grp = 0; grp < my_pairs; grp += kConsumers) {
          for (int seg = 0; seg < P.nseg; ++seg) {
            const SegInfo si = seg_info(P, seg, SSIZE);
            const int gseg = (seg * SEGC) / P.cpg;  // first scale group of this segment
#pragma unroll
            for (int w = 0; w < kConsumers; ++w) {
              const int lp = grp + w;
              if (lp >= my_pairs) continue;
              const int slot = w * spw + cur[w];
              if (used[w] >= spw) mbar_wait(empty0 + 8 * slot, ph[w] ^ 1);
              const PairSrc ps = resolve_pair(P, first + lp * gstride);
              const uint32_t dst = ring + (uint32_t)slot * R.slot_bytes;
              const uint32_t fb = full0 + 8 * slot;
              mbar_expect_tx(fb, 2u * (si.qb + si.sb + si.zb));
              bulk_g2s(dst, ps.r0 + (size_t)seg * SEGC * 16, si.qb, fb);
              bulk_g2s(dst + SEGC * 16, ps.r1 + (size_t)seg * SEGC * 16, si.qb, fb);
              const uint32_t so = dst + 2 * SEGC * 16;
              bulk_g2s(so, ps.r0 + P.sc_off + (size_t)gseg * SSIZE, si.sb, fb);
              bulk_g2s(so + P.sc_seg, ps.r1 + P.sc_off + (size_t)gseg * SSIZE, si.sb, fb);
              if (ASYM) {
                const uint32_t zo = so + 2 * P.sc_seg;
                bulk_g2s(zo, ps.r0 + P.zp_off + gseg, si.zb, fb);
                bulk_g2s(zo + P.zp_seg, ps.r1 + P.zp_off + gseg, si.zb, fb);
              }
              ++used[w];
              if (++cur[w] == spw) {
                cur[w] = 0;
                ph[w] ^= 1;
              }
            }
          }
        }
      }
    }
    return;
  }

  // ===================== consumers =====================
  constexpr int AMODE = (COMP == NS_COMP_INT8) ? A_U8 : A_S8;
  int cur = 0;
  uint32_t phase = 0;
  for (int oi = 0; oi < nops; ++oi) {
    // ---- op boundary: wait for the producers of this op's input, fetch the op descriptor ----
    {
      constexpr int kWords = (int)(sizeof(ProgOp) / 4);
      const ProgOp* src = ops ? ops + oi : &op0;
      if (threadIdx.x >= 32 && threadIdx.x < 32 + kWords)
        reinterpret_cast<uint32_t*>(&op_s)[threadIdx.x - 32] = reinterpret_cast<const uint32_t*>(src)[threadIdx.x - 32];
      if (R.pdl) {
        pdl_wait();  // single-op launch: activations come from the preceding kernel in the stream
      } else if (threadIdx.x == 0 && oi > 0 && ops[oi].barrier_before) {
        while ((int)(ld_acquire(counters + (oi - 1)) - target) < 0) {
        }
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory");
    const ProgOp& P = op_s;
    quantise_to_smem<COMP>(P, R.m, smem_base);
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory");

    const int my_pairs = first < P.npairs ? (P.npairs - first + gstride - 1) / gstride : 0;
    const uint32_t meta_s = smem_base + P.meta_off;
    for (int lp = warp; lp < my_pairs; lp += kConsumers) {
      const PairSrc ps = resolve_pair(P, first + lp * gstride);
      float acc[2][M];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int m = 0; m < M; ++m) acc[r][m] = 0.f;
      for (int seg = 0; seg < P.nseg; ++seg) {
        const int nch = min(SEGC, (P.kpad >> 5) - seg * SEGC);
        const int slot = warp * spw + cur;
        mbar_wait(full0 + 8 * slot, phase);
        const uint32_t r0 = ring + (uint32_t)slot * R.slot_bytes;
        const uint32_t r1 = r0 + SEGC * 16;
        const uint32_t s0 = r0 + 2 * SEGC * 16, s1 = s0 + P.sc_seg;
        const uint32_t z0 = s0 + 2 * P.sc_seg, z1 = z0 + P.zp_seg;
#pragma unroll 2
        for (int cl = lane; cl < nch; cl += 32) {
          const int c = seg * SEGC + cl;  // chunk index in the row
          const uint4 wv[2] = {lds128(r0 + 16 * cl), lds128(r1 + 16 * cl)};
          const int gl = (P.cpg == 1) ? cl : (int)__umulhi((uint32_t)cl, P.cpg_magic);  // group index inside the segment
          const float ws[2] = {lds_scale<STYPE>(s0, gl), lds_scale<STYPE>(s1, gl)};
          int off[2] = {8, 8};
          if (ASYM) {
            off[0] += lds8s(z0 + gl);
            off[1] += lds8s(z1 + gl);
          }
          uint32_t lo[2][4], hi[2][4];
          int su[2] = {0, 0};
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const uint32_t ww[4] = {wv[r].x, wv[r].y, wv[r].z, wv[r].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              lo[r][i] = ww[i] & 0x0F0F0F0Fu;
              hi[r][i] = ww[i] & 0xF0F0F0F0u;  // high nibbles as bytes * 16 (no shift): exact, divided out below
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
              // NSB4: word i pairs with activation words (Alo_i, Ahi_i) = ((a0,a4,a1,a5),(a2,a6,a3,a7)) of 8-group i
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
              // sum (a - za)(u - off) = sum a*u - off*Sa - za*(Su - 32*off): one exact integer per 32-element chunk
              int isum = pl + (ph >> 4) - off[r] * sa;
              if (AMODE == A_U8) {
                const int za = (int)((mt.y >> 16) & 0xff);
                isum -= za * (su[r] - 32 * off[r]);
              }
              acc[r][m] = fmaf((float)isum, a_scale * ws[r], acc[r][m]);
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * slot);  // slot may be refilled
        if (++cur == spw) {
          cur = 0;
          phase ^= 1u;
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
              const float sg = gt / (1.f + expf(-gt));  // swish alpha=-1 (kernel_ref.h:1574)
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
    }
    if (R.pdl) break;  // single op
    // ---- op done in this CTA: publish (release) ----
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory");
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(counters + oi, 1u);
    }
  }
  // last op finished everywhere -> advance the epoch exactly once (block 0), so the next launch sees fresh targets
  if (!R.pdl && blockIdx.x == 0 && threadIdx.x == 0) {
    while ((int)(ld_acquire(counters + (nops - 1)) - target) < 0) {
    }
    __threadfence();
    atomicAdd(epoch_ptr, 1u);
  }
}

// ---------------------------------------------------------------------------------------------------- host helpers
bool op_supported(const ns_weight* w0) {
  const bool imode = (w0->comp == NS_COMP_Q8_0 || w0->comp == NS_COMP_INT8 || w0->comp == NS_COMP_INT8_S8);
  const int qgroup = w0->comp == NS_COMP_Q8_0 ? 32 : w0->group;
  if (w0->wfmt != NS_W_S4 || !imode || w0->shuffle) return false;
  if (!(qgroup == 32 || qgroup == 64 || qgroup == 128 || qgroup == 256)) return false;  // activation block <= one warp
  if (w0->group % 32 != 0 || w0->k % qgroup != 0) return false;
  const int cpg = w0->group / 32;
  if (cpg > 8 || (cpg & (cpg - 1))) return false;  // scale bytes of a 64-chunk segment must be a 16-B multiple
  if (w0->asym && cpg > 4) return false;           // same for the int8 zero points
  return true;
}

int fill_op(ProgOp* op, int m_rows, const ns_weight* const* weights, int nw, int mode, const float* in, int lda, float* dst,
            int ldo, const float* bias, int bias_bcast, const float* residual, float* aux, int barrier_before, bool qkv_planes,
            int m_total) {
  const ns_weight* w0 = weights[0];
  memset(op, 0, sizeof(*op));
  long long ntot = 0;
  for (int i = 0; i < nw; ++i) {
    const ns_weight* wi = weights[i];
    if (wi->comp != w0->comp || wi->stype != w0->stype || wi->asym != w0->asym || wi->wfmt != NS_W_S4 || wi->k != w0->k ||
        wi->group != w0->group || wi->shuffle) {
      ns_set_error("fused matmul: weights differ in format");
      return NS_E_UNSUPPORTED;
    }
    if (mode == NS_GEMV_CONCAT && i + 1 < nw && (wi->n & 1)) {
      ns_set_error("fused matmul: every weight but the last needs an even n");
      return NS_E_UNSUPPORTED;
    }
    op->rows[i] = wi->rows;
    op->n[i] = wi->n;
    // concat: either one [m][n0+n1+n2] row (program API) or the reference's [nw][M][ldo] planes (ip_fusion_qkv.cpp:84-86)
    op->dst_off[i] = (mode == NS_GEMV_CONCAT) ? (qkv_planes ? (long long)i * m_total * ldo : ntot) : 0;
    ntot += wi->n;
  }
  if (mode == NS_GEMV_GATE_UP_SILU && (nw != 2 || weights[0]->n != weights[1]->n)) {
    ns_set_error("gate/up fusion needs two weights with equal n");
    return NS_E_INVALID;
  }
  const int ssize = ns_stype_size(w0->stype);
  op->nw = nw;
  op->mode = mode;
  op->k = w0->k;
  op->kpad = w0->kpad;
  op->pitch = w0->pitch;
  op->sc_off = w0->sc_off;
  op->zp_off = w0->zp_off;
  op->group = w0->group;
  op->cpg = w0->group / 32;
  op->cpg_magic = op->cpg > 1 ? (uint32_t)((0x100000000ull + (uint64_t)op->cpg - 1) / (uint64_t)op->cpg) : 0u;
  op->in = in;
  op->lda = lda;
  op->dst = dst;
  op->ldo = ldo;
  op->bias = bias;
  op->bias_bcast = bias_bcast;
  op->residual = residual;
  op->aux = aux;
  op->npairs = (mode == NS_GEMV_GATE_UP_SILU) ? w0->n : (int)((ntot + 1) / 2);
  op->barrier_before = barrier_before;
  op->act_row = (int)ns_round_up((size_t)w0->kpad, 1024);
  op->meta_stride = ns_meta_stride(w0->kpad);
  op->meta_off = m_rows * op->act_row;
  op->nseg = ((w0->kpad >> 5) + SEGC - 1) / SEGC;
  op->sc_seg = (int)ns_round_up((size_t)(SEGC / op->cpg) * ssize, 16);
  op->zp_seg = w0->asym ? (int)ns_round_up((size_t)(SEGC / op->cpg), 16) : 0;
  return NS_OK;
}

struct Plan {
  ProgCfg cfg;
  size_t smem;
  int grid;
};
int make_plan(const ProgOp* ops, int nops, int m, Plan* pl) {
  const int mt = m >= 3 ? 4 : m;
  size_t act_region = 0;
  int slot = 0;
  for (int i = 0; i < nops; ++i) {
    act_region = std::max(act_region, ns_round_up((size_t)mt * ops[i].act_row + (size_t)mt * ops[i].meta_stride * 8, 128));
    slot = std::max(slot, 2 * (SEGC * 16 + ops[i].sc_seg + ops[i].zp_seg));
  }
  const size_t budgets[2] = {113 * 1024, 200 * 1024};  // 2 x (113 KB + 1 KB reserved) = one SM
  int spw = 0;
  size_t budget = 0;
  for (int i = 0; i < 2; ++i) {
    budget = budgets[i];
    if (budget > act_region + 64) spw = (int)((budget - act_region - 64) / ((size_t)kConsumers * (slot + 16)));
    if (spw >= 3) break;
    spw = 0;
  }
  if (spw < 2) {
    ns_set_error("decode kernel: activation image (%zu B) leaves no room for the weight ring", act_region);
    return NS_E_UNSUPPORTED;
  }
  if (spw > 8) spw = 8;
  pl->cfg.ring_off = (int)act_region;
  pl->cfg.spw = spw;
  pl->cfg.slot_bytes = slot;
  pl->cfg.m = m;
  pl->cfg.pdl = 0;
  pl->smem = act_region + (size_t)spw * kConsumers * (slot + 16);
  pl->grid = ns_num_sms() * (budget > 113 * 1024 ? 1 : 2);
  return NS_OK;
}

template <int COMP, int M, bool ASYM, int STYPE>
int launch_k(const ProgOp& op0, const ProgOp* d_ops, int nops, const Plan& pl, unsigned* counters, unsigned* epoch,
             cudaStream_t st) {
  auto kern = program_kernel<COMP, M, ASYM, STYPE>;
  static bool attr_set = false;
  if (!attr_set) {
    NS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  int grid = pl.grid;
  if (pl.cfg.pdl && grid > op0.npairs) grid = op0.npairs > 0 ? op0.npairs : 1;
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = pl.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  if (pl.cfg.pdl) {
    static const bool no_pdl = getenv("NS_NO_PDL") != nullptr;
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = no_pdl ? 0 : 1;
  } else {
    attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident: they synchronise through global memory
    attr[0].val.cooperative = 1;
    cfg.numAttrs = 1;
  }
  cfg.attrs = attr;
  NS_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, op0, d_ops, nops, pl.cfg, counters, epoch));
  ns_count_launch();
  return NS_OK;
}
template <int COMP, bool ASYM, int STYPE>
int launch_m(int m, const ProgOp& op0, const ProgOp* d_ops, int nops, const Plan& pl, unsigned* c, unsigned* e, cudaStream_t st) {
  switch (m) {
    case 1: return launch_k<COMP, 1, ASYM, STYPE>(op0, d_ops, nops, pl, c, e, st);
    case 2: return launch_k<COMP, 2, ASYM, STYPE>(op0, d_ops, nops, pl, c, e, st);
    default: return launch_k<COMP, 4, ASYM, STYPE>(op0, d_ops, nops, pl, c, e, st);
  }
}
template <int COMP, bool ASYM>
int launch_s(int stype, int m, const ProgOp& op0, const ProgOp* d_ops, int nops, const Plan& pl, unsigned* c, unsigned* e,
             cudaStream_t st) {
  switch (stype) {
    case NS_S_F32: return launch_m<COMP, ASYM, NS_S_F32>(m, op0, d_ops, nops, pl, c, e, st);
    case NS_S_F16: return launch_m<COMP, ASYM, NS_S_F16>(m, op0, d_ops, nops, pl, c, e, st);
    default: return launch_m<COMP, ASYM, NS_S_BF16>(m, op0, d_ops, nops, pl, c, e, st);
  }
}
int launch_any(int comp, int asym, int stype, int m, const ProgOp& op0, const ProgOp* d_ops, int nops, const Plan& pl,
               unsigned* c, unsigned* e, cudaStream_t st) {
  switch (comp) {
    case NS_COMP_Q8_0:
      return asym ? launch_s<NS_COMP_Q8_0, true>(stype, m, op0, d_ops, nops, pl, c, e, st)
                  : launch_s<NS_COMP_Q8_0, false>(stype, m, op0, d_ops, nops, pl, c, e, st);
    case NS_COMP_INT8:
      return asym ? launch_s<NS_COMP_INT8, true>(stype, m, op0, d_ops, nops, pl, c, e, st)
                  : launch_s<NS_COMP_INT8, false>(stype, m, op0, d_ops, nops, pl, c, e, st);
    default:
      return asym ? launch_s<NS_COMP_INT8_S8, true>(stype, m, op0, d_ops, nops, pl, c, e, st)
                  : launch_s<NS_COMP_INT8_S8, false>(stype, m, op0, d_ops, nops, pl, c, e, st);
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------- per-op entry
// One fused launch: activation quantisation + GEMV for m <= 4 rows (the hot decode path of ns_mul_mat / ns_mul_qkv /
// ns_ffn_silu).  act: device fp32 [m][lda].  Returns NS_E_UNSUPPORTED (without an error message) when the weights are
// not eligible, so the caller can fall back to act_prep + gemv kernels.
bool ns_decode_op_supported(const ns_weight* w) { return op_supported(w); }

int ns_launch_decode_op(const ns_weight* const* weights, int nw, int mode, const float* act, int lda, float* dst, int ldo, int m,
                        int m_total, const float* bias, int bias_bcast, const float* residual, float* aux, cudaStream_t st) {
  if (m < 1 || m > 4) return NS_E_INVALID;
  ProgOp op;
  if (int rc = fill_op(&op, m, weights, nw, mode, act, lda, dst, ldo, bias, bias_bcast, residual, aux, 0, true, m_total)) return rc;
  Plan pl;
  if (int rc = make_plan(&op, 1, m, &pl)) return rc;
  pl.cfg.pdl = 1;
  const ns_weight* w0 = weights[0];
  return launch_any(w0->comp, w0->asym, w0->stype, m, op, nullptr, 1, pl, nullptr, nullptr, st);
}

// ---------------------------------------------------------------------------------------------------- program API
struct ns_program {
  int m;
  int comp, stype, asym;
  bool finalized;
  std::vector<ProgOp> ops;
  ProgOp* d_ops;
  unsigned* d_counters;  // [nops] + epoch at [nops]
  Plan plan;
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
  if (!op_supported(w0)) {
    ns_set_error("ns_program: only 4-bit integer weights with integer activations and groups of 32..256 are supported");
    return NS_E_UNSUPPORTED;
  }
  if (p->comp < 0) {
    p->comp = w0->comp;
    p->stype = w0->stype;
    p->asym = w0->asym;
  } else if (w0->comp != p->comp || w0->stype != p->stype || w0->asym != p->asym) {
    ns_set_error("ns_program: all weights of a program must share scale type, symmetry and compute type");
    return NS_E_UNSUPPORTED;
  }
  ProgOp op;
  if (int rc = fill_op(&op, p->m, weights, nw, mode, in, lda, dst, ldo, bias, bias_bcast, residual, aux, barrier_before, false, p->m))
    return rc;
  for (int i = 0; i < nw; ++i) p->alg_bytes += ns_weight_algorithmic_bytes(weights[i]);
  p->ops.push_back(op);
  return NS_OK;
}

extern "C" size_t ns_program_algorithmic_bytes(const ns_program* p) { return p ? p->alg_bytes : 0; }

extern "C" int ns_program_finalize(ns_program* p, void* queue) {
  if (!p || p->ops.empty()) return NS_E_INVALID;
  if (p->finalized) return NS_OK;
  cudaStream_t st = ns_stream_of(queue);
  if (int rc = make_plan(p->ops.data(), (int)p->ops.size(), p->m, &p->plan)) return rc;
  const size_t nops = p->ops.size();
  NS_CUDA_TRY(cudaMalloc((void**)&p->d_ops, nops * sizeof(ProgOp)));
  NS_CUDA_TRY(cudaMalloc((void**)&p->d_counters, (nops + 1) * sizeof(unsigned)));
  NS_CUDA_TRY(cudaMemcpyAsync(p->d_ops, p->ops.data(), nops * sizeof(ProgOp), cudaMemcpyHostToDevice, st));
  NS_CUDA_TRY(cudaMemsetAsync(p->d_counters, 0, (nops + 1) * sizeof(unsigned), st));
  NS_CUDA_TRY(cudaStreamSynchronize(st));
  p->finalized = true;
  return NS_OK;
}

extern "C" int ns_program_run(ns_program* p, void* queue) {
  if (int rc = ns_ensure_device()) return rc;
  if (!p || !p->finalized) {
    ns_set_error("ns_program_run: program not finalized");
    return NS_E_INVALID;
  }
  const int nops = (int)p->ops.size();
  return launch_any(p->comp, p->asym, p->stype, p->m, p->ops[0], p->d_ops, nops, p->plan, p->d_counters, p->d_counters + nops,
                    ns_stream_of(queue));
}

extern "C" void ns_program_free(ns_program* p) {
  if (!p) return;
  if (p->d_ops) cudaFree(p->d_ops);
  if (p->d_counters) cudaFree(p->d_counters);
  delete p;
}
