// program.cu -- persistent multi-op decode kernel: a whole token's worth of weight-only matmuls in ONE launch, one CTA per SM.
//
// What it replaces: the reference rebuilds an ne graph per token and walks it node by node (ne_graph_compute,
// neural_speed/core/ne_layers.c:11915; llama graph, models/llama/llama.cpp:136-143,217-231,586,612-618,718); every matmul
// node first quantises its activations (NE_TASK_INIT, ne_layers.c:7143-7157) and then runs the dots.  On B200 a decode GEMV
// lasts 1.5-8 us, so a kernel boundary (drain + launch + refill of the load pipeline) costs as much as the work.  Here an
// "ns_program" is the list of matmul nodes of one token; one cooperative launch of 148 CTAs x (15 consumer warps + 1 producer
// warp) executes all of them:
//   * producer warp: walks the op list and streams this CTA's weight rows with cp.async.bulk (TMA 1-D, SASS UBLKCP) into a
//     ring of ~30 fixed-size slots (~190 KB) guarded by full/empty mbarriers.  A unit is a row pair (or one row when a pair
//     does not fit a slot: K = 11008).  LANES issue units in parallel, 15 per batch: one thread needs ~1000 cycles per unit
//     (a chain of ~100 dependent scalar instructions, try_wait, R2UR, UBLKCP -- measured with the per-unit trace below), which
//     capped a single-thread producer at 9-12 GB/s per SM against the 45 GB/s an SM's share of HBM needs.  The producer never
//     waits for activations: it runs ahead across op boundaries, so HBM stays busy while the consumers synchronise.
//   * consumers, per op: wait until every CTA has finished the previous op (one red.release on a per-op counter, one polling
//     thread per CTA -- measured 1.3 us per round on 148 CTAs, profiles/ubench.cu; per-CTA flags polled by 148 threads cost
//     4.8 us), optional RMSNorm + activation quantisation of the op's fp32 input into the shared-memory image
//     (norm_quant.cuh: Q8_0 / BesTLA u8 / s8, bit-exact), dp4a over their FIFO units, epilogue (bias / residual / SiLU*mul /
//     GELU).  For M == 1 and K <= 4096 the quantised activations live in REGISTERS (40 per lane): the inner loop then only
//     reads the weights from shared memory (LOP3 and IDP.4A both issue at half rate on sm_100 -- measured -- so instruction
//     count, not bandwidth, bounds how fast the consumers catch up after a barrier).
// Same integer arithmetic and fp32 summation order as gemv_ring.cu: results are bit-identical to the per-op kernels.
// Activations are read with ld.global.cg (L2) because another SM rewrites them between ops within the same launch.
// Roofline: HBM; algorithmic bytes per launch = sum over ops of N*K/2 + N*ceil(K/g)*(scale_bytes [+1 if asym]).
#include <algorithm>
#include <vector>

#include "norm_quant.cuh"
#include "nsb.cuh"

namespace {

constexpr int kConsumers = 15;  // 15 + the producer warp = 512 threads: 128 registers per thread (17 warps would round up to 20: 96)
constexpr int kConsumerThreads = kConsumers * 32;
constexpr int kThreads = kConsumerThreads + 32;
constexpr int kMaxSlots = 60;  // ring slots: a multiple of kConsumers, so slot s is only ever consumed by warp s % kConsumers
constexpr int kBatch = 15;    // units issued per producer-warp step (one lane each); <= nslots
constexpr int kKcReg = 4;  // 32-element chunks per lane the register path holds (K <= 4096)
constexpr int kTl = 8;     // timeline words per (op, CTA)
constexpr int kUnitTrace = 8192;  // debug: per-unit stamps of CTA 0 (first units of a launch)

struct ProgOp {
  const uint8_t* rows[3];
  int n[3];
  long long dst_off[3];
  int nw, mode;
  int k, kpad, pitch, sc_off, zp_off, cpg, group;
  uint32_t cpg_magic;
  const float* in;
  int lda;
  const int* in_index;  // optional: effective input = in + (*in_index) * in_stride (embedding row picked on the device)
  long long in_stride;
  const float* norm_w;  // optional RMSNorm weight applied to the input before quantisation
  float norm_eps;
  float* dst;
  int ldo;
  const float* bias;
  int bias_bcast;
  const float* residual;
  const int* res_index;  // optional: residual += (*res_index) * res_stride
  long long res_stride;
  float* aux;
  int eltop;
  int npairs;
  int barrier_before;
  int act_row, meta_off, meta_stride;
  int use_reg;  // 1: activations in registers (M == 1, nchunks <= 32 * kKcReg)
  int unit_rows;  // 2: a unit is a row pair; 1: single rows (a pair does not fit a ring slot)
  int nunits;
};

struct ProgCfg {
  int ring_off, slot_bytes, nslots;
  int bar_off;  // full[nslots] | empty[nslots]
  int m;
  int iters;  // the op list is executed `iters` times (tokens) inside one launch
  int batch;     // units issued per producer-warp step (one lane each), <= nslots
  int inflight;  // at most this many units issued and not yet landed (>= batch); bounds the depth of the SM's request queue
  int pf_units;  // L2 prefetch distance in units ahead of the load cursor (0: off)
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
// TMA L2 prefetch of a byte range (no completion tracking; SASS UBLKPF)
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
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
__device__ __forceinline__ void sts64u(uint32_t a, uint32_t x, uint32_t y) {
  asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(a), "r"(x), "r"(y) : "memory");
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
__device__ __forceinline__ float ldcg1f(const float* p) {
  float r;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ int ldcg1i(const int* p) {
  int r;
  asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long clk64() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// unit u of an op = one row pair (the two rows are adjacent in memory except in gate/up mode: gate row p, up row p)
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

// unit_rows == 1: unit u is row u of the concatenated weights
__device__ __forceinline__ PairSrc resolve_single(const ProgOp& P, int u) {
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

// One 32-element chunk of one row against one activation chunk: the exact integer sum (a - za)(u - off)
//   = sum a*u - off*Sa - za*(Su - 32*off)        (Sa = sum of the activation codes, Su = sum of the weight codes)
template <int AMODE>
__device__ __forceinline__ int chunk_dot(const uint4& wv, const uint4& a0, const uint4& a1, int off, int neg_off_sa, int za) {
  const uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w};
  uint32_t lo[4], hi[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    lo[i] = ww[i] & 0x0F0F0F0Fu;
    hi[i] = ww[i] & 0xF0F0F0F0u;  // high nibbles as bytes * 16 (no shift): exact, divided out after the dot
  }
  int pl = neg_off_sa, ph = 0;  // the chain starts from -off * Sa
  if (AMODE == A_U8) {
    pl = dp4a_uu(a0.x, lo[0], pl); ph = dp4a_uu(a0.y, hi[0], ph);
    pl = dp4a_uu(a0.z, lo[1], pl); ph = dp4a_uu(a0.w, hi[1], ph);
    pl = dp4a_uu(a1.x, lo[2], pl); ph = dp4a_uu(a1.y, hi[2], ph);
    pl = dp4a_uu(a1.z, lo[3], pl); ph = dp4a_uu(a1.w, hi[3], ph);
  } else {
    pl = dp4a_us(lo[0], (int)a0.x, pl); ph = dp4a_us(hi[0], (int)a0.y, ph);
    pl = dp4a_us(lo[1], (int)a0.z, pl); ph = dp4a_us(hi[1], (int)a0.w, ph);
    pl = dp4a_us(lo[2], (int)a1.x, pl); ph = dp4a_us(hi[2], (int)a1.y, ph);
    pl = dp4a_us(lo[3], (int)a1.z, pl); ph = dp4a_us(hi[3], (int)a1.w, ph);
  }
  int isum = pl + (ph >> 4);  // ph is an exact multiple of 16
  if (AMODE == A_U8) {
    int sl = 0, sh = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      sl = dp4a_uu(lo[i], 0x01010101u, sl);
      sh = dp4a_uu(hi[i], 0x01010101u, sh);
    }
    isum -= za * (sl + (sh >> 4) - 32 * off);
  }
  return isum;
}

// One unit (two weight rows) against register-resident activations: KC chunks per lane, same per-lane summation order as the
// shared-memory loop (chunk c = lane + 32 i in increasing i).
template <int KC, int AMODE, bool ASYM, int STYPE>
__device__ __forceinline__ void reg_unit(const uint32_t (&wb)[2], const uint32_t (&rb)[2], const uint4 (&A0)[kKcReg],
                                         const uint4 (&A1)[kKcReg], const float (&asc)[kKcReg], const int (&asa)[kKcReg],
                                         const int (&aza)[kKcReg], const uint32_t (&soff)[kKcReg], const uint32_t (&zoff)[kKcReg],
                                         float& acc0, float& acc1) {
  uint4 wv[KC][2];
  float ws[KC][2];
  int off[KC][2];
#pragma unroll
  for (int i = 0; i < KC; ++i)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      wv[i][r] = lds128(wb[r] + 512u * i);
      ws[i][r] = lds_scale<STYPE>(rb[r] + soff[i], 0);
      off[i][r] = 8;
      if (ASYM) off[i][r] += lds8s(rb[r] + zoff[i]);
    }
#pragma unroll
  for (int i = 0; i < KC; ++i) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int isum = ASYM ? chunk_dot<AMODE>(wv[i][r], A0[i], A1[i], off[i][r], -off[i][r] * asa[i], aza[i])
                            : chunk_dot<AMODE>(wv[i][r], A0[i], A1[i], 8, asa[i], aza[i]);
      float& acc = r ? acc1 : acc0;
      acc = fmaf((float)isum, asc[i] * ws[i][r], acc);
    }
  }
}

// One unit (NR weight rows) against the activation image in shared memory (any M, any K)
struct SmemUnitArgs {
  uint32_t r0, r1, smem_base, meta_s;
  int sc_off, zp_off, cpg;
  uint32_t cpg_magic;
  int act_row, meta_stride, nchunks, lane;
};
template <int NR, int M, int AMODE, bool ASYM, int STYPE>
__device__ __forceinline__ void smem_unit(const SmemUnitArgs& U, float (&acc)[2][M]) {
  const uint32_t rb[2] = {U.r0, U.r1};
#pragma unroll 2
  for (int c = U.lane; c < U.nchunks; c += 32) {
    const int gi = (U.cpg == 1) ? c : (int)__umulhi((uint32_t)c, U.cpg_magic);
    uint4 wv[NR];
    float ws[NR];
    int off[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      wv[r] = lds128(rb[r] + 16 * c);
      ws[r] = lds_scale<STYPE>(rb[r] + U.sc_off, gi);
      off[r] = 8;
      if (ASYM) off[r] += lds8s(rb[r] + U.zp_off + gi);
    }
    const uint32_t a_off = (uint32_t)(c >> 5) * 1024u + (uint32_t)(c & 31) * 16u;
#pragma unroll
    for (int m = 0; m < M; ++m) {
      const uint32_t ab = U.smem_base + (uint32_t)m * U.act_row + a_off;
      const uint4 a0 = lds128(ab), a1 = lds128(ab + 512);
      const uint2 mt = lds64(U.meta_s + 8u * (uint32_t)(m * U.meta_stride + c));
      const float a_scale = __uint_as_float(mt.x);
      const int sa = (int)(short)(mt.y & 0xffff);
      const int za = (int)((mt.y >> 16) & 0xff);
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const int isum = chunk_dot<AMODE>(wv[r], a0, a1, off[r], -off[r] * sa, za);
        acc[r][m] = fmaf((float)isum, a_scale * ws[r], acc[r][m]);
      }
    }
  }
}

// descriptor global -> shared by the consumer threads 32 .. 32 + words
__device__ __forceinline__ void copy_op(ProgOp* dst, const ProgOp* src, int tid) {
  constexpr int kWords = (int)(sizeof(ProgOp) / 4);
  static_assert(kWords <= kConsumerThreads - 32, "descriptor copy uses threads 32..");
  if (tid >= 32 && tid < 32 + kWords) reinterpret_cast<uint32_t*>(dst)[tid - 32] = reinterpret_cast<const uint32_t*>(src)[tid - 32];
}

template <int COMP, int M, bool ASYM, int STYPE>
__global__ void __launch_bounds__(kThreads, 1)
    program_kernel(const ProgOp* __restrict__ ops, int nops, const ProgCfg R, unsigned* __restrict__ counters,
                   unsigned* __restrict__ epoch_ptr, unsigned long long* __restrict__ tl, unsigned long long* __restrict__ tu_dbg) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ ProgOp op_s[2];  // the consumers' current and next op (the next one is fetched during the current one)
  __shared__ ProgOp op_p;     // the producer's current op
  __shared__ ProgOp op_q;     // the op under the producer's L2-prefetch cursor
  __shared__ float red_s[kConsumers];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t ring = smem_base + R.ring_off;
  const uint32_t full0 = smem_base + R.bar_off;
  const uint32_t empty0 = full0 + 8u * (uint32_t)R.nslots;
  const int NS = R.nslots;
  const int first = blockIdx.x, G = (int)gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // launches completed so far: every CTA arrives once per op per launch on counters[op], so after `epoch` launches of `iters`
  // iterations each counter stands at (sum of earlier iterations) * G; epoch_ptr[0] holds that sum of iterations.
  const unsigned base_iters = ld_acquire(epoch_ptr);

  if (warp == kConsumers) {
    // ===================== producer: streams the weights of ALL ops, never waits for activations =====================
    // L1 is all but carved away by the 224 KB of shared memory, so every read of the descriptor array costs an L2 round trip:
    // the warp copies the op's descriptor to shared memory once per op.  Lane l < kBatch issues unit jb + l of each batch.
    int ubase = 0;  // CTA-wide index of the op's first unit
    constexpr int kWords = (int)(sizeof(ProgOp) / 4);
    // L2 prefetch cursor: runs pf_units units ahead of the load cursor, across op boundaries.  While the ring is full (the
    // consumers are in a grid barrier) the prefetches already issued keep HBM streaming into L2; the ring refills from L2.
    int pf_seq = 0, pf_j = 0, pf_my = 0, pf_ahead = 0;
    bool pf_done = R.pf_units <= 0;
    const int total_ops_p = nops * R.iters;
    auto pf_open = [&]() {  // descriptor of op pf_seq under the prefetch cursor
      __syncwarp();
      for (int w = lane; w < kWords; w += 32)
        reinterpret_cast<uint32_t*>(&op_q)[w] = reinterpret_cast<const uint32_t*>(ops + pf_seq % nops)[w];
      __syncwarp();
      const int nu = op_q.nunits;
      pf_my = first < nu ? (nu - first + G - 1) / G : 0;
      pf_j = 0;
    };
    if (!pf_done) pf_open();
    auto pf_top_up = [&]() {
      while (!pf_done && pf_ahead < R.pf_units) {
        const int j = pf_j + lane;
        if (j < pf_my) {
          const PairSrc ps = op_q.unit_rows == 2 ? resolve_pair(op_q, first + j * G) : resolve_single(op_q, first + j * G);
          const uint32_t pq = (uint32_t)op_q.pitch;
          if (ps.valid1 && ps.r1 != ps.r0 + pq) {
            bulk_prefetch_l2(ps.r0, pq);
            bulk_prefetch_l2(ps.r1, pq);
          } else {
            bulk_prefetch_l2(ps.r0, (ps.valid1 ? 2u : 1u) * pq);
          }
        }
        const int n = min(32, pf_my - pf_j);
        pf_j += n;
        pf_ahead += n;
        if (pf_j >= pf_my) {
          if (++pf_seq >= total_ops_p) pf_done = true;
          else pf_open();
        }
      }
    };
    for (int it = 0; it < R.iters; ++it) {
      for (int oi = 0; oi < nops; ++oi) {
        __syncwarp();
        for (int w = lane; w < kWords; w += 32)
          reinterpret_cast<uint32_t*>(&op_p)[w] = reinterpret_cast<const uint32_t*>(ops + oi)[w];
        __syncwarp();
        const int nunits = op_p.nunits;
        const uint32_t pitch = (uint32_t)op_p.pitch;
        const int my_units = first < nunits ? (nunits - first + G - 1) / G : 0;
        const size_t tli = ((size_t)(it * nops + oi) * G + first) * kTl;
        if (tl && lane == 0) tl[tli + 5] = clk64();
        for (int jb = 0; jb < my_units; jb += R.batch) {
          pf_top_up();
          pf_ahead -= min(R.batch, my_units - jb);
          const int j = jb + lane;
          if (lane < R.batch && j < my_units) {
            const int i = ubase + j;
            const bool trace = tu_dbg && first == 0 && i < kUnitTrace;
            if (trace) tu_dbg[(size_t)i * 8 + 0] = clk64();
            const PairSrc ps = op_p.unit_rows == 2 ? resolve_pair(op_p, first + j * G) : resolve_single(op_p, first + j * G);
            const int slot = i % NS, lap = i / NS;
            if (lap > 0) mbar_wait(empty0 + 8u * slot, (uint32_t)(lap - 1) & 1u);  // the slot's previous unit has been consumed
            if (i >= R.inflight) {  // unit i - inflight (issued by an earlier batch) has landed
              const int pi = i - R.inflight;
              mbar_wait(full0 + 8u * (pi % NS), (uint32_t)(pi / NS) & 1u);
            }
            if (trace) tu_dbg[(size_t)i * 8 + 1] = clk64();
            const uint32_t dst = ring + (uint32_t)slot * (uint32_t)R.slot_bytes;
            const uint32_t bar = full0 + 8u * slot;
            if (ps.valid1 && ps.r1 != ps.r0 + pitch) {  // gate row, up row
              mbar_expect_tx(bar, 2u * pitch);
              bulk_g2s(dst, ps.r0, pitch, bar);
              bulk_g2s(dst + pitch, ps.r1, pitch, bar);
            } else {
              const uint32_t bytes = (ps.valid1 ? 2u : 1u) * pitch;
              mbar_expect_tx(bar, bytes);
              bulk_g2s(dst, ps.r0, bytes, bar);
            }
            if (trace) tu_dbg[(size_t)i * 8 + 2] = clk64();
          }
          __syncwarp();
        }
        ubase += my_units;
        if (tl && lane == 0) tl[tli + 6] = clk64();
      }
    }
    return;
  }

  // ===================== consumers =====================
  constexpr int AMODE = (COMP == NS_COMP_INT8) ? A_U8 : A_S8;
  const int total_ops = nops * R.iters;
  int ubase = 0;  // CTA-wide index of this op's first unit
  copy_op(&op_s[0], ops, threadIdx.x);
  nsq::bar_sync<1, kConsumerThreads>();
  for (int seq = 0; seq < total_ops; ++seq) {
    const int it = seq / nops, oi = seq - it * nops;
    const size_t tli = ((size_t)seq * G + first) * kTl;
    // ---- op boundary: wait until every CTA has published the previous op (the descriptor is already in shared memory) ----
    {
      if (threadIdx.x == 0) {
        if (tl) {
          tl[tli + 0] = clk64();
          tl[tli + 7] = gtimer();
        }
        if (seq > 0 && op_s[seq & 1].barrier_before) {
          // the previous op in execution order: (it, oi - 1) or (it - 1, nops - 1)
          const int po = oi > 0 ? oi - 1 : nops - 1;
          const unsigned want = (base_iters + (unsigned)(oi > 0 ? it : it - 1) + 1u) * (unsigned)G;
          while ((int)(ld_acquire(counters + po) - want) < 0) {
          }
        }
        if (tl) tl[tli + 1] = clk64();
      }
    }
    nsq::bar_sync<1, kConsumerThreads>();
    const ProgOp& P = op_s[seq & 1];
    if (seq + 1 < total_ops) copy_op(&op_s[(seq + 1) & 1], ops + (seq + 1) % nops, threadIdx.x);  // read again only after >= 2 barriers
    {
      const float* in = P.in;
      if (P.in_index) in += (long long)ldcg1i(P.in_index) * P.in_stride;
      const nsq::NormQuantIn qi{in, P.norm_w, P.norm_eps, P.lda, P.k, P.kpad, COMP == NS_COMP_Q8_0 ? 32 : P.group,
                                P.act_row, P.meta_off, P.meta_stride};
      nsq::norm_quantise_to_smem<COMP, kConsumerThreads, 1>(qi, R.m, smem_base, red_s, threadIdx.x);
    }
    nsq::bar_sync<1, kConsumerThreads>();
    if (tl && threadIdx.x == 0) tl[tli + 2] = clk64();

    // hot op fields into registers (op_s sits in shared memory; the inline-asm loads below would otherwise re-read it)
    const int nunits = P.nunits, unit_rows = P.unit_rows, pitch = P.pitch, sc_off = P.sc_off, zp_off = P.zp_off, cpg = P.cpg, mode = P.mode;
    const int act_row = P.act_row, meta_stride = P.meta_stride, ldo = P.ldo, eltop = P.eltop, bias_bcast = P.bias_bcast;
    const uint32_t cpg_magic = P.cpg_magic;
    const float* bias = P.bias;
    float* dst = P.dst;
    float* aux = P.aux;
    const int my_units = first < nunits ? (nunits - first + G - 1) / G : 0;
    const uint32_t meta_s = smem_base + P.meta_off;
    const int nchunks = P.kpad >> 5;
    const float* residual = P.residual;
    if (residual && P.res_index) residual += (long long)ldcg1i(P.res_index) * P.res_stride;

    // register-resident activations (M == 1, nchunks a multiple of 32): chunk c = lane + 32 i, i < kc
    uint4 A0[kKcReg], A1[kKcReg];
    float asc[kKcReg];
    int asa[kKcReg], aza[kKcReg];
    uint32_t soff[kKcReg], zoff[kKcReg];  // byte offsets of the chunk's scale / zero point inside a weight row
    const bool use_reg = (M == 1) && P.use_reg;
    const int kc = nchunks >> 5;
    if (use_reg) {
#pragma unroll
      for (int i = 0; i < kKcReg; ++i) {
        const int c = lane + 32 * i;
        if (i < kc) {
          const uint32_t ab = smem_base + (uint32_t)(c >> 5) * 1024u + (uint32_t)(c & 31) * 16u;
          A0[i] = lds128(ab);
          A1[i] = lds128(ab + 512);
          const uint2 mt = lds64(meta_s + 8u * (uint32_t)c);
          asc[i] = __uint_as_float(mt.x);
          asa[i] = (int)(short)(mt.y & 0xffff);
          if (!ASYM) asa[i] *= -8;  // the dp4a chain starts from -off * Sa; off == 8 for symmetric weights
          aza[i] = (int)((mt.y >> 16) & 0xff);
          const int gi = (cpg == 1) ? c : (int)__umulhi((uint32_t)c, cpg_magic);
          soff[i] = (uint32_t)sc_off + (uint32_t)gi * (STYPE == NS_S_F32 ? 4u : 2u);
          zoff[i] = (uint32_t)zp_off + (uint32_t)gi;
        } else {
          A0[i] = A1[i] = make_uint4(0, 0, 0, 0);
          asc[i] = 0.f;
          asa[i] = aza[i] = 0;
          soff[i] = zoff[i] = 0;
        }
      }
    }

    // this warp's units of the op: CTA-wide unit index ubase + j with (ubase + j) % kConsumers == warp
    int j0 = warp - (ubase % kConsumers);
    if (j0 < 0) j0 += kConsumers;
    for (int j = j0; j < my_units; j += kConsumers) {
      const int ui = ubase + j;
      const int d = ui % NS;
      const PairSrc ps = unit_rows == 2 ? resolve_pair(P, first + j * G) : resolve_single(P, first + j * G);
      // epilogue operands fetched before the wait (their L2 latency hides behind the dot products)
      float ep_res = 0.f, ep_bias = 0.f;
      const bool gate_up = mode == NS_GEMV_GATE_UP_SILU;
      const int e_m = gate_up ? lane : (lane >> 1);
      const int e_r = gate_up ? 0 : (lane & 1);
      const bool e_live = e_m < R.m && e_m < M && (e_r == 0 || ps.valid1);
      const long long e_out = e_r ? ps.out1 : ps.out0;
      if (e_live && !gate_up) {
        const size_t o = (size_t)e_m * ldo + e_out;
        if (bias) ep_bias = bias_bcast ? ldcg1f(bias + e_out) : ldcg1f(bias + o);
        if (residual) ep_res = ldcg1f(residual + o);
      }
      const bool trace = tu_dbg && first == 0 && ui < kUnitTrace && lane == 0;
      if (trace) tu_dbg[(size_t)ui * 8 + 3] = clk64();
      mbar_wait(full0 + 8u * d, (uint32_t)(ui / NS) & 1u);
      if (trace) tu_dbg[(size_t)ui * 8 + 4] = clk64();
      const uint32_t r0 = ring + (uint32_t)d * (uint32_t)R.slot_bytes;
      const uint32_t r1 = ps.valid1 ? r0 + pitch : r0;
      const uint32_t rb[2] = {r0, r1};
      float acc[2][M];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int m = 0; m < M; ++m) acc[r][m] = 0.f;

      if (use_reg) {
        const uint32_t wb[2] = {r0 + 16u * lane, r1 + 16u * lane};
        switch (kc) {  // straight-line code per row length: the loads of all chunks are in flight before the first dot
          case 4: reg_unit<4, AMODE, ASYM, STYPE>(wb, rb, A0, A1, asc, asa, aza, soff, zoff, acc[0][0], acc[1][0]); break;
          case 3: reg_unit<3, AMODE, ASYM, STYPE>(wb, rb, A0, A1, asc, asa, aza, soff, zoff, acc[0][0], acc[1][0]); break;
          case 2: reg_unit<2, AMODE, ASYM, STYPE>(wb, rb, A0, A1, asc, asa, aza, soff, zoff, acc[0][0], acc[1][0]); break;
          default: reg_unit<1, AMODE, ASYM, STYPE>(wb, rb, A0, A1, asc, asa, aza, soff, zoff, acc[0][0], acc[1][0]); break;
        }
      } else {
        const SmemUnitArgs ua{rb[0], rb[1], smem_base, meta_s, sc_off, zp_off, cpg, cpg_magic, act_row, meta_stride, nchunks, lane};
        if (unit_rows == 2) smem_unit<2, M, AMODE, ASYM, STYPE>(ua, acc);
        else smem_unit<1, M, AMODE, ASYM, STYPE>(ua, acc);
      }
      __syncwarp();
      if (trace) tu_dbg[(size_t)ui * 8 + 5] = clk64();
      if (lane == 0) mbar_arrive(empty0 + 8u * d);  // the unit's bytes may be overwritten

#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int m = 0; m < M; ++m) acc[r][m] = warp_sum(acc[r][m]);
      if (gate_up) {
        float g = 0.f, up = 0.f;
#pragma unroll
        for (int m = 0; m < M; ++m)
          if (lane == m) {
            g = acc[0][m];
            up = acc[1][m];
          }
        if (e_live) {
          const float sg = eltop == NS_ELT_GELU ? ns_gelu(g) : ns_silu(g);  // kernel_ref.h:1569-1576
          if (aux) aux[(size_t)e_m * ldo + ps.out0] = sg;
          dst[(size_t)e_m * ldo + ps.out0] = sg * up;
        }
      } else {
        float v = 0.f;
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int m = 0; m < M; ++m)
            if (lane == 2 * m + r) v = acc[r][m];
        if (e_live) {
          if (bias) v += ep_bias;
          if (eltop == NS_ELT_GELU) v = ns_gelu(v);
          if (residual) v += ep_res;
          dst[(size_t)e_m * ldo + e_out] = v;
        }
      }
    }
    ubase += my_units;
    // ---- op done in this CTA: publish (release) ----
    nsq::bar_sync<1, kConsumerThreads>();
    if (threadIdx.x == 0) {
      if (tl) tl[tli + 3] = clk64();
      red_release_add(counters + oi, 1u);
      if (tl) tl[tli + 4] = clk64();
    }
  }
  // last op finished everywhere -> advance the epoch exactly once (block 0), so the next launch sees fresh targets
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const unsigned want = (base_iters + (unsigned)R.iters) * (unsigned)G;
    while ((int)(ld_acquire(counters + (nops - 1)) - want) < 0) {
    }
    red_release_add(epoch_ptr, (unsigned)R.iters);
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
  unsigned* d_counters;      // [nops] arrivals per op + iterations completed (epoch) at [nops]
  unsigned long long* d_tl;  // debug timeline [nops][grid][kTl] (NS_PROG_TIMELINE), else NULL
  unsigned long long* d_tu;  // debug per-unit trace of CTA 0 [kUnitTrace][8]
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
  p->d_tl = nullptr;
  p->d_tu = nullptr;
  p->alg_bytes = 0;
  return p;
}

extern "C" int ns_program_add_matmul_ex(ns_program* p, const ns_weight* const* weights, int nw, int mode, const float* in, int lda,
                                        float* dst, int ldo, const float* bias, int bias_bcast, const float* residual, float* aux,
                                        int barrier_before, const float* norm_w, float norm_eps, const int* in_index,
                                        long long in_stride, const int* res_index, long long res_stride, int eltop) {
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
  if (norm_w && w0->k % 4 != 0) {
    ns_set_error("ns_program: the fused RMSNorm needs k %% 4 == 0");
    return NS_E_UNSUPPORTED;
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
  op.in_index = in_index;
  op.in_stride = in_stride;
  op.norm_w = norm_w;
  op.norm_eps = norm_eps;
  op.dst = dst;
  op.ldo = ldo;
  op.bias = bias;
  op.bias_bcast = bias_bcast;
  op.residual = residual;
  op.res_index = res_index;
  op.res_stride = res_stride;
  op.aux = aux;
  op.eltop = eltop;
  op.npairs = (mode == NS_GEMV_GATE_UP_SILU) ? w0->n : (int)((ntot + 1) / 2);
  op.barrier_before = barrier_before;
  op.act_row = (int)ns_round_up((size_t)w0->kpad, 1024);
  op.meta_stride = ns_meta_stride(w0->kpad);
  op.meta_off = p->m * op.act_row;
  static const bool no_reg = getenv("NS_PROG_NO_REG") != nullptr;  // tuning aid
  op.use_reg = (p->m == 1 && (w0->kpad >> 5) <= 32 * kKcReg && (w0->kpad >> 5) % 32 == 0 && !no_reg) ? 1 : 0;
  p->ops.push_back(op);
  return NS_OK;
}

extern "C" int ns_program_add_matmul(ns_program* p, const ns_weight* const* weights, int nw, int mode, const float* in, int lda,
                                     float* dst, int ldo, const float* bias, int bias_bcast, const float* residual,
                                     float* aux, int barrier_before) {
  return ns_program_add_matmul_ex(p, weights, nw, mode, in, lda, dst, ldo, bias, bias_bcast, residual, aux, barrier_before,
                                  nullptr, 0.f, nullptr, 0, nullptr, 0, NS_ELT_DEFAULT);
}

extern "C" size_t ns_program_algorithmic_bytes(const ns_program* p) { return p ? p->alg_bytes : 0; }

extern "C" int ns_program_finalize(ns_program* p, void* queue) {
  if (!p || p->ops.empty()) return NS_E_INVALID;
  if (p->finalized) return NS_OK;
  cudaStream_t st = ns_stream_of(queue);
  const int mt = p->m >= 3 ? 4 : p->m;
  size_t act_region = 0;
  int unit_max = 0;
  for (const ProgOp& o : p->ops) {
    act_region = std::max(act_region, ns_round_up((size_t)mt * o.act_row + (size_t)mt * o.meta_stride * 8, 128));
    unit_max = std::max(unit_max, 2 * o.pitch);
  }
  static const int env_kb = getenv("NS_PROG_SMEM_KB") ? atoi(getenv("NS_PROG_SMEM_KB")) : 0;  // tuning aid
  const size_t budget = (size_t)(env_kb > 0 ? env_kb : 222) * 1024;  // the static shared memory (descriptors, reductions) rides on top
  // slot = the longest row, or a gate/up pair (both rows of such a unit feed one epilogue); ops whose row pair fits use pairs
  int slot = 0;
  for (const ProgOp& o : p->ops) slot = std::max(slot, o.mode == NS_GEMV_GATE_UP_SILU ? 2 * o.pitch : o.pitch);
  static const int env_pair = getenv("NS_PROG_PAIR_SLOTS") ? atoi(getenv("NS_PROG_PAIR_SLOTS")) : 0;  // tuning aid: slots of a pair of the longest rows
  if (env_pair) slot = unit_max;
  slot = (int)ns_round_up((size_t)slot, 128);
  int nslots = 0;
  if (budget > act_region + 64) nslots = (int)((budget - act_region - 64) / ((size_t)slot + 16));
  nslots -= nslots % kConsumers;
  if (nslots > kMaxSlots) nslots = kMaxSlots;
  if (nslots < kConsumers) {  // kBatch <= nslots: the units of one producer batch land in distinct slots
    ns_set_error("ns_program: rows too long for the shared-memory ring (%d B per slot)", slot);
    return NS_E_UNSUPPORTED;
  }
  for (ProgOp& o : p->ops) {
    o.unit_rows = (o.mode == NS_GEMV_GATE_UP_SILU || 2 * o.pitch <= slot) ? 2 : 1;
    long long rows = 0;
    for (int i = 0; i < o.nw; ++i) rows += o.n[i];
    o.nunits = o.unit_rows == 2 ? o.npairs : (int)rows;
  }
  p->cfg.ring_off = (int)act_region;
  p->cfg.slot_bytes = slot;
  p->cfg.nslots = nslots;
  p->cfg.bar_off = (int)(act_region + (size_t)nslots * slot);
  p->cfg.m = p->m;
  p->cfg.iters = 1;
  static const int env_batch = getenv("NS_PROG_BATCH") ? atoi(getenv("NS_PROG_BATCH")) : 0;        // tuning aids
  static const int env_inflight = getenv("NS_PROG_INFLIGHT") ? atoi(getenv("NS_PROG_INFLIGHT")) : 0;
  p->cfg.batch = std::max(1, std::min(std::min(32, nslots), env_batch > 0 ? env_batch : kBatch));
  p->cfg.inflight = std::max(p->cfg.batch, env_inflight > 0 ? env_inflight : nslots);
  static const int env_pf = getenv("NS_PROG_PF_UNITS") ? atoi(getenv("NS_PROG_PF_UNITS")) : -1;
  p->cfg.pf_units = env_pf >= 0 ? env_pf : 0;
  p->smem = act_region + (size_t)nslots * slot + (size_t)nslots * 16;
  p->grid = ns_num_sms();
  const size_t nops = p->ops.size();
  NS_CUDA_TRY(cudaMalloc((void**)&p->d_ops, nops * sizeof(ProgOp)));
  NS_CUDA_TRY(cudaMalloc((void**)&p->d_counters, (nops + 1) * sizeof(unsigned)));
  NS_CUDA_TRY(cudaMemcpyAsync(p->d_ops, p->ops.data(), nops * sizeof(ProgOp), cudaMemcpyHostToDevice, st));
  NS_CUDA_TRY(cudaMemsetAsync(p->d_counters, 0, (nops + 1) * sizeof(unsigned), st));
  if (getenv("NS_PROG_TIMELINE")) {
    const size_t words = nops * (size_t)p->grid * kTl;
    NS_CUDA_TRY(cudaMalloc((void**)&p->d_tl, words * sizeof(unsigned long long)));
    NS_CUDA_TRY(cudaMemsetAsync(p->d_tl, 0, words * sizeof(unsigned long long), st));
    NS_CUDA_TRY(cudaMalloc((void**)&p->d_tu, (size_t)kUnitTrace * 8 * sizeof(unsigned long long)));
    NS_CUDA_TRY(cudaMemsetAsync(p->d_tu, 0, (size_t)kUnitTrace * 8 * sizeof(unsigned long long), st));
  }
  NS_CUDA_TRY(cudaStreamSynchronize(st));
  p->finalized = true;
  return NS_OK;
}

template <int COMP, int M, bool ASYM, int STYPE>
static int run_one(ns_program* p, int iters, cudaStream_t st) {
  auto kern = program_kernel<COMP, M, ASYM, STYPE>;
  static bool attr_set = false;
  if (!attr_set) {
    NS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
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
  ProgCfg c = p->cfg;
  c.iters = iters;
  unsigned long long* tl = (p->d_tl && iters == 1) ? p->d_tl : nullptr;
  unsigned long long* tu = tl ? p->d_tu : nullptr;
  NS_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, ops, nops, c, counters, epoch, tl, tu));
  ns_count_launch();
  return NS_OK;
}
template <int COMP, bool ASYM, int STYPE>
static int run_m(ns_program* p, int iters, cudaStream_t st) {
  switch (p->m) {
    case 1: return run_one<COMP, 1, ASYM, STYPE>(p, iters, st);
    case 2: return run_one<COMP, 2, ASYM, STYPE>(p, iters, st);
    default: return run_one<COMP, 4, ASYM, STYPE>(p, iters, st);
  }
}
template <int COMP, bool ASYM>
static int run_s(ns_program* p, int iters, cudaStream_t st) {
  switch (p->stype) {
    case NS_S_F32: return run_m<COMP, ASYM, NS_S_F32>(p, iters, st);
    case NS_S_F16: return run_m<COMP, ASYM, NS_S_F16>(p, iters, st);
    default: return run_m<COMP, ASYM, NS_S_BF16>(p, iters, st);
  }
}
template <int COMP>
static int run_a(ns_program* p, int iters, cudaStream_t st) {
  return p->asym ? run_s<COMP, true>(p, iters, st) : run_s<COMP, false>(p, iters, st);
}

// the op list `iters` times inside ONE launch (iters tokens of a generation loop)
extern "C" int ns_program_run_n(ns_program* p, int iters, void* queue) {
  if (int rc = ns_ensure_device()) return rc;
  if (!p || !p->finalized || iters < 1) {
    ns_set_error("ns_program_run: program not finalized");
    return NS_E_INVALID;
  }
  cudaStream_t st = ns_stream_of(queue);
  switch (p->comp) {
    case NS_COMP_Q8_0: return run_a<NS_COMP_Q8_0>(p, iters, st);
    case NS_COMP_INT8: return run_a<NS_COMP_INT8>(p, iters, st);
    default: return run_a<NS_COMP_INT8_S8>(p, iters, st);
  }
}

extern "C" int ns_program_run(ns_program* p, void* queue) { return ns_program_run_n(p, 1, queue); }

// debug: copies the [nops][grid][8] clock stamps of the last run to `host` (needs NS_PROG_TIMELINE at finalize)
extern "C" int ns_program_timeline(ns_program* p, unsigned long long* host, size_t cap_words, int* nops, int* grid) {
  if (!p || !p->d_tl) return NS_E_INVALID;
  const size_t n = p->ops.size() * (size_t)p->grid * kTl;
  if (nops) *nops = (int)p->ops.size();
  if (grid) *grid = p->grid;
  if (cap_words < n) return NS_E_INVALID;
  NS_CUDA_TRY(cudaDeviceSynchronize());
  NS_CUDA_TRY(cudaMemcpy(host, p->d_tl, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return NS_OK;
}

// debug: per-unit stamps of CTA 0: [unit][8] = producer {alloc start, space ok, issued}, consumer {wait start, data ready, done}
extern "C" int ns_program_unit_trace(ns_program* p, unsigned long long* host, size_t cap_words) {
  if (!p || !p->d_tu || cap_words < (size_t)kUnitTrace * 8) return NS_E_INVALID;
  NS_CUDA_TRY(cudaDeviceSynchronize());
  NS_CUDA_TRY(cudaMemcpy(host, p->d_tu, (size_t)kUnitTrace * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return NS_OK;
}

extern "C" void ns_program_free(ns_program* p) {
  if (!p) return;
  if (p->d_tu) cudaFree(p->d_tu);
  if (p->d_tl) cudaFree(p->d_tl);
  if (p->d_ops) cudaFree(p->d_ops);
  if (p->d_counters) cudaFree(p->d_counters);
  delete p;
}
