// program.cu -- persistent multi-op decode kernel: a whole token's worth of weight-only matmuls in ONE launch, one CTA per SM.
//
// What it replaces: the reference rebuilds an ne graph per token and walks it node by node (ne_graph_compute,
// neural_speed/core/ne_layers.c:11915; llama graph, models/llama/llama.cpp:136-143,217-231,586,612-618,718); every matmul
// node first quantises its activations (NE_TASK_INIT, ne_layers.c:7143-7157) and then runs the dots.  On B200 a decode GEMV
// lasts 1.5-8 us, so a kernel boundary (drain + launch + refill of the load pipeline) costs as much as the work.  Here an
// "ns_program" is the list of matmul nodes of one token; one cooperative launch of 148 CTAs x (15 consumer warps + 1 producer
// warp) executes all of them:
//   * producer warp: walks the op list and streams this CTA's weight rows with cp.async.bulk (TMA 1-D, SASS UBLKCP) into a
//     ring of 8 slots (~200 KB) guarded by full/empty mbarriers.  A CTA owns a CONTIGUOUS range of an op's rows; a unit is a
//     row pair (what one consumer warp computes at a time), a GROUP of up to 8 units is one contiguous byte range = one bulk
//     copy = one ring slot.  Lanes issue groups in parallel: a thread needs ~1000 cycles per copy (a chain of ~100 dependent
//     scalar instructions around try_wait, R2UR, UBLKCP -- measured with the per-unit trace), which capped a producer issuing
//     one 4.6 KB pair per copy at 9-12 GB/s per SM against the 45 GB/s an SM's share of HBM needs.  The producer never waits
//     for activations: it runs ahead across op boundaries, so HBM stays busy while the consumers synchronise.
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
constexpr int kMaxGs = 8;   // units per group (= arrivals on a slot's empty barrier) at most
constexpr int kKcReg = 4;  // 32-element chunks per lane the register path holds (K <= 4096)
constexpr int kTl = 12;    // timeline words per (op, CTA)
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
  int unit_rows;  // 2: a unit is a row pair; 1: single rows
  int nunits;
  int gs;  // units per group (one bulk copy, one ring slot)
  // flag-in-data hand-over between ops (NCCL's LL idea): the op also writes every output as an 8-byte word {value, tag} to
  // dst_tag [m][ldo] (tag = ops executed so far, one 64-bit store); an op with in_tagged reads `in` as such words and polls each
  // one until its tag matches -- no counter, no fence, no grid barrier between the two ops
  unsigned long long* dst_tag;
  int in_tagged;
  int publish;  // 1: a later op (or the end of the launch) waits on this op's arrival counter
  // divisions done once on the host: kConsumers / gs, kConsumers % gs, ceil(2^32 / gs), 32 / cpg
  int dgi, dwi, cstep;
  uint32_t gs_magic;
};

struct ProgCfg {
  int ring_off, group_bytes, ngroups;  // ring of `ngroups` slots, each holding one group of up to gs units
  int gs_max;   // arrival count of the empty barriers (= the largest gs of any op)
  int bar_off;  // full[ngroups] | empty[ngroups]
  int m;
  int iters;  // the op list is executed `iters` times (tokens) inside one launch
  int batch;    // groups issued per producer-warp step (one lane each), <= ngroups
  uint32_t g_magic;  // ceil(2^32 / gridDim.x): first * nunits / G without a 64-bit division (exact while first * nunits < 2^32 / G)
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
// non-blocking probe of a phase
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
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
// shared-memory loop (chunk c = lane + 32 i in increasing i).  Two chunks (four 16-byte weight loads) are in flight at a time:
// with 44 registers pinned by the activations, eight loads in flight spilled -- and with the L1 carved away a spill costs an
// L2 round trip.
template <int KC, int AMODE, bool ASYM, int STYPE>
__device__ __forceinline__ void reg_unit(const uint32_t (&wb)[2], const uint32_t (&rb)[2], const uint4 (&A0)[kKcReg],
                                         const uint4 (&A1)[kKcReg], const float (&asc)[kKcReg], const int (&asa)[kKcReg],
                                         const int (&aza)[kKcReg], uint32_t soff0, uint32_t sstep, uint32_t zoff0, uint32_t zstep,
                                         float& acc0, float& acc1) {
#pragma unroll
  for (int i0 = 0; i0 < KC; i0 += 2) {
    uint4 wv[2][2];
    float ws[2][2];
    int off[2][2];
#pragma unroll
    for (int ii = 0; ii < 2; ++ii) {
      const int i = i0 + ii;
      if (i < KC) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          wv[ii][r] = lds128(wb[r] + 512u * i);
          ws[ii][r] = lds_scale<STYPE>(rb[r] + soff0 + sstep * i, 0);
          off[ii][r] = 8;
          if (ASYM) off[ii][r] += lds8s(rb[r] + zoff0 + zstep * i);
        }
      }
    }
#pragma unroll
    for (int ii = 0; ii < 2; ++ii) {
      const int i = i0 + ii;
      if (i < KC) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int isum = ASYM ? chunk_dot<AMODE>(wv[ii][r], A0[i], A1[i], off[ii][r], -off[ii][r] * asa[i], aza[i])
                                : chunk_dot<AMODE>(wv[ii][r], A0[i], A1[i], 8, asa[i], aza[i]);
          float& acc = r ? acc1 : acc0;
          acc = fmaf((float)isum, asc[i] * ws[ii][r], acc);
        }
      }
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

// Where unit u of an op lands in its output, and whether its second row exists (unit_rows == 2: rows 2u, 2u + 1 of the
// concatenated weights -- pairs never straddle two weights, the launcher checks even n; gate/up: gate row u, up row u)
struct UnitOut {
  int out0;
  bool valid1;
};
// (the op's fields are read from its descriptor in shared memory: a handful of LDS per unit instead of nine pinned registers;
// outputs are indexed with ints: m * ldo + n < 2^31)
__device__ __forceinline__ UnitOut unit_out(const ProgOp& g, int u) {
  UnitOut r;
  if (g.mode == NS_GEMV_GATE_UP_SILU) {
    r.out0 = u;
    r.valid1 = true;
    return r;
  }
  const int unit_rows = g.unit_rows, nw = g.nw;
  int row = unit_rows * u, nn = g.n[0];
  int off = (int)g.dst_off[0];
  if (nw > 1 && row >= nn) {
    row -= nn;
    nn = g.n[1];
    off = (int)g.dst_off[1];
    if (nw > 2 && row >= nn) {
      row -= nn;
      nn = g.n[2];
      off = (int)g.dst_off[2];
    }
  }
  r.out0 = off + row;
  r.valid1 = unit_rows == 2 && row + 1 < nn;
  return r;
}
// both sums of a unit with six shuffles: after the xor-16 step lanes < 16 carry row 0, lanes >= 16 row 1 (every lane of a
// butterfly holds the same bits at every step, so lane 0 / lane 16 end with exactly what two full butterflies give)
__device__ __forceinline__ void warp_sum2(float& a, float& b, int lane) {
  a += __shfl_xor_sync(0xffffffffu, a, 16);
  b += __shfl_xor_sync(0xffffffffu, b, 16);
  float v = lane < 16 ? a : b;
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  a = __shfl_sync(0xffffffffu, v, 0);
  b = __shfl_sync(0xffffffffu, v, 16);
}

// units [lo, hi) of an op owned by CTA `first` of G: floor(first * nunits / G); the host checks 148 * nunits < 2^32 / G
__device__ __forceinline__ int cta_unit_lo(int first, int nunits, uint32_t g_magic) { return (int)__umulhi((uint32_t)first * (uint32_t)nunits, g_magic); }

// {value, tag} as one 64-bit store: the reader polls the tag (flag-in-data, no fence needed: the word is written atomically)
__device__ __forceinline__ void st_tagged(unsigned long long* base, size_t idx, float v, unsigned tag) {
  if (base) {
    const unsigned long long w = (unsigned long long)__float_as_uint(v) | ((unsigned long long)tag << 32);
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(base + idx), "l"(w) : "memory");
  }
}

// ---- a consumer warp's walk over its units of one op ---------------------------------------------------------------------
struct UnitCursor {
  int j;          // unit index inside this CTA's range of the op
  int gi, within; // its group inside the op, its position inside the group
  int slot;       // ring slot of that group
  uint32_t lap;   // how often the ring had wrapped when the group was issued (mbarrier phase)
};
struct UnitEnv {
  uint32_t ring, full0, empty0, group_bytes;
  int NG, gs, pitch, lo, ucount, lane, m;
  unsigned long long* tu;  // per-unit trace (CTA 0 of a TRACE build) or NULL
  unsigned out_tag;        // tag of this op's outputs (ops executed before it + 1)
};
struct UnitPre {
  uint32_t r0, r1;  // shared-memory addresses of the unit's rows
  float ep_res, ep_bias;
  int e_m, e_out;
  bool e_live;
};
__device__ __forceinline__ void unit_advance(UnitCursor& c, const UnitEnv& e, const ProgOp& P) {  // to unit j + kConsumers
  c.j += kConsumers;
  int dg = P.dgi;
  c.within += P.dwi;
  if (c.within >= e.gs) {
    c.within -= e.gs;
    ++dg;
  }
  c.gi += dg;
  c.slot += dg;
  while (c.slot >= e.NG) {
    c.slot -= e.NG;
    ++c.lap;
  }
}
// epilogue operands are fetched before the wait (their L2 latency hides behind the dot products); then the unit's bytes
template <int M>
__device__ __forceinline__ UnitPre unit_prologue(const ProgOp& P, const UnitEnv& e, const UnitCursor& c, const UnitOut& uo, bool gate_up,
                                                 const float* residual) {
  UnitPre p;
  p.ep_res = p.ep_bias = 0.f;
  p.e_m = gate_up ? e.lane : (e.lane >> 1);
  const int e_r = gate_up ? 0 : (e.lane & 1);
  p.e_live = p.e_m < e.m && p.e_m < M && (e_r == 0 || uo.valid1);
  p.e_out = uo.out0 + e_r;
  if (p.e_live && !gate_up) {
    const size_t o = (size_t)p.e_m * P.ldo + p.e_out;
    const float* bias = P.bias;
    if (bias) p.ep_bias = P.bias_bcast ? ldcg1f(bias + p.e_out) : ldcg1f(bias + o);
    if (residual) p.ep_res = ldcg1f(residual + o);
  }
  const int ui = e.ucount + c.j;
  const bool trace = e.tu && ui < kUnitTrace && e.lane == 0;
  if (trace) e.tu[(size_t)ui * 8 + 3] = clk64();
  mbar_wait(e.full0 + 8u * c.slot, c.lap & 1u);
  if (trace) e.tu[(size_t)ui * 8 + 4] = clk64();
  const uint32_t gbase = e.ring + (uint32_t)c.slot * e.group_bytes;
  const int unit_rows = P.unit_rows;
  p.r0 = gbase + (uint32_t)(gate_up ? c.within : unit_rows * c.within) * (uint32_t)e.pitch;
  p.r1 = gate_up ? gbase + (uint32_t)(e.gs + c.within) * (uint32_t)e.pitch : (uo.valid1 ? p.r0 + e.pitch : p.r0);
  return p;
}
template <int M>
__device__ __forceinline__ void unit_epilogue(const ProgOp& P, const UnitEnv& e, const UnitCursor& c, const UnitOut& uo, const UnitPre& p,
                                              bool gate_up, const float* residual, float (&acc)[2][M]) {
  __syncwarp();
  const int ui = e.ucount + c.j;
  if (e.tu && ui < kUnitTrace && e.lane == 0) e.tu[(size_t)ui * 8 + 5] = clk64();
  if (e.lane == 0) mbar_arrive(e.empty0 + 8u * c.slot);  // one of the group's units is done with the slot
  if (M == 1) {
    warp_sum2(acc[0][0], acc[1][0], e.lane);
  } else {
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int m = 0; m < M; ++m) acc[r][m] = warp_sum(acc[r][m]);
  }
  float* dst = P.dst;
  const int ldo = P.ldo, eltop = P.eltop;
  if (gate_up) {
    float g = 0.f, up = 0.f;
#pragma unroll
    for (int m = 0; m < M; ++m)
      if (e.lane == m) {
        g = acc[0][m];
        up = acc[1][m];
      }
    if (p.e_live) {
      const float sg = eltop == NS_ELT_GELU ? ns_gelu(g) : ns_silu(g);  // kernel_ref.h:1569-1576
      float* aux = P.aux;
      if (aux) aux[(size_t)p.e_m * ldo + uo.out0] = sg;
      dst[(size_t)p.e_m * ldo + uo.out0] = sg * up;
      st_tagged(P.dst_tag, (size_t)p.e_m * ldo + uo.out0, sg * up, e.out_tag);
    }
  } else {
    float v = 0.f;
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int m = 0; m < M; ++m)
        if (e.lane == 2 * m + r) v = acc[r][m];
    if (p.e_live) {
      if (P.bias) v += p.ep_bias;
      if (eltop == NS_ELT_GELU) v = ns_gelu(v);
      if (residual) v += p.ep_res;
      dst[(size_t)p.e_m * ldo + p.e_out] = v;
      st_tagged(P.dst_tag, (size_t)p.e_m * ldo + p.e_out, v, e.out_tag);
    }
  }
}

// TRACE: clock stamps per (op, CTA) and per unit of CTA 0 (debug builds of the common configuration only: the stamps cost
// registers in the hot loop, and a spilled register costs an L2 round trip here)
template <int COMP, int M, bool ASYM, int STYPE, bool TRACE>
__global__ void __launch_bounds__(kThreads, 1)
    program_kernel(const ProgOp* __restrict__ ops, int nops, const ProgCfg R, unsigned* __restrict__ counters,
                   unsigned* __restrict__ epoch_ptr, unsigned long long* __restrict__ tl_, unsigned long long* __restrict__ tu_) {
  unsigned long long* const tl = TRACE ? tl_ : nullptr;
  unsigned long long* const tu_dbg = TRACE ? tu_ : nullptr;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ ProgOp op_s[2];  // the consumers' current and next op (the next one is fetched during the current one)
  __shared__ ProgOp op_p;     // the producer's current op
  __shared__ float red_s[kConsumers];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t ring = smem_base + R.ring_off;
  const uint32_t full0 = smem_base + R.bar_off;
  const int NG = R.ngroups;
  const uint32_t empty0 = full0 + 8u * (uint32_t)NG;
  const int first = blockIdx.x, G = (int)gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NG; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, R.gs_max);  // one arrival per unit of the group (the producer stands in for missing ones)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // launches completed so far: every CTA arrives once per op per launch on counters[op], so after `epoch` launches of `iters`
  // iterations each counter stands at (sum of earlier iterations) * G; epoch_ptr[0] holds that sum of iterations.
  const unsigned base_iters = ld_acquire(epoch_ptr);
  const unsigned tag_base = ld_acquire(epoch_ptr + 1);  // ops executed by earlier launches + 1 (tags are never 0)

  if (warp == kConsumers) {
    // ===================== producer: streams the weights of ALL ops, never waits for activations =====================
    // A CTA owns a CONTIGUOUS range of an op's units, so a group of gs units is one contiguous byte range per weight: one
    // bulk copy per group.  Lane l issues group gb + l of each step.  L1 is all but carved away by the shared memory, so the
    // op descriptor is copied to shared memory once per op instead of being re-read through L2.
    int gslot0 = 0;        // ring slot of the op's first group
    uint32_t glap0 = 0;    // and how many times the ring has wrapped before it
    int ucount = 0;        // CTA-wide index of the op's first unit (trace only)
    constexpr int kWords = (int)(sizeof(ProgOp) / 4);
    for (int it = 0; it < R.iters; ++it) {
      for (int oi = 0; oi < nops; ++oi) {
        __syncwarp();
        for (int w = lane; w < kWords; w += 32)
          reinterpret_cast<uint32_t*>(&op_p)[w] = reinterpret_cast<const uint32_t*>(ops + oi)[w];
        __syncwarp();
        const int nunits = op_p.nunits, gs = op_p.gs, unit_rows = op_p.unit_rows, mode = op_p.mode;
        const uint32_t pitch = (uint32_t)op_p.pitch;
        const int lo = cta_unit_lo(first, nunits, R.g_magic), hi = cta_unit_lo(first + 1, nunits, R.g_magic);
        const int my_units = hi - lo, ng = op_p.gs_magic ? (int)__umulhi((uint32_t)(my_units + gs - 1), op_p.gs_magic) : my_units;
        const size_t tli = ((size_t)(it * nops + oi) * G + first) * kTl;
        if (tl && lane == 0) tl[tli + 5] = clk64();
        for (int gb = 0; gb < ng; gb += R.batch) {
          const int gi = gb + lane;
          bool pending = lane < R.batch && gi < ng;
          int slot = gslot0 + gi;
          uint32_t lap = glap0;
          while (slot >= NG) {
            slot -= NG;
            ++lap;
          }
          const int u0 = lo + gi * gs, n = min(gs, hi - u0);
          const bool trace = pending && tu_dbg && first == 0 && ucount + gi * gs < kUnitTrace;
          if (trace) tu_dbg[(size_t)(ucount + gi * gs) * 8 + 0] = clk64();
          // every lane issues its group as soon as ITS slot is free (a blocking wait would hold the whole warp back until the
          // slowest slot of the step frees: the refill of the other slots then comes late and the consumers starve)
          while (__any_sync(0xffffffffu, pending)) {
            const bool ready = pending && (lap == 0 || mbar_test(empty0 + 8u * slot, (lap - 1) & 1u));
            if (ready) {
              if (trace) tu_dbg[(size_t)(ucount + gi * gs) * 8 + 1] = clk64();
              const uint32_t dst = ring + (uint32_t)slot * (uint32_t)R.group_bytes;
              const uint32_t bar = full0 + 8u * slot;
              if (mode == NS_GEMV_GATE_UP_SILU) {  // [gs gate rows][gs up rows]
                mbar_expect_tx(bar, 2u * (uint32_t)n * pitch);
                bulk_g2s(dst, op_p.rows[0] + (size_t)u0 * pitch, (uint32_t)n * pitch, bar);
                bulk_g2s(dst + (uint32_t)gs * pitch, op_p.rows[1] + (size_t)u0 * pitch, (uint32_t)n * pitch, bar);
              } else {
                // rows [ra, rb) of the concatenated weights, split where they cross from one weight into the next
                int ra = unit_rows * u0, rb = unit_rows * (u0 + n);
                const int ntot = op_p.n[0] + (op_p.nw > 1 ? op_p.n[1] : 0) + (op_p.nw > 2 ? op_p.n[2] : 0);
                if (rb > ntot) rb = ntot;
                mbar_expect_tx(bar, (uint32_t)(rb - ra) * pitch);
                int seg0 = 0;
                uint32_t d = dst;
#pragma unroll
                for (int sgi = 0; sgi < 3; ++sgi) {
                  if (sgi < op_p.nw) {
                    const int seg1 = seg0 + op_p.n[sgi];
                    const int a2 = max(ra, seg0), b2 = min(rb, seg1);
                    if (a2 < b2) {
                      bulk_g2s(d, op_p.rows[sgi] + (size_t)(a2 - seg0) * pitch, (uint32_t)(b2 - a2) * pitch, bar);
                      d += (uint32_t)(b2 - a2) * pitch;
                    }
                    seg0 = seg1;
                  }
                }
              }
              for (int e = n; e < R.gs_max; ++e) mbar_arrive(empty0 + 8u * slot);  // arrivals of the units this group does not have
              if (trace) tu_dbg[(size_t)(ucount + gi * gs) * 8 + 2] = clk64();
              pending = false;
            }
          }
          __syncwarp();
        }
        gslot0 += ng;
        while (gslot0 >= NG) {
          gslot0 -= NG;
          ++glap0;
        }
        ucount += my_units;
        if (tl && lane == 0) tl[tli + 6] = clk64();
      }
    }
    return;
  }

  // ===================== consumers =====================
  constexpr int AMODE = (COMP == NS_COMP_INT8) ? A_U8 : A_S8;
  const int total_ops = nops * R.iters;
  int gslot0 = 0;         // ring slot of this op's first group
  uint32_t glap0 = 0;     // and how many times the ring has wrapped before it
  int ucount = 0;         // CTA-wide index of this op's first unit
  copy_op(&op_s[0], ops, threadIdx.x);
  nsq::bar_sync<1, kConsumerThreads>();
  for (int seq = 0; seq < total_ops; ++seq) {
    const int it = seq / nops, oi = seq - it * nops;
    const size_t tli = ((size_t)seq * G + first) * kTl;
    // ---- op boundary: wait until every CTA has published the previous op (the descriptor is already in shared memory) ----
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
    nsq::bar_sync<1, kConsumerThreads>();
    const ProgOp& P = op_s[seq & 1];
    if (seq + 1 < total_ops) copy_op(&op_s[(seq + 1) & 1], ops + (seq + 1) % nops, threadIdx.x);  // read again only after >= 2 barriers
    {
      const float* in = P.in;
      if (P.in_index) in += (long long)ldcg1i(P.in_index) * P.in_stride;
      const nsq::NormQuantIn qi{in, P.in_tagged ? tag_base + (unsigned)seq : 0u, P.norm_w, P.norm_eps, P.lda, P.k, P.kpad, COMP == NS_COMP_Q8_0 ? 32 : P.group,
                                P.act_row, P.meta_off, P.meta_stride};
      nsq::norm_quantise_to_smem<COMP, kConsumerThreads, 1>(qi, R.m, smem_base, red_s, threadIdx.x);
    }
    nsq::bar_sync<1, kConsumerThreads>();
    if (tl && threadIdx.x == 0) tl[tli + 2] = clk64();

    // Only what the dot-product loop needs lives in registers; the epilogue re-reads the op's fields from shared memory (a few
    // LDS per unit), and the two activation paths are two separate loops: with 44 registers pinned by register-resident
    // activations, one merged loop spilled, and a spill costs an L2 round trip here (the L1 is carved away).
    const int pitch = P.pitch, gs = P.gs;
    const bool gate_up = P.mode == NS_GEMV_GATE_UP_SILU;
    const int lo = cta_unit_lo(first, P.nunits, R.g_magic), hi = cta_unit_lo(first + 1, P.nunits, R.g_magic);
    const int my_units = hi - lo, ng = P.gs_magic ? (int)__umulhi((uint32_t)(my_units + gs - 1), P.gs_magic) : my_units;
    const float* residual = P.residual;
    if (residual && P.res_index) residual += (long long)ldcg1i(P.res_index) * P.res_stride;
    const bool use_reg = (M == 1) && P.use_reg;

    // this warp's units of the op: unit j (0 .. my_units) with (ucount + j) % kConsumers == warp
    UnitCursor cur;
    cur.j = warp - (ucount % kConsumers);
    if (cur.j < 0) cur.j += kConsumers;
    cur.gi = P.gs_magic ? (int)__umulhi((uint32_t)cur.j, P.gs_magic) : cur.j;
    cur.within = cur.j - cur.gi * gs;
    cur.slot = gslot0 + cur.gi;
    cur.lap = glap0;
    while (cur.slot >= NG) {
      cur.slot -= NG;
      ++cur.lap;
    }
    if (tl && threadIdx.x == 0) tl[tli + 8] = clk64();
    const UnitEnv env{ring, full0, empty0, (uint32_t)R.group_bytes, NG, gs, pitch, lo, ucount, lane, R.m,
                      TRACE && first == 0 ? tu_dbg : nullptr, tag_base + (unsigned)seq + 1u};
    const unsigned out_tag = env.out_tag;

    if (use_reg) {
      // register-resident activations (M == 1, nchunks a multiple of 32): chunk c = lane + 32 i, i < kc
      const uint32_t meta_s = smem_base + P.meta_off;
      const int kc = P.kpad >> 10;
      uint4 A0[kKcReg], A1[kKcReg];
      float asc[kKcReg];
      int asa[kKcReg], aza[kKcReg];
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
        } else {
          A0[i] = A1[i] = make_uint4(0, 0, 0, 0);
          asc[i] = 0.f;
          asa[i] = aza[i] = 0;
        }
      }
      // byte offset of chunk i's scale / zero point inside a weight row: base + i * step (cpg divides 32 on this path)
      const int cpg = P.cpg;
      const int g0 = (cpg == 1) ? lane : (int)__umulhi((uint32_t)lane, P.cpg_magic);
      const uint32_t ssz = STYPE == NS_S_F32 ? 4u : 2u;
      const uint32_t soff0 = (uint32_t)P.sc_off + (uint32_t)g0 * ssz, sstep = (uint32_t)P.cstep * ssz;
      const uint32_t zoff0 = (uint32_t)P.zp_off + (uint32_t)g0, zstep = (uint32_t)P.cstep;
      if (tl && threadIdx.x == 0) tl[tli + 9] = clk64();
      // TWO units per trip: the part of a unit after its dot products (reduction shuffles, epilogue, cursor, the next try_wait)
      // is ~250 dependent instructions that all 15 warps run at the same time -- the SM idles through it (measured: 1000
      // cycles of dots, 1100 cycles of the rest per unit).  Sharing that part between two units halves it per unit.
      while (cur.j < my_units) {
        const UnitCursor c0 = cur;
        unit_advance(cur, env, P);
        const bool has1 = cur.j < my_units;
        const UnitCursor c1 = cur;
        if (has1) unit_advance(cur, env, P);
        const UnitOut uo0 = unit_out(P, lo + c0.j);
        const UnitOut uo1 = has1 ? unit_out(P, lo + c1.j) : uo0;
        // the lane that stores (unit k, row r): 16 r + 8 k; it fetches that output's bias / residual ahead of the dots
        const int sk = (lane >> 3) & 1, sr = lane >> 4;
        const UnitOut& suo = sk ? uo1 : uo0;
        const bool s_live = (lane & 7) == 0 && (sk == 0 || has1) && (gate_up ? sr == 0 : (sr == 0 || suo.valid1));
        const int s_out = suo.out0 + (gate_up ? 0 : sr);
        float ep_res = 0.f, ep_bias = 0.f;
        if (s_live && !gate_up) {
          const float* bias = P.bias;
          if (bias) ep_bias = ldcg1f(bias + s_out);  // M == 1: broadcast and per-row bias coincide
          if (residual) ep_res = ldcg1f(residual + s_out);
        }
        float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};  // [unit][row]
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          if (k == 1 && !has1) break;
          const UnitCursor& c = k ? c1 : c0;
          const UnitOut& uo = k ? uo1 : uo0;
          const int ui = env.ucount + c.j;
          const bool trace = env.tu && ui < kUnitTrace && lane == 0;
          if (trace) env.tu[(size_t)ui * 8 + 3] = clk64();
          mbar_wait(full0 + 8u * c.slot, c.lap & 1u);
          if (trace) env.tu[(size_t)ui * 8 + 4] = clk64();
          const uint32_t gbase = ring + (uint32_t)c.slot * (uint32_t)R.group_bytes;
          const uint32_t r0 = gbase + (uint32_t)(gate_up ? c.within : 2 * c.within) * (uint32_t)pitch;
          const uint32_t r1 = gate_up ? gbase + (uint32_t)(gs + c.within) * (uint32_t)pitch : (uo.valid1 ? r0 + pitch : r0);
          const uint32_t rb[2] = {r0, r1};
          const uint32_t wb[2] = {r0 + 16u * lane, r1 + 16u * lane};
          switch (kc) {  // straight-line code per row length
            case 4: reg_unit<4, AMODE, ASYM, STYPE>(wb, rb, A0, A1, asc, asa, aza, soff0, sstep, zoff0, zstep, acc[k][0], acc[k][1]); break;
            case 3: reg_unit<3, AMODE, ASYM, STYPE>(wb, rb, A0, A1, asc, asa, aza, soff0, sstep, zoff0, zstep, acc[k][0], acc[k][1]); break;
            case 2: reg_unit<2, AMODE, ASYM, STYPE>(wb, rb, A0, A1, asc, asa, aza, soff0, sstep, zoff0, zstep, acc[k][0], acc[k][1]); break;
            default: reg_unit<1, AMODE, ASYM, STYPE>(wb, rb, A0, A1, asc, asa, aza, soff0, sstep, zoff0, zstep, acc[k][0], acc[k][1]); break;
          }
          __syncwarp();
          if (trace) env.tu[(size_t)ui * 8 + 5] = clk64();
          if (lane == 0) mbar_arrive(empty0 + 8u * c.slot);  // one of the group's units is done with the slot
        }
        // four sums with nine shuffles, same butterfly order (16, 8, 4, 2, 1) as warp_sum: after the xor-16 step lanes < 16
        // carry row 0 and lanes >= 16 row 1; after the xor-8 step lanes with bit 3 clear carry unit 0, the others unit 1
        float v0, v1;
        {
          const float a0 = acc[0][0] + __shfl_xor_sync(0xffffffffu, acc[0][0], 16);
          const float b0 = acc[0][1] + __shfl_xor_sync(0xffffffffu, acc[0][1], 16);
          const float a1 = acc[1][0] + __shfl_xor_sync(0xffffffffu, acc[1][0], 16);
          const float b1 = acc[1][1] + __shfl_xor_sync(0xffffffffu, acc[1][1], 16);
          v0 = lane < 16 ? a0 : b0;
          v1 = lane < 16 ? a1 : b1;
        }
        v0 += __shfl_xor_sync(0xffffffffu, v0, 8);
        v1 += __shfl_xor_sync(0xffffffffu, v1, 8);
        float v = (lane & 8) ? v1 : v0;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        // lane 16 r + 8 k holds row r of unit k
        if (gate_up) {
          const float up = __shfl_sync(0xffffffffu, v, (lane & 8) | 16);  // the up row of this lane's unit
          if (s_live) {
            const float sg = P.eltop == NS_ELT_GELU ? ns_gelu(v) : ns_silu(v);  // kernel_ref.h:1569-1576
            float* aux = P.aux;
            if (aux) aux[s_out] = sg;
            P.dst[s_out] = sg * up;
            st_tagged(P.dst_tag, s_out, sg * up, out_tag);
          }
        } else if (s_live) {
          if (P.bias) v += ep_bias;
          if (P.eltop == NS_ELT_GELU) v = ns_gelu(v);
          if (residual) v += ep_res;
          P.dst[s_out] = v;
          st_tagged(P.dst_tag, s_out, v, out_tag);
        }
      }
    } else {
      const SmemUnitArgs ua0{0, 0, smem_base, smem_base + (uint32_t)P.meta_off, P.sc_off, P.zp_off, P.cpg, P.cpg_magic, P.act_row,
                             P.meta_stride, P.kpad >> 5, lane};
      for (; cur.j < my_units; unit_advance(cur, env, P)) {
        const UnitOut uo = unit_out(P, lo + cur.j);
        const UnitPre pre = unit_prologue<M>(P, env, cur, uo, gate_up, residual);
        SmemUnitArgs ua = ua0;
        ua.r0 = pre.r0;
        ua.r1 = pre.r1;
        float acc[2][M];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int m = 0; m < M; ++m) acc[r][m] = 0.f;
        if (P.unit_rows == 2) smem_unit<2, M, AMODE, ASYM, STYPE>(ua, acc);
        else smem_unit<1, M, AMODE, ASYM, STYPE>(ua, acc);
        unit_epilogue<M>(P, env, cur, uo, pre, gate_up, residual, acc);
      }
    }
    gslot0 += ng;
    while (gslot0 >= NG) {
      gslot0 -= NG;
      ++glap0;
    }
    ucount += my_units;
    // ---- op done in this CTA: publish (release) ----
    if (P.publish) {
      nsq::bar_sync<1, kConsumerThreads>();
      if (threadIdx.x == 0) {
        if (tl) tl[tli + 3] = clk64();
        red_release_add(counters + oi, 1u);
        if (tl) tl[tli + 4] = clk64();
      }
    } else {
      if (tl && threadIdx.x == 0) tl[tli + 3] = tl[tli + 4] = clk64();
      nsq::bar_sync<1, kConsumerThreads>();  // every warp is done with the activation image and the op descriptor
    }
  }
  // last op finished everywhere -> advance the epoch exactly once (block 0), so the next launch sees fresh targets
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const unsigned want = (base_iters + (unsigned)R.iters) * (unsigned)G;
    while ((int)(ld_acquire(counters + (nops - 1)) - want) < 0) {
    }
    red_release_add(epoch_ptr, (unsigned)R.iters);
    red_release_add(epoch_ptr + 1, (unsigned)total_ops);
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
  op.use_reg = (p->m == 1 && (w0->kpad >> 5) <= 32 * kKcReg && (w0->kpad >> 5) % 32 == 0 && 32 % op.cpg == 0 && !no_reg) ? 1 : 0;
  p->ops.push_back(op);
  return NS_OK;
}

extern "C" int ns_program_add_matmul(ns_program* p, const ns_weight* const* weights, int nw, int mode, const float* in, int lda,
                                     float* dst, int ldo, const float* bias, int bias_bcast, const float* residual,
                                     float* aux, int barrier_before) {
  return ns_program_add_matmul_ex(p, weights, nw, mode, in, lda, dst, ldo, bias, bias_bcast, residual, aux, barrier_before,
                                  nullptr, 0.f, nullptr, 0, nullptr, 0, NS_ELT_DEFAULT);
}

// Flag-in-data hand-over for the op added last: in_tagged != 0 -> its input `in` is a [m][lda] array of 8-byte {value, tag}
// words written by the PREVIOUS op of the program (that op's dst_tag), polled word by word instead of waiting on a grid barrier
// (barrier_before is cleared); dst_tag != NULL -> the op also writes its outputs as such words to dst_tag [m][ldo].
extern "C" int ns_program_tag_last(ns_program* p, int in_tagged, void* dst_tag) {
  if (!p || p->finalized || p->ops.empty()) return NS_E_INVALID;
  ProgOp& o = p->ops.back();
  if (in_tagged && (p->ops.size() < 2 || o.in_index || o.kpad > 3 * kConsumerThreads * 8)) {
    ns_set_error("ns_program_tag_last: a tagged input needs a producing op before it, no input indirection and k <= %d",
                 3 * kConsumerThreads * 8);
    return NS_E_INVALID;
  }
  o.in_tagged = in_tagged ? 1 : 0;
  if (in_tagged) o.barrier_before = 0;
  o.dst_tag = (unsigned long long*)dst_tag;
  return NS_OK;
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
  static const int env_kb = getenv("NS_PROG_SMEM_KB") ? atoi(getenv("NS_PROG_SMEM_KB")) : 0;  // tuning aids
  static const int env_gb = getenv("NS_PROG_GROUP_KB") ? atoi(getenv("NS_PROG_GROUP_KB")) : 0;
  static const int env_batch = getenv("NS_PROG_BATCH") ? atoi(getenv("NS_PROG_BATCH")) : 0;
  const size_t budget = (size_t)(env_kb > 0 ? env_kb : 222) * 1024;  // the static shared memory (descriptors, reductions) rides on top
  // Ring slot = one GROUP of units fetched by one bulk copy.  Sized for two pairs of the longest rows when at least six such
  // slots fit, else one pair, else one row; shorter rows pack floor(slot / unit) units (<= kMaxGs) into a slot.
  int max_pitch = 0;
  for (const ProgOp& o : p->ops) max_pitch = std::max(max_pitch, o.pitch);
  int group = 0, ngroups = 0;
  const int cand[3] = {4 * max_pitch, 2 * max_pitch, max_pitch};
  for (int ci = 0; ci < 3; ++ci) {
    group = env_gb > 0 ? env_gb * 1024 : cand[ci];
    group = (int)ns_round_up((size_t)group, 128);
    ngroups = budget > act_region + 64 ? (int)((budget - act_region - 64) / ((size_t)group + 16)) : 0;
    if (ngroups >= 6 || env_gb > 0) break;
  }
  if (ngroups > 32) ngroups = 32;
  if (ngroups < 2 || group < max_pitch) {
    ns_set_error("ns_program: rows too long for the shared-memory ring (%d B per row)", max_pitch);
    return NS_E_UNSUPPORTED;
  }
  int gs_max = 1;
  for (ProgOp& o : p->ops) {
    const bool gate_up = o.mode == NS_GEMV_GATE_UP_SILU;
    if (gate_up && 2 * o.pitch > group) {
      ns_set_error("ns_program: a gate/up row pair (%d B) does not fit a ring slot", 2 * o.pitch);
      return NS_E_UNSUPPORTED;
    }
    o.unit_rows = (gate_up || 2 * o.pitch <= group) ? 2 : 1;
    long long rows = 0;
    for (int i = 0; i < o.nw; ++i) rows += o.n[i];
    o.nunits = o.unit_rows == 2 ? o.npairs : (int)rows;
    o.gs = std::max(1, std::min(kMaxGs, group / (o.unit_rows * o.pitch)));
    gs_max = std::max(gs_max, o.gs);
    o.dgi = kConsumers / o.gs;
    o.dwi = kConsumers % o.gs;
    o.cstep = 32 / std::max(1, std::min(32, o.cpg));
    o.gs_magic = o.gs > 1 ? (uint32_t)((0x100000000ull + (uint64_t)o.gs - 1) / (uint64_t)o.gs) : 0u;  // 0: gs == 1, no division
  }
  p->cfg.ring_off = (int)act_region;
  p->cfg.group_bytes = group;
  p->cfg.ngroups = ngroups;
  p->cfg.gs_max = gs_max;
  p->cfg.bar_off = (int)(act_region + (size_t)ngroups * group);
  p->cfg.m = p->m;
  p->cfg.iters = 1;
  p->cfg.batch = std::max(1, std::min(ngroups, env_batch > 0 ? env_batch : 8));
  p->grid = ns_num_sms();
  p->cfg.g_magic = (uint32_t)((0x100000000ull + (uint64_t)p->grid - 1) / (uint64_t)p->grid);
  for (const ProgOp& o : p->ops)
    if ((uint64_t)(p->grid + 1) * (uint64_t)o.nunits >= 0x100000000ull / (uint64_t)p->grid) {
      ns_set_error("ns_program: too many rows (%d units) for the 32-bit unit split", o.nunits);
      return NS_E_UNSUPPORTED;
    }
  p->smem = act_region + (size_t)ngroups * group + (size_t)ngroups * 16;
  p->grid = ns_num_sms();
  const size_t nops = p->ops.size();
  for (size_t i = 0; i < nops; ++i)  // an arrival counter is only bumped where somebody waits on it
    p->ops[i].publish = (i + 1 == nops || p->ops[i + 1].barrier_before) ? 1 : 0;
  NS_CUDA_TRY(cudaMalloc((void**)&p->d_ops, nops * sizeof(ProgOp)));
  NS_CUDA_TRY(cudaMalloc((void**)&p->d_counters, (nops + 2) * sizeof(unsigned)));
  NS_CUDA_TRY(cudaMemcpyAsync(p->d_ops, p->ops.data(), nops * sizeof(ProgOp), cudaMemcpyHostToDevice, st));
  NS_CUDA_TRY(cudaMemsetAsync(p->d_counters, 0, (nops + 2) * sizeof(unsigned), st));
  {
    const unsigned one = 1;  // tags start at 1: zero-initialised tagged buffers never match
    NS_CUDA_TRY(cudaMemcpyAsync(p->d_counters + nops + 1, &one, sizeof(one), cudaMemcpyHostToDevice, st));
  }
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

template <int COMP, int M, bool ASYM, int STYPE, bool TRACE>
static int run_one(ns_program* p, int iters, cudaStream_t st) {
  auto kern = program_kernel<COMP, M, ASYM, STYPE, TRACE>;
  static size_t attr_smem = 0;
  if (p->smem > attr_smem) {
    NS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem));
    attr_smem = p->smem;
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
    case 1:
      if (COMP == NS_COMP_Q8_0 && !ASYM && p->d_tl && iters == 1) return run_one<COMP, 1, ASYM, STYPE, (COMP == NS_COMP_Q8_0 && !ASYM)>(p, iters, st);
      return run_one<COMP, 1, ASYM, STYPE, false>(p, iters, st);
    case 2: return run_one<COMP, 2, ASYM, STYPE, false>(p, iters, st);
    default: return run_one<COMP, 4, ASYM, STYPE, false>(p, iters, st);
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
