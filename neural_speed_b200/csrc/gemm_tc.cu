// gemm_tc.cu -- prefill / batched weight-only matmul on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), M > 4 rows.
//
// Replaces the reference's blocked GEMM for int4 weights: LauncherBase::run_block / LauncherIntKBlock::run_block
// (bestla/bestla/bestla_wrapper.h:501,768) with WeightKBlockNInteger::getWeight + decompress_kblock_s4_fp
// (bestla_prologue_b.h:642-733, kernel_ref.h:1113) feeding the JIT AMX/AVX512 micro-kernels (bestla_gemm.h), and the
// ggml Q4_0 mul_mat for prompt batches (core/ne_layers.c:7085).  Numerics: weights are dequantised to bf16
// ((u - 8 - zp) exactly in bf16, times the bf16-rounded scale), activations rounded to bf16, fp32 accumulation in TMEM --
// the reference's own CompBf16 numerics (BF16 tolerance 2e-2 at K=4096, bestla_ut.h:80-94); logits stay within the
// north-star's 1e-2 relative bar of the CPU path (tests/test_gpu_gemm_tc.py).
//
// Orientation ("swap-AB"): the WEIGHT tile is the UMMA M operand (128 output channels = 128 TMEM lanes) and the token tile
// is the UMMA N operand (T = 32..256 columns), so small batches do not waste the 128-row MMA and the epilogue's stores
// are coalesced along n.  Per CTA and per 64-wide k block:
//   warp 0   TMA producer: packed nibbles [128 rows][32 B] (cp.async.bulk.tensor 2-D over the NSB rows, row pitch = pitch)
//            and bf16 activations [T][64] (128B-swizzled) into two mbarrier rings
//   warps 8-11  dequant: 1 thread = 1 weight row: 2 x LDS.128 of nibbles -> 64 bf16 via (w >> 4j) & 0x000F000F | 0x4300
//            (bf16 128+u), HSUB2 (exact), HMUL2 by the group scale -> 8 x STS.128 into the 128B-swizzled K-major tile,
//            fence.proxy.async, arrive
//   warp 1   MMA issuer: 4 x tcgen05.mma.cta_group::1.kind::f16 (M=128, N=T, K=16) per k block, accumulator in TMEM,
//            tcgen05.commit releases the smem slots
//   warps 4-7   epilogue: tcgen05.ld 32x32b.x32 -> (+bias, +residual) -> fp32 global stores, coalesced over n
//   warp 2   TMEM alloc / dealloc
// Roofline: tensor (bf16, fp32 accumulate): flops = 2*M*N*K per launch.
#include <cuda.h>

#include "nsb.cuh"

namespace {

constexpr int BLOCK_N = 128;  // weight rows per CTA (UMMA M)
constexpr int BLOCK_K = 64;   // bf16 elements per k block (one 128-byte swizzle atom)
constexpr int UMMA_K = 16;
constexpr int PACKED_TILE = BLOCK_N * (BLOCK_K / 2);  // 4096: nibbles of one 128-row sub-tile and k block (x2 for int8 weights)
constexpr int DEQ_TILE = BLOCK_N * BLOCK_K * 2;       // 16384: the same as bf16

struct GemmParams {
  const uint8_t* rows;
  int wfmt, f4kind;
  int pitch, sc_off, zp_off, stype, asym, group, ngroups;
  int n, k, kpad, m;
  float* dst;
  int ldo;
  const float* bias;
  int bias_bcast;
  const float* residual;
  int kb_per_split;  // k blocks per grid.z slice (>= total: no split)
  int dbg;  // NS_TC_DEBUG: 1 = dequant warps skip the conversion, 2 = no MMAs issued, 4 = epilogue skipped (timing experiments)
};

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
// TMA 2-D tile load (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, fp32 accumulate (SASS: UTCHMMA)
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, 128B-swizzled tile: 8-row groups 1024 B apart (SBO), LBO unused (=1), version 1 (Blackwell), layout SWIZZLE_128B
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// c_format f32 (1<<4), a/b format bf16 (1<<7, 1<<10), both K-major, N>>3 at bit 17, M>>4 at bit 24
__device__ __forceinline__ uint32_t make_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

#define TMEM_LD_32X32B_X32(taddr, r)                                                                                        \
  asm volatile(                                                                                                             \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                             \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                              \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                              \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),         \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), \
        "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),             \
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                                        \
      : "r"(taddr)                                                                                                          \
      : "memory")

// NB = number of 128-row weight sub-tiles a CTA owns (each with its own TMEM accumulator).  The bf16 activation tile is the
// expensive operand to re-read (2 B/element against 0.5 B for the weights): at T = 256, NB = 1 a CTA pulls 36 KB per 4.2 MFLOP
// and the GEMM is bound by L2 -> SM bandwidth (measured 36 % of the tensor peak); NB = 2 halves the activation traffic per flop.
// W8: int8 weights (64 packed bytes per row and k block instead of 32).
template <int T, int NB, bool W8 = false>
struct Smem {
  static constexpr int SP = NB == 2 ? (W8 ? 3 : 4) : 6;      // packed-weight stages
  static constexpr int SD = NB == 2 ? 2 : 3;                 // dequantised-weight stages
  static constexpr int SA = NB == 2 ? (T >= 256 ? 3 : 4) : ((T >= 256) ? 4 : 6);  // activation stages
  static constexpr int ROW_BYTES = W8 ? BLOCK_K : BLOCK_K / 2;  // packed bytes per weight row and k block
  static constexpr int PACKED_STAGE = NB * BLOCK_N * ROW_BYTES;
  static constexpr int DEQ_STAGE = NB * DEQ_TILE;
  static constexpr int kThreads = 256 + 128 * NB;
  static constexpr int ACT_STAGE = T * BLOCK_K * 2;
  static constexpr int off_deq = 0;
  static constexpr int off_act = off_deq + SD * DEQ_STAGE;
  static constexpr int off_packed = off_act + SA * ACT_STAGE;
  static constexpr int off_bar = off_packed + SP * PACKED_STAGE;
  static constexpr int num_bars = 2 * SP + 2 * SA + 2 * SD + 1;
  static constexpr int off_tmem_ptr = off_bar + num_bars * 8;
  static constexpr int total = off_tmem_ptr + 16 + 1024;  // + slack for manual 1024-B alignment
};

template <int T, int NB, bool W8>
__global__ void __launch_bounds__(256 + 128 * NB, 1)
    gemm_w4_tc_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_a, const GemmParams P) {
  using L = Smem<T, NB, W8>;
  constexpr int ROW_BYTES = L::ROW_BYTES;
  constexpr int SA = L::SA, SP = L::SP, SD = L::SD, PACKED_STAGE = L::PACKED_STAGE, DEQ_STAGE = L::DEQ_STAGE;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  unsigned char* deq = smem + L::off_deq;
  unsigned char* act = smem + L::off_act;
  unsigned char* packed = smem + L::off_packed;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::off_bar);
  uint64_t* p_full = bars;
  uint64_t* p_empty = p_full + SP;
  uint64_t* a_full = p_empty + SP;
  uint64_t* a_empty = a_full + SA;
  uint64_t* d_full = a_empty + SA;
  uint64_t* d_empty = d_full + SD;
  uint64_t* tmem_full = d_empty + SD;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::off_tmem_ptr);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * (BLOCK_N * NB);
  const int t0 = blockIdx.y * T;
  // split-K (small M: too few output tiles to fill 148 SMs): this CTA owns k blocks [kb0, kb0 + num_kb) and adds its partial
  // tile to dst with fp32 atomics (dst zeroed by the launcher; bias / residual applied by split 0)
  const int total_kb = (P.kpad + BLOCK_K - 1) / BLOCK_K;
  const int kb0 = blockIdx.z * P.kb_per_split;
  const int num_kb = (total_kb - kb0 < P.kb_per_split) ? total_kb - kb0 : P.kb_per_split;

  __shared__ float nf4_lut[16];  // constant memory serialises divergent indices; shared memory broadcasts per bank
  if (threadIdx.x >= 32 && threadIdx.x < 48) nf4_lut[threadIdx.x - 32] = NS_F4_LUT[P.f4kind][threadIdx.x - 32];
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    for (int i = 0; i < SP; ++i) {
      mbar_init(&p_full[i], 1);
      mbar_init(&p_empty[i], 4 * NB);
    }
    for (int i = 0; i < SA; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < SD; ++i) {
      mbar_init(&d_full[i], 4 * NB);
      mbar_init(&d_empty[i], 1);
    }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"((uint32_t)(T * NB))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      // packed weights never depend on the previous kernel: prefetch the first ring before griddepcontrol.wait
      const int pre = num_kb < SP ? num_kb : SP;
      for (int kb = 0; kb < pre; ++kb) {
        mbar_expect_tx(&p_full[kb], PACKED_STAGE);
        tma_load_2d(packed + kb * PACKED_STAGE, &tmap_w, (kb0 + kb) * ROW_BYTES, n0, &p_full[kb]);
      }
      pdl_wait();  // activations were written by the preceding kernel
      for (int kb = 0; kb < num_kb; ++kb) {
        const int sa = kb % SA;
        if (kb >= SA) mbar_wait(&a_empty[sa], ((kb / SA) - 1) & 1);
        mbar_expect_tx(&a_full[sa], L::ACT_STAGE);
        tma_load_2d(act + sa * L::ACT_STAGE, &tmap_a, (kb0 + kb) * BLOCK_K, t0, &a_full[sa]);
        const int kp = kb + SP;  // keep the packed ring SP blocks ahead
        if (kp < num_kb) {
          const int sp = kp % SP;
          mbar_wait(&p_empty[sp], ((kp / SP) - 1) & 1);
          mbar_expect_tx(&p_full[sp], PACKED_STAGE);
          tma_load_2d(packed + sp * PACKED_STAGE, &tmap_w, (kb0 + kp) * ROW_BYTES, n0, &p_full[sp]);
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    const uint32_t idesc = make_idesc_bf16(BLOCK_N, T);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int sd = kb % SD, sa = kb % SA;
      mbar_wait(&d_full[sd], (kb / SD) & 1);
      mbar_wait(&a_full[sa], (kb / SA) & 1);
      tcgen05_fence_after();
      if (lane == 0) {
        const uint32_t a_addr = smem_u32(deq + sd * DEQ_STAGE);
        const uint32_t b_addr = smem_u32(act + sa * L::ACT_STAGE);
        if (!(P.dbg & 2))
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
#pragma unroll
          for (int k4 = 0; k4 < BLOCK_K / UMMA_K; ++k4) {
            umma_bf16(tmem_base + (uint32_t)(nb * T), make_desc_sw128(a_addr + nb * DEQ_TILE + k4 * UMMA_K * 2),
                      make_desc_sw128(b_addr + k4 * UMMA_K * 2), idesc, (uint32_t)((kb | k4) != 0));
          }
        umma_commit(&d_empty[sd]);  // fires when the MMAs above have finished reading shared memory
        umma_commit(&a_empty[sa]);
        if (kb == num_kb - 1) umma_commit(tmem_full);
      }
      __syncwarp();
    }
  } else if (warp >= 8) {
    // ============================ dequant: one thread per weight row (128 * NB rows) ============================
    const int r = threadIdx.x - 256;  // row inside the CTA's 128*NB rows; packed rows are 32 B apart, bf16 sub-tiles DEQ_TILE apart
    int grow = n0 + r;
    if (grow >= P.n) grow = P.n - 1;  // rows past N are zero-filled by TMA; keep the scale loads in bounds
    const uint8_t* rowp = P.rows + (size_t)grow * P.pitch;
    const int swz = r & 7;
    unsigned char* drow_base = nullptr;
    for (int kb = 0; kb < num_kb; ++kb) {
      const int sp = kb % SP, sd = kb % SD;
      // group scale / zero point of the two 32-element halves of this k block
      int g0 = ((kb0 + kb) * BLOCK_K) / P.group, g1 = ((kb0 + kb) * BLOCK_K + 32) / P.group;
      if (g0 >= P.ngroups) g0 = P.ngroups - 1;
      if (g1 >= P.ngroups) g1 = P.ngroups - 1;
      const float sc0 = ns_scale_at(rowp + P.sc_off, P.stype, g0);
      const float sc1 = (g1 == g0) ? sc0 : ns_scale_at(rowp + P.sc_off, P.stype, g1);
      // offset subtracted in bf16 before the scale: 128 (bf16 magic) + 8 (nibble bias) [+ zero point] for packed int4,
      // the zero point alone for int8 weights (all exactly representable)
      float of0 = W8 ? 0.f : 136.f, of1 = of0;
      if (P.asym) {
        of0 += (float)(int)(signed char)rowp[P.zp_off + g0];
        of1 += (float)(int)(signed char)rowp[P.zp_off + g1];
      }
      const __nv_bfloat162 s2[2] = {__float2bfloat162_rn(sc0), __float2bfloat162_rn(sc1)};
      const __nv_bfloat162 o2[2] = {__float2bfloat162_rn(of0), __float2bfloat162_rn(of1)};

      mbar_wait(&p_full[sp], (kb / SP) & 1);
      const uint4* pk = reinterpret_cast<const uint4*>(packed + sp * PACKED_STAGE + r * ROW_BYTES);
      uint32_t words[W8 ? 16 : 8];
#pragma unroll
      for (int i = 0; i < (W8 ? 4 : 2); ++i) {
        const uint4 qv = pk[i];
        words[4 * i] = qv.x, words[4 * i + 1] = qv.y, words[4 * i + 2] = qv.z, words[4 * i + 3] = qv.w;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_empty[sp]);  // packed bytes are in registers
      if (kb >= SD) mbar_wait(&d_empty[sd], ((kb / SD) - 1) & 1);
      drow_base = deq + sd * DEQ_STAGE + (r >> 7) * DEQ_TILE + ((r & 127) >> 3) * 1024 + swz * 128;
      if (!(P.dbg & 1))
#pragma unroll
      for (int c = 0; c < 8; ++c) {  // 16-byte chunk c of the bf16 row = k 8c..8c+7
        const int h = c >> 2;
        uint32_t o[4];
        if (W8) {
          // int8 weights, natural byte order: words[2c] = k 8c..8c+3, words[2c+1] = k 8c+4..8c+7
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t wsrc = words[2 * c + (j >> 1)];
            const int b0 = (int)(signed char)((wsrc >> (16 * (j & 1))) & 0xffu), b1 = (int)(signed char)((wsrc >> (16 * (j & 1) + 8)) & 0xffu);
            __nv_bfloat162 v = __floats2bfloat162_rn((float)b0, (float)b1);
            v = __hmul2(__hsub2(v, o2[h]), s2[h]);
            o[j] = *reinterpret_cast<uint32_t*>(&v);
          }
        } else if (P.wfmt == NS_W_NF4) {
          // NF4 codes: level from the table (kernel_ref.h:1325-1368), rounded to bf16, times the bf16 scale
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float l0 = nf4_lut[(words[c] >> (4 * j)) & 0xFu], l1 = nf4_lut[(words[c] >> (4 * j + 16)) & 0xFu];
            __nv_bfloat162 v = __hmul2(__floats2bfloat162_rn(l0, l1), s2[h]);
            o[j] = *reinterpret_cast<uint32_t*>(&v);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t t = ((words[c] >> (4 * j)) & 0x000F000Fu) | 0x43004300u;  // bf16x2 (128 + e(2j), 128 + e(2j+1))
            __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&t);
            v = __hmul2(__hsub2(v, o2[h]), s2[h]);
            o[j] = *reinterpret_cast<uint32_t*>(&v);
          }
        }
        *reinterpret_cast<uint4*>(drow_base + ((c ^ swz) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA
      __syncwarp();
      if (lane == 0) mbar_arrive(&d_full[sd]);
    }
  }
  // ============================ epilogue: TMEM -> registers -> global, ALL warps ============================
  // A warp can read the TMEM lanes of quarter (warp % 4); the warps of a quarter split the accumulator columns.  With only
  // four epilogue warps the 128 x 512 fp32 tile took 27 us of a 112 us GEMM (nothing else runs in the CTA by then).
  {
    __syncwarp();
    const int q = warp & 3;
    const int per_quarter = (int)(blockDim.x >> 7);  // warps per lane quarter: 3 (NB = 1) or 4 (NB = 2)
    pdl_wait();  // dst / residual may still be in use by the preceding kernel
    mbar_wait(tmem_full, 0);
    tcgen05_fence_after();
#pragma unroll 1
    for (int cc = (warp >> 2) * 32; cc < ((P.dbg & 4) ? 0 : NB * T); cc += per_quarter * 32) {
      const int nb = cc / T, c0 = cc % T;
      const int nrow = n0 + nb * BLOCK_N + q * 32 + lane;
      const bool nvalid = nrow < P.n;
      const float bcast_bias = (P.bias && P.bias_bcast && nvalid) ? P.bias[nrow] : 0.f;
      uint32_t v[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cc;
      TMEM_LD_32X32B_X32(taddr, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int t = t0 + c0 + j;
        if (nvalid && t < P.m) {
          const size_t o = (size_t)t * P.ldo + nrow;
          float x = __uint_as_float(v[j]);
          if (blockIdx.z == 0) {
            x += bcast_bias;
            if (P.bias && !P.bias_bcast) x += P.bias[o];
            if (P.residual) x += P.residual[o];
          }
          if (gridDim.z > 1) atomicAdd(P.dst + o, x);
          else P.dst[o] = x;
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(T * NB)) : "memory");
  }
}

// fp32 -> bf16 activations [m][kpad], zero padded, optional column gather
__global__ void __launch_bounds__(256) act_to_bf16_kernel(const float* __restrict__ A, int lda, int M, int K, int kpad,
                                                          const int* __restrict__ shuffle, __nv_bfloat16* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per 2 elements
  const size_t total = (size_t)M * (kpad >> 1);
  if (idx >= total) return;
  const int m = (int)(idx / (kpad >> 1)), k = (int)(idx - (size_t)m * (kpad >> 1)) * 2;
  float a = 0.f, b = 0.f;
  if (k < K) a = A[(size_t)m * lda + (shuffle ? shuffle[k] : k)];
  if (k + 1 < K) b = A[(size_t)m * lda + (shuffle ? shuffle[k + 1] : k + 1)];
  reinterpret_cast<__nv_bfloat162*>(out)[idx] = __floats2bfloat162_rn(a, b);
}
// same, 8 elements per thread (2 x 16-byte loads, one 16-byte store): K % 8 == 0, K == kpad, lda % 4 == 0, no gather.
// The scalar version above moved 1.6 TB/s; a 2048-token prefill converts 0.35 GB per layer.
__global__ void __launch_bounds__(256) act_to_bf16_v8_kernel(const float* __restrict__ A, int lda, int M, int K,
                                                             __nv_bfloat16* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int k8 = K >> 3;
  const size_t total = (size_t)M * k8;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int m = (int)(idx / k8), c = (int)(idx - (size_t)m * k8);
    const float4* src = (const float4*)(A + (size_t)m * lda) + 2 * c;
    const float4 x = src[0], y = src[1];
    __nv_bfloat162 p0 = __floats2bfloat162_rn(x.x, x.y), p1 = __floats2bfloat162_rn(x.z, x.w);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(y.x, y.y), p3 = __floats2bfloat162_rn(y.z, y.w);
    uint4 o;
    o.x = *(uint32_t*)&p0, o.y = *(uint32_t*)&p1, o.z = *(uint32_t*)&p2, o.w = *(uint32_t*)&p3;
    ((uint4*)out)[idx] = o;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

template <int T, int NB, bool W8>
int launch_t(const CUtensorMap& mw, const CUtensorMap& ma, const GemmParams& P, cudaStream_t st) {
  using L = Smem<T, NB, W8>;
  auto kern = gemm_w4_tc_kernel<T, NB, W8>;
  static bool attr_set = false;
  if (!attr_set) {
    NS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total));
    attr_set = true;
  }
  dim3 grid((P.n + BLOCK_N * NB - 1) / (BLOCK_N * NB), (P.m + T - 1) / T);
  GemmParams Q = P;
  const int total_kb = (P.kpad + BLOCK_K - 1) / BLOCK_K;
  int splits = 1;
  const int tiles = (int)(grid.x * grid.y);
  static const int env_splits = getenv("NS_TC_SPLITS") ? atoi(getenv("NS_TC_SPLITS")) : 0;  // tuning aid
  if (tiles < 96 && P.residual != P.dst && P.bias != P.dst) {
    splits = ns_num_sms() / tiles;  // floor: one CTA per SM is resident (TMEM + ~100-200 KB smem), a partial second wave doubles the time
    if (splits > total_kb / 8) splits = total_kb / 8;  // keep >= 8 k blocks (512 k) per slice
    if (splits > 16) splits = 16;
    if (splits < 1) splits = 1;
  }
  if (env_splits > 0) splits = env_splits;
  Q.kb_per_split = (total_kb + splits - 1) / splits;
  splits = (total_kb + Q.kb_per_split - 1) / Q.kb_per_split;
  grid.z = (unsigned)splits;
  if (splits > 1) NS_CUDA_TRY(cudaMemset2DAsync(P.dst, (size_t)P.ldo * 4, 0, (size_t)P.n * 4, (size_t)P.m, st));
  NS_CUDA_TRY(ns_launch_pdl(kern, grid, dim3(L::kThreads), (size_t)L::total, st, mw, ma, Q));
  ns_count_launch();
  return NS_OK;
}

}  // namespace

size_t ns_gemm_tc_workspace_bytes(int m, int kpad) { return ns_round_up((size_t)m * kpad * 2, 256); }

bool ns_gemm_tc_supported(const ns_weight* w) {
  return (w->wfmt == NS_W_S4 || w->wfmt == NS_W_NF4 || w->wfmt == NS_W_S8) && (w->group % 32 == 0 || w->group == w->k);
}

// phase 1: fp32 activations -> bf16 [m][kpad] in ws (ns_gemm_tc_workspace_bytes(m, kpad) bytes)
int ns_launch_act_bf16(const ns_weight* w, const float* act, int lda, int m, void* ws, cudaStream_t st) {
  if (!w->shuffle && w->k == w->kpad && (w->k & 7) == 0 && (lda & 3) == 0 && ((uintptr_t)act & 15) == 0) {
    const size_t total8 = (size_t)m * (w->k >> 3);
    size_t blocks = (total8 + 255) / 256;
    if (blocks > (size_t)ns_num_sms() * 16) blocks = (size_t)ns_num_sms() * 16;
    NS_CUDA_TRY(ns_launch_pdl(act_to_bf16_v8_kernel, dim3((unsigned)blocks), dim3(256), 0, st, act, lda, m, w->k, (__nv_bfloat16*)ws));
    ns_count_launch();
    return NS_OK;
  }
  const size_t total = (size_t)m * (w->kpad >> 1);
  NS_CUDA_TRY(ns_launch_pdl(act_to_bf16_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, act, lda, m, w->k,
                            w->kpad, (const int*)w->shuffle, (__nv_bfloat16*)ws));
  ns_count_launch();
  return NS_OK;
}

// silu(gate) * up, elementwise (epilogues Swish alpha=-1 + Mul of ip_fusion_ffn.cpp:408-470, kernel_ref.h:1574); 4 elements per
// thread when the pointers and the count allow it
template <bool V4>
__global__ void __launch_bounds__(256) silu_mul_kernel(const float* __restrict__ g, const float* __restrict__ u,
                                                       float* __restrict__ out, float* __restrict__ aux, size_t total,
                                                       int eltop) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t n = V4 ? total >> 2 : total;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    if (V4) {
      const float4 x = ((const float4*)g)[i], y = ((const float4*)u)[i];
      float4 sg;
      sg.x = eltop == NS_ELT_GELU ? ns_gelu(x.x) : ns_silu(x.x);
      sg.y = eltop == NS_ELT_GELU ? ns_gelu(x.y) : ns_silu(x.y);
      sg.z = eltop == NS_ELT_GELU ? ns_gelu(x.z) : ns_silu(x.z);
      sg.w = eltop == NS_ELT_GELU ? ns_gelu(x.w) : ns_silu(x.w);
      if (aux) ((float4*)aux)[i] = sg;
      ((float4*)out)[i] = make_float4(sg.x * y.x, sg.y * y.y, sg.z * y.z, sg.w * y.w);
    } else {
      const float x = g[i];
      const float sg = eltop == NS_ELT_GELU ? ns_gelu(x) : ns_silu(x);
      if (aux) aux[i] = sg;
      out[i] = sg * u[i];
    }
  }
}
// silu(gate) * up written straight into the bf16 activation image of the down projection: the fp32 product (m x fmid, 90 MB at 2048
// tokens of Llama-2-7B) is neither stored nor read back.  Same values as silu_mul_kernel followed by act_to_bf16_v8_kernel (the product
// is formed in fp32 and rounded once).  Used when the caller does not look at the intermediate (the eval step).
__global__ void __launch_bounds__(256) silu_mul_bf16_kernel(const float* __restrict__ g, const float* __restrict__ u, size_t total8,
                                                            __nv_bfloat16* __restrict__ out, int eltop) {
  pdl_launch_dependents();
  pdl_wait();
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total8; idx += (size_t)gridDim.x * blockDim.x) {
    const float4 x0 = ((const float4*)g)[2 * idx], x1 = ((const float4*)g)[2 * idx + 1];
    const float4 y0 = ((const float4*)u)[2 * idx], y1 = ((const float4*)u)[2 * idx + 1];
    const float xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
    const float ys[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
    float p[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = (eltop == NS_ELT_GELU ? ns_gelu(xs[i]) : ns_silu(xs[i])) * ys[i];
    __nv_bfloat162 q0 = __floats2bfloat162_rn(p[0], p[1]), q1 = __floats2bfloat162_rn(p[2], p[3]);
    __nv_bfloat162 q2 = __floats2bfloat162_rn(p[4], p[5]), q3 = __floats2bfloat162_rn(p[6], p[7]);
    uint4 o;
    o.x = *(uint32_t*)&q0, o.y = *(uint32_t*)&q1, o.z = *(uint32_t*)&q2, o.w = *(uint32_t*)&q3;
    ((uint4*)out)[idx] = o;
  }
}
// g, u: [m][w2->k] fp32 contiguous.  false: the layout needs the general conversion kernel (act-order gather, padded K) -- the caller
// takes the two-step path
bool ns_launch_silu_mul_bf16(const ns_weight* w2, const float* g, const float* u, int m, void* ws, cudaStream_t st, int eltop, int* rc) {
  if (w2->shuffle || w2->k != w2->kpad || (w2->k & 7) || (((uintptr_t)g | (uintptr_t)u | (uintptr_t)ws) & 15)) return false;
  const size_t total8 = (size_t)m * (w2->k >> 3);
  size_t blocks = (total8 + 255) / 256;
  if (blocks > (size_t)ns_num_sms() * 16) blocks = (size_t)ns_num_sms() * 16;
  if (blocks < 1) blocks = 1;
  *rc = ns_cuda_ok(ns_launch_pdl(silu_mul_bf16_kernel, dim3((unsigned)blocks), dim3(256), 0, st, g, u, total8, (__nv_bfloat16*)ws, eltop),
                   "silu_mul_bf16_kernel")
            ? NS_OK
            : NS_E_CUDA;
  ns_count_launch();
  return true;
}
int ns_launch_silu_mul(const float* g, const float* u, float* out, float* aux, size_t total, cudaStream_t st, int eltop) {
  const bool v4 = (total & 3) == 0 && (((uintptr_t)g | (uintptr_t)u | (uintptr_t)out | (uintptr_t)aux) & 15) == 0;
  const size_t n = v4 ? total >> 2 : total;
  size_t blocks = (n + 255) / 256;
  if (blocks > (size_t)ns_num_sms() * 16) blocks = (size_t)ns_num_sms() * 16;
  if (blocks < 1) blocks = 1;
  if (v4)
    NS_CUDA_TRY(ns_launch_pdl(silu_mul_kernel<true>, dim3((unsigned)blocks), dim3(256), 0, st, g, u, out, aux, total, eltop));
  else
    NS_CUDA_TRY(ns_launch_pdl(silu_mul_kernel<false>, dim3((unsigned)blocks), dim3(256), 0, st, g, u, out, aux, total, eltop));
  ns_count_launch();
  return NS_OK;
}
__global__ void __launch_bounds__(256) gelu_kernel(float* __restrict__ x, size_t total) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) x[i] = ns_gelu(x[i]);
}
int ns_launch_gelu(float* x, size_t total, cudaStream_t st) {
  NS_CUDA_TRY(ns_launch_pdl(gelu_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, x, total));
  ns_count_launch();
  return NS_OK;
}

// phase 2: weights x bf16 activations already in ws
int ns_launch_gemm_tc(const ns_weight* w, const void* ws, float* dst, int ldo, int m, const float* bias, int bias_bcast,
                      const float* residual, cudaStream_t st) {
  if (!ns_gemm_tc_supported(w)) {
    ns_set_error("tensor-core GEMM: int4 / NF4 / int8 weights with 32-multiple groups are supported");
    return NS_E_UNSUPPORTED;
  }
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    ns_set_error("cuTensorMapEncodeTiled not available from the driver");
    return NS_E_CUDA;
  }
  const __nv_bfloat16* abf = (const __nv_bfloat16*)ws;
  int T = m <= 32 ? 32 : (m <= 64 ? 64 : (m <= 128 ? 128 : 256));
  static const int force_nb = getenv("NS_TC_NB") ? atoi(getenv("NS_TC_NB")) : 0;  // tuning aid
  const int NBsel = force_nb ? force_nb : ((T >= 128 && w->n >= 2 * BLOCK_N) ? 2 : 1);
  const bool w8 = w->wfmt == NS_W_S8;
  CUtensorMap mw, ma;
  {
    // packed nibbles: uint8 [n][q_bytes] with row pitch `pitch`; box = 32 bytes (64 k) x 128 rows, no swizzle
    cuuint64_t dims[2] = {(cuuint64_t)w->q_bytes, (cuuint64_t)w->n};
    cuuint64_t strides[1] = {(cuuint64_t)w->pitch};
    cuuint32_t box[2] = {(cuuint32_t)(w8 ? BLOCK_K : BLOCK_K / 2), (cuuint32_t)(BLOCK_N * NBsel)};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&mw, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)w->rows, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      ns_set_error("cuTensorMapEncodeTiled(weights) failed: %d", (int)r);
      return NS_E_CUDA;
    }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)w->kpad, (cuuint64_t)m};
    cuuint64_t strides[1] = {(cuuint64_t)w->kpad * 2};
    cuuint32_t box[2] = {BLOCK_K, (cuuint32_t)T};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)abf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      ns_set_error("cuTensorMapEncodeTiled(activations) failed: %d", (int)r);
      return NS_E_CUDA;
    }
  }
  GemmParams P;
  P.rows = w->rows;
  P.wfmt = w->wfmt;
  P.f4kind = w->f4kind;
  P.pitch = w->pitch;
  P.sc_off = w->sc_off;
  P.zp_off = w->zp_off;
  P.stype = w->stype;
  P.asym = w->asym;
  P.group = w->group;
  P.ngroups = w->ngroups;
  P.n = w->n;
  P.k = w->k;
  P.kpad = w->kpad;
  P.m = m;
  P.dst = dst;
  P.ldo = ldo;
  P.bias = bias;
  P.bias_bcast = bias_bcast;
  P.residual = residual;
  static const int dbg = getenv("NS_TC_DEBUG") ? atoi(getenv("NS_TC_DEBUG")) : 0;
  P.dbg = dbg;
  if (w8) {
    if (NBsel == 2) return T == 128 ? launch_t<128, 2, true>(mw, ma, P, st) : launch_t<256, 2, true>(mw, ma, P, st);
    switch (T) {
      case 32: return launch_t<32, 1, true>(mw, ma, P, st);
      case 64: return launch_t<64, 1, true>(mw, ma, P, st);
      case 128: return launch_t<128, 1, true>(mw, ma, P, st);
      default: return launch_t<256, 1, true>(mw, ma, P, st);
    }
  }
  if (NBsel == 2) return T == 128 ? launch_t<128, 2, false>(mw, ma, P, st) : launch_t<256, 2, false>(mw, ma, P, st);
  switch (T) {
    case 32: return launch_t<32, 1, false>(mw, ma, P, st);
    case 64: return launch_t<64, 1, false>(mw, ma, P, st);
    case 128: return launch_t<128, 1, false>(mw, ma, P, st);
    default: return launch_t<256, 1, false>(mw, ma, P, st);
  }
}
