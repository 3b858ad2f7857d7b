// comm.cu -- one-shot sum all-reduce over NVLink peer memory for the tensor-parallel exchange step.
//
// Replaces reduce_add / ne_compute_forward_all_reduce (core/parallel_context.cpp:47, core/ne_layers.c:5466: oneCCL / shm
// all-reduce of the [M, n_embd] fp32 partials after the o-projection and the down-projection, models/llama/llama.cpp:592,693).
// At decode the message is 16-32 KB: the cost is latency, not bandwidth, so instead of a ring the kernel does a one-shot
// exchange: every rank stores its vector straight into a slot of EVERY peer's buffer through NVLink (P2P stores, NVSwitch
// gives full bandwidth to every peer), raises a per-source flag with system-scope release, waits for the other sources'
// flags, and adds the `world` slots in rank order (+ optional residual).  The sum order is the same on every rank, so all
// ranks hold bit-identical results (the reference's ring/shm reduction does not guarantee that).
//
// Memory: per rank ONE device allocation shared with the peers by cudaIpc handles (exchanged by the host through
// torch.distributed, neural_speed_b200/tp.py):  [2 parities][world][max_elems] floats | [2][world] arrival counters.
// Two parities make step s+1 safe while a slow peer still reads step s; a rank cannot reach step s+2 before every peer
// has finished step s (it needs their step s+1 data, which they send after completing s).
#include <vector>

#include "nsb.cuh"

namespace {
constexpr int kCtas = 8;       // CTAs per all-reduce (fixed: arrival counters count CTAs)
constexpr int kThreads = 256;
constexpr int kMaxWorld = 16;

struct Peers {
  float* slots[kMaxWorld];          // base of every rank's buffer (own mapping for self)
  unsigned int* flags[kMaxWorld];   // arrival counters of every rank's buffer
};

__device__ __forceinline__ void st_release_sys_add(unsigned int* p) {
  asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// data[i] = sum_r partial_r[i] (+ residual[i]); n % 4 == 0
// The step number lives in device memory (ctl[0]; ctl[1] = CTAs done), so the launch has no per-step arguments and can be
// replayed from a CUDA graph: every CTA reads the step before it arrives anywhere, and no CTA can finish before all local CTAs
// have arrived, so the last CTA to finish may safely advance it.
__global__ void __launch_bounds__(kThreads) allreduce_oneshot_kernel(Peers P, int rank, int world, size_t max_elems,
                                                                     unsigned int* __restrict__ ctl, float* __restrict__ data, size_t n,
                                                                     const float* __restrict__ residual) {
  pdl_launch_dependents();
  pdl_wait();
  const unsigned int step = *((volatile unsigned int*)ctl) + 1u;
  const int parity = (int)(step & 1u);
  const unsigned int expect = ((step + 1u) / 2u) * kCtas;  // arrivals per (parity, source) up to and including this step
  const size_t n4 = n >> 2;
  const size_t per = (n4 + gridDim.x - 1) / gridDim.x;
  const size_t lo = blockIdx.x * per, hi = (lo + per < n4) ? lo + per : n4;
  // 1. scatter my slice to slot[parity][rank] of every rank (self included: one code path, one sum order)
  const size_t slot_off = ((size_t)parity * world + rank) * max_elems;
  for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const float4 v = ((const float4*)data)[i];
    for (int p = 0; p < world; ++p) ((float4*)(P.slots[p] + slot_off))[i] = v;
  }
  __syncthreads();
  // 2. publish: one arrival per CTA and destination (release at system scope orders the stores above)
  if (threadIdx.x < world) {
    __threadfence_system();
    st_release_sys_add(P.flags[threadIdx.x] + parity * world + rank);
  }
  // 3. wait until every source's kCtas CTAs have arrived for this step
  if (threadIdx.x < world) {
    const unsigned int* f = P.flags[rank] + parity * world + threadIdx.x;
    long long spins = 0;
    while (ld_acquire_sys(f) < expect) {
      if (++spins > (1ll << 28)) {  // a peer never arrived (mismatched call sequence): give up instead of hanging the GPU
        P.flags[rank][2 * world] = 0xdeadu;
        break;
      }
    }
  }
  __syncthreads();
  // 4. reduce my slice in rank order
  const float* mine = P.slots[rank] + (size_t)parity * world * max_elems;
  for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    float4 acc = ((const float4*)mine)[i];
    for (int r = 1; r < world; ++r) {
      const float4 v = ((const float4*)(mine + (size_t)r * max_elems))[i];
      acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
    }
    if (residual) {
      const float4 v = ((const float4*)residual)[i];
      acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
    }
    ((float4*)data)[i] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(ctl + 1, 1u) == kCtas - 1) {
      ctl[1] = 0u;
      __threadfence();
      *((volatile unsigned int*)ctl) = step;
    }
  }
}
}  // namespace

struct ns_comm {
  int rank, world;
  size_t max_elems;
  void* local = nullptr;  // own buffer
  size_t bytes = 0;
  void* peer_base[kMaxWorld] = {};
  bool opened[kMaxWorld] = {};
  Peers peers;
  unsigned int* ctl = nullptr;  // device: {step, CTAs done}
  bool ready = false;
};

extern "C" size_t ns_comm_handle_bytes(void) { return sizeof(cudaIpcMemHandle_t); }

extern "C" ns_comm* ns_comm_create(int rank, int world, size_t max_elems, void* queue) {
  if (ns_ensure_device()) return nullptr;
  if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world || max_elems == 0) {
    ns_set_error("ns_comm_create: invalid arguments (rank %d world %d)", rank, world);
    return nullptr;
  }
  ns_comm* c = new ns_comm();
  c->rank = rank;
  c->world = world;
  c->max_elems = ns_round_up(max_elems, 4);
  const size_t slot_bytes = (size_t)2 * world * c->max_elems * sizeof(float);
  c->bytes = ns_round_up(slot_bytes, 256) + 256;
  if (cudaMalloc(&c->local, c->bytes) != cudaSuccess) {
    ns_set_error("ns_comm_create: cudaMalloc(%zu) failed", c->bytes);
    delete c;
    return nullptr;
  }
  cudaStream_t st = ns_stream_of(queue);
  cudaMemsetAsync(c->local, 0, c->bytes, st);
  cudaStreamSynchronize(st);
  return c;
}

extern "C" int ns_comm_get_handle(ns_comm* c, void* handle_out) {
  if (!c || !handle_out) return NS_E_INVALID;
  cudaIpcMemHandle_t h;
  NS_CUDA_TRY(cudaIpcGetMemHandle(&h, c->local));
  memcpy(handle_out, &h, sizeof(h));
  return NS_OK;
}

// all_handles: world * ns_comm_handle_bytes() bytes, rank-major (what an all_gather of ns_comm_get_handle produces)
extern "C" int ns_comm_open_peers(ns_comm* c, const void* all_handles) {
  if (!c || !all_handles) return NS_E_INVALID;
  const size_t slot_bytes = ns_round_up((size_t)2 * c->world * c->max_elems * sizeof(float), 256);
  for (int r = 0; r < c->world; ++r) {
    void* base = c->local;
    if (r != c->rank) {
      cudaIpcMemHandle_t h;
      memcpy(&h, (const char*)all_handles + (size_t)r * sizeof(h), sizeof(h));
      if (!ns_cuda_ok(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle")) return NS_E_CUDA;
      c->opened[r] = true;
    }
    c->peer_base[r] = base;
    c->peers.slots[r] = (float*)base;
    c->peers.flags[r] = (unsigned int*)((char*)base + slot_bytes);
  }
  c->ctl = c->peers.flags[c->rank] + 2 * c->world + 1;  // after the counters and the status word
  c->ready = true;
  return NS_OK;
}

// Same-process wiring (several ranks' communicators living in one process, e.g. a loopback test of the exchange kernel with all
// ranks on ONE device, or a single-process multi-GPU host with peer access enabled): the ranks' buffers are plain device
// pointers, no cudaIpc handles.  comms[r] must be rank r of the same world, all with the same max_elems.
extern "C" int ns_comm_link_local(ns_comm* const* comms, int world) {
  if (!comms || world < 1 || world > kMaxWorld) return NS_E_INVALID;
  for (int r = 0; r < world; ++r)
    if (!comms[r] || comms[r]->rank != r || comms[r]->world != world || comms[r]->max_elems != comms[0]->max_elems) {
      ns_set_error("ns_comm_link_local: communicator %d does not match (rank / world / max_elems)", r);
      return NS_E_INVALID;
    }
  const size_t slot_bytes = ns_round_up((size_t)2 * world * comms[0]->max_elems * sizeof(float), 256);
  for (int me = 0; me < world; ++me) {
    ns_comm* c = comms[me];
    for (int r = 0; r < world; ++r) {
      c->peer_base[r] = comms[r]->local;
      c->peers.slots[r] = (float*)comms[r]->local;
      c->peers.flags[r] = (unsigned int*)((char*)comms[r]->local + slot_bytes);
    }
    c->ctl = c->peers.flags[me] + 2 * world + 1;
    c->ready = true;
  }
  return NS_OK;
}

// in-place: data[0..n) = sum over ranks (+ residual).  Every rank must call it with the same n, in the same order.
extern "C" int ns_comm_all_reduce_f32(ns_comm* c, float* data, size_t n, const float* residual, void* queue) {
  if (!c || !c->ready || !data || n == 0 || n > c->max_elems || (n & 3) || ((uintptr_t)data & 15) || ((uintptr_t)residual & 15)) {
    ns_set_error("ns_comm_all_reduce_f32: invalid arguments (n=%zu, max %zu; n %% 4 == 0 and 16-byte alignment required)", n,
                 c ? c->max_elems : (size_t)0);
    return NS_E_INVALID;
  }
  NS_CUDA_TRY(ns_launch_pdl(allreduce_oneshot_kernel, dim3(kCtas), dim3(kThreads), 0, ns_stream_of(queue), c->peers, c->rank, c->world,
                            c->max_elems, c->ctl, data, n, residual));
  ns_count_launch();
  return NS_OK;
}

// 0 = healthy; non-zero once a wait gave up (results of that step are undefined)
extern "C" int ns_comm_status(ns_comm* c) {
  if (!c || !c->ready) return NS_E_INVALID;
  unsigned int v = 0;
  NS_CUDA_TRY(cudaMemcpy(&v, c->peers.flags[c->rank] + 2 * c->world, sizeof(v), cudaMemcpyDeviceToHost));
  return (int)v;
}

extern "C" void ns_comm_free(ns_comm* c) {
  if (!c) return;
  cudaDeviceSynchronize();
  for (int r = 0; r < c->world; ++r)
    if (c->opened[r]) cudaIpcCloseMemHandle(c->peer_base[r]);
  if (c->local) cudaFree(c->local);
  delete c;
}
