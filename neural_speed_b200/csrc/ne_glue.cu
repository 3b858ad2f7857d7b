// ne_glue.cu -- the entry points of neural_speed/core/ne_bestla.h that take the graph engine's own structs, so the reference's
// ne_graph_compute (core/ne_layers.c:11915-12010) can run on libns_b200.so with NO source change:
//   bestla_support          core/layers/ne_bestla.cpp:205-292   which nodes the kernel library takes, host workspace, n_tasks = 1
//   bestla_backend_support  core/layers/ne_bestla.cpp:176-203   backend of a node's result
//   bestla_parallel_for     core/layers/ne_bestla.cpp:42-70     INIT / COMPUTE / FINALIZE over the node's tasks
//   bestla_mul / bestla_add core/layers/ne_bestla.cpp:119-166   contiguous element-wise ops (called by ne_layers.c:4622,5677)
//   bestla_layernormalization  core/layers/ne_bestla.cpp:114-117 -> kernel_ref.h:2199-2250 (called by ne_layers.c:6541,6625)
// The struct layouts come from include/ns_ne_abi.h (restated from ne.h; tests/test_ne_abi_cpu.py pins every offset against the
// reference's header).  The element-wise entry points receive HOST buffers like the matmul drop-ins: they stage through device
// memory and run a CUDA kernel -- there is no CPU compute path in this library.
#include <omp.h>

#include "../../include/ns_ne_abi.h"
#include "nsb.cuh"

namespace {

struct Stage {
  float* p = nullptr;
  size_t elems = 0;
};
Stage g_a, g_b, g_o;
bool reserve(Stage& s, size_t need) {
  if (s.elems >= need) return true;
  if (s.p) cudaFree(s.p);
  s.p = nullptr;
  s.elems = 0;
  if (cudaMalloc((void**)&s.p, need * sizeof(float)) != cudaSuccess) return false;
  s.elems = need;
  return true;
}

// out[b][i] = t[b][i] (op) v[b * vstep + i]     (vstep == 0: one vector broadcast over the batch)
template <bool MUL>
__global__ void __launch_bounds__(256) eltwise_kernel(const float* __restrict__ t, const float* __restrict__ v, float* __restrict__ out,
                                                      int batch, int vsize, int vstep) {
  const size_t total = (size_t)batch * vsize;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t b = e / vsize, i = e - b * vsize;
    const float x = t[e], y = v[b * vstep + i];
    out[e] = MUL ? x * y : x + y;
  }
}

// kernel_ref.h:2199-2250 without scale / bias: one CTA per row; `simplified` = RMS norm
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ in, float* __restrict__ out, int n, float eps, int simplified) {
  const float* x = in + (size_t)blockIdx.x * n;
  float* y = out + (size_t)blockIdx.x * n;
  float s = 0.f, ss = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) {
    const float v = x[i];
    s += v;
    ss = fmaf(v, v, ss);
  }
  __shared__ float rs[8], rss[8];
  s = warp_sum(s);
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) {
    rs[threadIdx.x >> 5] = s;
    rss[threadIdx.x >> 5] = ss;
  }
  __syncthreads();
  float ts = 0.f, tss = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    ts += rs[i];
    tss += rss[i];
  }
  const float mean = ts / (float)n;
  const float var = simplified ? tss / (float)n + eps : tss / (float)n - mean * mean + eps;
  const float inv = 1.f / sqrtf(var);
  for (int i = threadIdx.x; i < n; i += 256) y[i] = simplified ? x[i] * inv : (x[i] - mean) * inv;
}

template <bool MUL>
void host_eltwise(int batch, int vsize, const float* tensor, const float* vector, int vstep, float* out) {
  if (ns_ensure_device()) ns_fatal("bestla_%s: no CUDA device", MUL ? "mul" : "add");
  if (batch <= 0 || vsize <= 0) return;
  cudaStream_t st = ns_stream_of(nullptr);
  const size_t nt = (size_t)batch * vsize, nv = vstep ? (size_t)(batch - 1) * vstep + vsize : (size_t)vsize;
  if (!reserve(g_a, nt) || !reserve(g_b, nv) || !reserve(g_o, nt)) ns_fatal("device staging allocation failed");
  if (cudaMemcpyAsync(g_a.p, tensor, nt * 4, cudaMemcpyHostToDevice, st) != cudaSuccess ||
      cudaMemcpyAsync(g_b.p, vector, nv * 4, cudaMemcpyHostToDevice, st) != cudaSuccess)
    ns_fatal("H2D failed");
  const int grid = (int)((nt + 255) / 256 < 2048 ? (nt + 255) / 256 : 2048);
  eltwise_kernel<MUL><<<grid, 256, 0, st>>>(g_a.p, g_b.p, g_o.p, batch, vsize, vstep);
  ns_count_launch();
  if (cudaMemcpyAsync(out, g_o.p, nt * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)
    ns_fatal("element-wise kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
}

inline bool ne_contiguous(const ns_ne_tensor* t) { return t->nb[0] <= t->nb[1] && t->nb[1] <= t->nb[2] && t->nb[2] <= t->nb[3]; }
inline long long ne_rows(const ns_ne_tensor* t) { return (long long)(t->ne[1] * t->ne[2] * t->ne[3]); }

}  // namespace

extern "C" {

NS_API void bestla_mul(int batch, int vsize, const float* tensor, const float* vector, int vstep, float* out) {
  host_eltwise<true>(batch, vsize, tensor, vector, vstep, out);
}
NS_API void bestla_add(int batch, int vsize, const float* tensor, const float* vector, int vstep, float* out) {
  host_eltwise<false>(batch, vsize, tensor, vector, vstep, out);
}

NS_API void bestla_layernormalization(int norm_count, int norm_size, bool isrms, float epsilon, const float* FpIn, float* FpOut) {
  if (ns_ensure_device()) ns_fatal("bestla_layernormalization: no CUDA device");
  if (norm_count <= 0 || norm_size <= 0) return;
  cudaStream_t st = ns_stream_of(nullptr);
  const size_t n = (size_t)norm_count * norm_size;
  if (!reserve(g_a, n) || !reserve(g_o, n)) ns_fatal("device staging allocation failed");
  if (cudaMemcpyAsync(g_a.p, FpIn, n * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) ns_fatal("H2D failed");
  layernorm_kernel<<<norm_count, 256, 0, st>>>(g_a.p, g_o.p, norm_size, epsilon, isrms ? 1 : 0);
  ns_count_launch();
  if (cudaMemcpyAsync(FpOut, g_o.p, n * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)
    ns_fatal("layernorm kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
}

// ne_bestla.cpp:176-203 outside NS_SYCL: every result lives on the host (the device-resident route is INTEGRATION.md B)
NS_API int bestla_backend_support(struct ns_ne_tensor* src0, struct ns_ne_tensor* src1, int op) {
  (void)src0;
  (void)src1;
  (void)op;
  return NS_NE_BACKEND_CPU;
}

// ne_bestla.cpp:205-292: same decisions as the reference -- the nodes answered `true` are entered ONCE (n_tasks = 1) and the
// callee parallelises internally (here: on the GPU)
NS_API bool bestla_support(struct ns_ne_tensor* node, int n_threads, size_t* workspace, size_t* dev_workspace) {
  (void)n_threads;
  size_t ws_h = 0;
  bool support = node->backend == NS_NE_BACKEND_SYCL;
  switch (node->op) {
    case NS_NE_OP_MUL_MAT_ID:
    case NS_NE_OP_MUL_MAT_BIAS:
    case NS_NE_OP_MUL_MAT: {
      const ns_ne_tensor* wei = node->op == NS_NE_OP_MUL_MAT_ID ? node->opt[0] : node->src0;
      if (node->src0->type == NS_NE_TYPE_BTLA) {
        if (node->src0->backend == NS_NE_BACKEND_CPU)
          ws_h = (size_t)bestla_f32f32_get_workspace_size((int)node->src1->ne[1], (int)wei->ne[1], (int)node->src1->ne[0], wei->data);
        support = true;
      }
    } break;
    case NS_NE_OP_ROPE:
      if (node->type == NS_NE_TYPE_BTLA) support = true;
      break;
    case NS_NE_OP_MUL:
    case NS_NE_OP_ADD:
      if (ne_contiguous(node->src1) && ne_contiguous(node->src0) &&
          (ne_rows(node->src1) == 1 || ne_rows(node->src1) == ne_rows(node->src0)) && node->src0->ne[0] == node->src1->ne[0] &&
          node->nb[0] == sizeof(float))
        support = true;
      break;
    case NS_NE_OP_MUL_FFN_SILU:
    case NS_NE_OP_MUL_FFN_GELU:
    case NS_NE_OP_MUL_FFN_GELU_MUL:
    case NS_NE_OP_MUL_FFN_ADD_GELU:
      if (node->src0->backend == NS_NE_BACKEND_CPU) {
        ws_h = (size_t)bestla_fusion_FFN_f32f32_get_workspace_size((int)node->src0->ne[1], (int)node->src0->ne[0], (int)node->src1->ne[1],
                                                                   (int)node->opt[0]->ne[1], node->src1->data, node->opt[0]->data);
        support = true;
      }
      break;
    case NS_NE_OP_MUL_ID_FFN_GELU:
    case NS_NE_OP_MUL_ID_FFN_SILU:
      if (node->src0->backend == NS_NE_BACKEND_CPU) {
        ws_h = (size_t)bestla_fusion_FFN_f32f32_get_workspace_size((int)node->src0->ne[1], (int)node->src0->ne[0], (int)node->opt[0]->ne[1],
                                                                   (int)node->opt[9]->ne[1], node->opt[0]->data, node->opt[9]->data);
        support = true;
      }
      break;
    case NS_NE_OP_MUL_QKV:
      ws_h = (size_t)bestla_fusion_QKV_f32f32_get_workspace_size((int)node->src0->ne[1], (int)node->src1->ne[1], (int)node->src1->ne[0],
                                                                 node->src1->data);
      support = true;
      break;
    case NS_NE_OP_NORM:
    case NS_NE_OP_RMS_NORM:
      if (ne_contiguous(node->src0)) support = true;
      break;
    default: break;
  }
  if (support) node->n_tasks = 1;
  *workspace = ws_h;
  *dev_workspace = 0;
  return support;
}

// ne_bestla.cpp:42-70: nth == 1 runs the three phases inline; otherwise nth host threads, INIT on thread 0, a barrier between
// the phases.  (Nodes this library computes have n_tasks == 1; the threaded branch serves the engine's own ggml-type nodes.)
NS_API void bestla_parallel_for(ns_forward_compute_fptr fcomp, struct ns_ne_compute_params* mainparams, struct ns_ne_tensor* node) {
  if (mainparams->nth <= 1) {
    struct ns_ne_compute_params params = *mainparams;
    params.type = NS_NE_TASK_INIT;
    fcomp(&params, node);
    params.type = NS_NE_TASK_COMPUTE;
    fcomp(&params, node);
    params.type = NS_NE_TASK_FINALIZE;
    fcomp(&params, node);
    return;
  }
  const int nth = mainparams->nth;
#pragma omp parallel num_threads(nth)
  {
    // the team may come out smaller than requested: every task index is still visited, phases stay separated by barriers
    const int team = omp_get_num_threads(), me = omp_get_thread_num();
    struct ns_ne_compute_params params = *mainparams;
    params.type = NS_NE_TASK_INIT;
    params.ith = 0;
    if (me == 0) fcomp(&params, node);
#pragma omp barrier
    params.type = NS_NE_TASK_COMPUTE;
    for (int t = me; t < nth; t += team) {
      params.ith = t;
      fcomp(&params, node);
    }
#pragma omp barrier
    params.type = NS_NE_TASK_FINALIZE;
    for (int t = me; t < nth; t += team) {
      params.ith = t;
      fcomp(&params, node);
    }
  }
}

}  // extern "C"
