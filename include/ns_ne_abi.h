/* ns_ne_abi.h -- the slice of neural-speed's `ne` graph-engine ABI that crosses the kernel boundary.
 *
 * bestla_support / bestla_backend_support / bestla_parallel_for (neural_speed/core/ne_bestla.h:25-27,
 * core/layers/ne_bestla.cpp:42,176,205) take `struct ne_tensor*` / `struct ne_compute_params*`: the callee reads the node's
 * op, type, shape, strides, sources and data pointers and writes n_tasks.  libns_b200.so is built without the reference's
 * headers, so the layouts are restated here (field order and sizes of neural_speed/core/ne.h:161-206 and :242-256, enum values
 * of core/data_types.h:32-55, core/layers/Ops.h:20-105, ne.h:88-91); tests/test_ne_abi_cpu.py compiles this header next to the
 * reference's own ne.h and checks every offset and enum value, so a drift shows up as a test failure, not as a crash.
 */
#ifndef NS_NE_ABI_H
#define NS_NE_ABI_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NS_NE_MAX_DIMS 4        /* ne.h:44 */
#define NS_NE_MAX_OPT 36        /* ne.h:48 */
#define NS_NE_MAX_OP_PARAMS 32  /* ne.h:50 */

/* core/data_types.h:32-55 */
enum ns_ne_type { NS_NE_TYPE_F32 = 0, NS_NE_TYPE_F16 = 1, NS_NE_TYPE_Q4_0 = 2, NS_NE_TYPE_Q6_K = 14, NS_NE_TYPE_BTLA = 19 };
/* ne.h:88-91 */
enum ns_ne_backend { NS_NE_BACKEND_CPU = 0, NS_NE_BACKEND_SYCL = 1 };
/* ne.h:235-239 */
enum ns_ne_task_type { NS_NE_TASK_INIT = 0, NS_NE_TASK_COMPUTE = 1, NS_NE_TASK_FINALIZE = 2 };
/* core/layers/Ops.h:20-105 (only the ops the boundary functions look at) */
enum ns_ne_op {
  NS_NE_OP_ADD = 2,
  NS_NE_OP_MUL = 6,
  NS_NE_OP_NORM = 24,
  NS_NE_OP_RMS_NORM = 25,
  NS_NE_OP_MUL_MAT = 28,
  NS_NE_OP_MUL_MAT_BIAS = 29,
  NS_NE_OP_MUL_MAT_ID = 30,
  NS_NE_OP_ROPE = 46,
  NS_NE_OP_MUL_QKV = 52,
  NS_NE_OP_MUL_FFN_SILU = 53,
  NS_NE_OP_MUL_FFN_GELU = 54,
  NS_NE_OP_MUL_FFN_GELU_MUL = 55,
  NS_NE_OP_MUL_FFN_ADD_GELU = 56,
  NS_NE_OP_MUL_ID_FFN_SILU = 57,
  NS_NE_OP_MUL_ID_FFN_GELU = 58
};

/* struct ne_tensor, ne.h:161-206 */
struct ns_ne_tensor {
  int type;    /* enum ne_type */
  int backend; /* enum ne_backend */
  int n_dims;
  int64_t ne[NS_NE_MAX_DIMS];
  size_t nb[NS_NE_MAX_DIMS];
  int op; /* enum ne_op */
  bool is_param;
  int32_t op_params[NS_NE_MAX_OP_PARAMS / sizeof(int32_t)];
  struct ns_ne_tensor* grad;
  struct ns_ne_tensor* src0;
  struct ns_ne_tensor* src1;
  struct ns_ne_tensor* opt[NS_NE_MAX_OPT];
  int n_tasks;
  int perf_runs;
  int64_t perf_cycles;
  int64_t perf_time_us;
  void* data;
  size_t size;
  char name[32];
  char padding[8];
};

/* struct ne_compute_params, ne.h:242-256 */
struct ns_ne_compute_params {
  int type; /* enum ne_task_type */
  int ith, nth;
  size_t wsize;
  void* wdata;
  size_t dev_wsize;
  void* dev_wdata;
  void* dev_queue;
};

/* forward_compute_fptr, ne_bestla.h:24 */
typedef void (*ns_forward_compute_fptr)(struct ns_ne_compute_params* params, struct ns_ne_tensor* node);

#ifdef __cplusplus
}
#endif
#endif /* NS_NE_ABI_H */
