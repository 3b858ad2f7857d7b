"""One Llama-2-7B layer's prefill GEMMs (M = 2048) through ns_mul_qkv / ns_mul_mat / ns_ffn_silu -- run under ncu to capture
gemm_w4_tc_kernel (profiles/*gemm_tc*)."""
import ctypes as C
import sys

import torch

import neural_speed_b200 as ns

M = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
E, FF = 4096, 11008
L = ns.lib()
L.bestla_init()


def mk(n, k):
    w = torch.randn(n, k, device="cuda") * 0.02
    rows = torch.empty(n * (k // 32) * 18, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    assert L.ns_device_quantize_q4_0(C.c_void_p(w.data_ptr()), C.c_void_p(rows.data_ptr()), n, k, None) == 0
    L.bestla_device_sync(None)
    return ns.Weight.from_q4_0_device(rows.data_ptr(), n, k, (k // 32) * 18)


wq, wk, wv, wo, w1, w3, w2 = mk(E, E), mk(E, E), mk(E, E), mk(E, E), mk(FF, E), mk(FF, E), mk(E, FF)
x = torch.randn(M, E, device="cuda")
qkv = torch.zeros(3, M, E, device="cuda")
o = torch.zeros(M, E, device="cuda")
tmp = torch.zeros(2, M, FF, device="cuda")
out = torch.zeros(M, E, device="cuda")
torch.cuda.synchronize()
for _ in range(2):
    ns.mul_qkv(wq, wk, wv, x.data_ptr(), E, qkv.data_ptr(), E, M)
    ns.mul_mat(wo, x.data_ptr(), E, o.data_ptr(), E, M)
    ns.ffn_silu(w1, w2, w3, x.data_ptr(), E, tmp.data_ptr(), out.data_ptr(), E, M)
L.bestla_device_sync(None)
print("ok", float(out.abs().mean()))
