"""ncu target: the integer tensor-core batch kernel on one lm_head-sized weight, M = 8 and 32, ggml Q4_0 and int4 g128 asym"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import neural_speed_b200 as ns
L = ns.lib(); L.bestla_init()
cp = lambda t: C.c_void_p(t.data_ptr())
n, k = 32000, 4096
for kw in (dict(group=32, stype=ns.S_F16, comp=ns.COMP_Q8_0, asym=False), dict(group=128, stype=ns.S_BF16, comp=ns.COMP_INT8, asym=True)):
    w = ns.Weight.random(n, k, seed=3, **kw)
    for M in (8, 32):
        x = torch.randn(M, k, device="cuda"); y = torch.zeros(M, n, device="cuda")
        for _ in range(2):
            assert L.ns_mul_mat(w.h, cp(x), k, cp(y), n, M, None, None, 0, None, None) == 0, ns.last_error()
        L.bestla_device_sync(None)
