"""Build a Llama-2-7B-shaped engine with a few layers and run a short greedy generation -- meant to be run under
`ncu --metrics gpu__time_duration.sum` to get the per-kernel launch list of the decode step (profiles/summarize.py)."""
import sys

import numpy as np
import torch

import neural_speed_b200 as ns
import ctypes as C

n_layer = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n_new = int(sys.argv[2]) if len(sys.argv) > 2 else 4
n_prompt = int(sys.argv[3]) if len(sys.argv) > 3 else 32
E, FF, V = 4096, 11008, 32000
L = ns.lib()
L.bestla_init()


def mk(n, k):
    w = torch.randn(n, k, device="cuda") * 0.02
    rows = torch.empty(n * (k // 32) * 18, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    assert L.ns_device_quantize_q4_0(C.c_void_p(w.data_ptr()), C.c_void_p(rows.data_ptr()), n, k, None) == 0
    L.bestla_device_sync(None)
    return ns.Weight.from_q4_0_device(rows.data_ptr(), n, k, (k // 32) * 18)


eng = ns.Llama(V, E, 32, 32, n_layer, FF, max(1024, n_prompt + 64), 1e-5)
eng.set_f32(ns.Llama.TOK_EMBD, 0, torch.randn(V, E).numpy())
ones = np.ones(E, np.float32)
eng.set_f32(ns.Llama.OUT_NORM, 0, ones)
eng.set_weight(ns.Llama.OUTPUT, 0, mk(V, E))
ids = dict(wq=ns.Llama.WQ, wk=ns.Llama.WK, wv=ns.Llama.WV, wo=ns.Llama.WO, w1=ns.Llama.W1, w2=ns.Llama.W2, w3=ns.Llama.W3)
shapes = dict(wq=(E, E), wk=(E, E), wv=(E, E), wo=(E, E), w1=(FF, E), w2=(E, FF), w3=(FF, E))
for il in range(n_layer):
    eng.set_f32(ns.Llama.ATTN_NORM, il, ones)
    eng.set_f32(ns.Llama.FFN_NORM, il, ones)
    for name, (n, k) in shapes.items():
        eng.set_weight(ids[name], il, mk(n, k))
prompt = (np.arange(n_prompt, dtype=np.int32) % 30000) + 1
_, nxt = eng.eval(prompt, 0, want_logits=False)
print("generated", eng.generate(int(nxt), n_prompt, n_new))
