"""diagnostic: one decoder layer at 7B shapes, position 0, stage by stage: CUDA ops (C-ABI) against the numpy oracle"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import neural_speed_b200 as ns
import oracle
from oracle.llama_model import rms_norm
L = ns.lib(); L.bestla_init()
rng = np.random.default_rng(5)
E, FF = 4096, 11008
qw = lambda n, k: oracle.quantize_q4_0(rng.standard_normal((n, k), dtype=np.float32) * np.float32(1.0 / np.sqrt(k)))
rows = {nm: qw(n, k) for nm, (n, k) in dict(wq=(E, E), wk=(E, E), wv=(E, E), wo=(E, E), w1=(FF, E), w2=(E, FF), w3=(FF, E)).items()}
W = {nm: ns.Weight.from_q4_0_host(r, r.shape[0], (E if nm != "w2" else FF)) for nm, r in rows.items()}
x = rng.standard_normal((1, E), dtype=np.float32)
wn = rng.uniform(0.5, 1.5, E).astype(np.float32)
def rel(a, b): return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))
mm = lambda r, a: oracle.mul_mat_q4_0_f32(r, np.ascontiguousarray(a, np.float32))
def gmm(w, a, n):
    at = torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda(); out = torch.zeros(1, n, device="cuda")
    assert L.ns_mul_mat(w.h, C.c_void_p(at.data_ptr()), a.shape[1], C.c_void_p(out.data_ptr()), n, 1, None, None, 0, None, None) == 0
    L.bestla_device_sync(None); return out.cpu().numpy()
xn = rms_norm(x, 1e-5) * wn
for nm in ("wq", "wk", "wv"):
    print(nm, rel(gmm(W[nm], xn, E), mm(rows[nm], xn)))
# fused qkv
at = torch.from_numpy(xn).cuda(); qkv = torch.zeros(3, 1, E, device="cuda")
assert L.ns_mul_qkv(W["wq"].h, W["wk"].h, W["wv"].h, C.c_void_p(at.data_ptr()), E, C.c_void_p(qkv.data_ptr()), E, 1, None, None) == 0
L.bestla_device_sync(None)
q3 = qkv.cpu().numpy()
for i, nm in enumerate(("wq", "wk", "wv")):
    print("fused", nm, rel(q3[i], mm(rows[nm], xn)))
v = mm(rows["wv"], xn)
attn = v.astype(np.float16).astype(np.float32)
o_or = mm(rows["wo"], attn) + x
# wo with residual on GPU
at = torch.from_numpy(attn).cuda(); rs = torch.from_numpy(x).cuda(); out = torch.zeros(1, E, device="cuda")
assert L.ns_mul_mat(W["wo"].h, C.c_void_p(at.data_ptr()), E, C.c_void_p(out.data_ptr()), E, 1, None, C.c_void_p(rs.data_ptr()), 0, None, None) == 0
L.bestla_device_sync(None)
print("wo+res", rel(out.cpu().numpy(), o_or))
h = rms_norm(o_or, 1e-5) * wn
g = mm(rows["w1"], h); u = mm(rows["w3"], h)
silu = np.array([[oracle.lib().orc_silu(float(z)) for z in row] for row in g], np.float32)
mid = silu * u
dn = mm(rows["w2"], mid)
at = torch.from_numpy(h).cuda(); tmp = torch.zeros(2, 1, FF, device="cuda"); out = torch.zeros(1, E, device="cuda")
assert L.ns_ffn_silu(W["w1"].h, W["w2"].h, W["w3"].h, C.c_void_p(at.data_ptr()), E, C.c_void_p(tmp.data_ptr()), C.c_void_p(out.data_ptr()), E, 1, None, None) == 0
L.bestla_device_sync(None)
print("ffn mid", rel(tmp[0].cpu().numpy(), mid), " ffn out", rel(out.cpu().numpy(), dn))
print("gate alone", rel(gmm(W["w1"], h, FF), g), " up alone", rel(gmm(W["w3"], h, FF), u), " down alone (oracle mid)", rel(gmm(W["w2"], mid, E), dn))
