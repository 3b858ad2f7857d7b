"""diagnostic: per-step logit error of the CUDA eval step against the reference engine at Llama-2-7B layer shapes"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import neural_speed_b200 as ns
import oracle
from oracle.llama_model import OracleLlama, greedy
ns.lib().bestla_init()
rng = np.random.default_rng(2024)
NL = int(sys.argv[1]) if len(sys.argv) > 1 else 2
hp = dict(n_vocab=32000, n_embd=4096, n_head=32, n_head_kv=32, n_layer=NL, n_ff=11008, n_ctx=64, norm_eps=1e-5, rope_theta=10000.0, rope_scale=1.0)
E, FF, V = hp["n_embd"], hp["n_ff"], hp["n_vocab"]
tok = rng.standard_normal((V, E), dtype=np.float32)
out_norm = rng.uniform(0.5, 1.5, E).astype(np.float32)
qw = lambda n, k: oracle.quantize_q4_0(rng.standard_normal((n, k), dtype=np.float32) * np.float32(1.0 / np.sqrt(k)))
shapes = dict(wq=(E, E), wk=(E, E), wv=(E, E), wo=(E, E), w1=(FF, E), w2=(E, FF), w3=(FF, E))
layers = []
for _ in range(NL):
    lay = dict(attn_norm=rng.uniform(0.5, 1.5, E).astype(np.float32), ffn_norm=rng.uniform(0.5, 1.5, E).astype(np.float32))
    for name, (n, k) in shapes.items():
        lay[name] = qw(n, k)
    layers.append(lay)
out_rows = qw(V, E)
ref = oracle.RefNeLlama(hp, tok, out_norm, out_rows, layers)
orc = OracleLlama(hp, tok, out_norm, out_rows, layers)
eng = ns.Llama(**hp)
eng.set_f32(ns.Llama.TOK_EMBD, 0, tok); eng.set_f32(ns.Llama.OUT_NORM, 0, out_norm)
eng.set_weight(ns.Llama.OUTPUT, 0, ns.Weight.from_q4_0_host(out_rows, V, E))
ids = dict(wq=ns.Llama.WQ, wk=ns.Llama.WK, wv=ns.Llama.WV, wo=ns.Llama.WO, w1=ns.Llama.W1, w2=ns.Llama.W2, w3=ns.Llama.W3)
for il, lay in enumerate(layers):
    eng.set_f32(ns.Llama.ATTN_NORM, il, lay["attn_norm"]); eng.set_f32(ns.Llama.FFN_NORM, il, lay["ffn_norm"])
    for name, (n, k) in shapes.items():
        eng.set_weight(ids[name], il, ns.Weight.from_q4_0_host(lay[name], n, k))
prompt = [1] + [int(t) for t in rng.integers(3, V, 11)]
t, pos = prompt[0], 0
for step in range(20):
    w_ref = ref.eval([t], pos); w_orc = orc.eval([t], pos) if step < 8 else w_ref
    got, nxt = eng.eval([t], pos)
    sc = max(1.0, float(np.abs(w_ref).max()))
    top = np.sort(w_ref)[-2:]
    print(f"step {step} pos {pos}: |gpu-ref| {np.abs(got-w_ref).max()/sc:.2e}  |gpu-orc| {np.abs(got-w_orc).max()/sc:.2e}  |orc-ref| {np.abs(w_orc-w_ref).max()/sc:.2e}  rms {np.sqrt(np.mean((got-w_ref)**2))/sc:.2e} margin {(top[1]-top[0])/sc:.2e} same_id {nxt==greedy(w_ref)}", flush=True)
    pos += 1
    t = prompt[pos] if pos < len(prompt) else greedy(w_ref)
