"""Generate tests/golden/*.npz from the REFERENCE's own code (oracle/_ref/*.so = /root/reference compiled in place).

Run in the authoring container (needs /root/reference):  python tests/golden/make_golden.py
The fixtures travel with the repo; the GPU box never sees /root/reference.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    assert oracle.ref_ggml() is not None and oracle.ref_btla() is not None, "needs oracle/_ref (build with /root/reference)"
    rng = np.random.default_rng(1234)  # seed borrowed from the reference CI (--seed 1234)

    # ---- ggml Q4_0 x Q8_0: ne_compute_forward_mul_mat_q_f32 on a small problem, M in {1, 5}
    N, K, M = 64, 512, 5
    w = rng.normal(0, 0.02, (N, K)).astype(np.float32)
    a = rng.normal(0, 1.0, (M, K)).astype(np.float32)
    a[1, :32] = 0.0
    wq = oracle.quantize_q4_0(w, "ref")
    aq = oracle.quantize_q8_0(a, "ref", "runtime")
    out = oracle.mul_mat_q4_0_f32(wq, a, "ref", nth=1)
    np.savez_compressed(os.path.join(HERE, "ggml_q4_0.npz"), w=w, a=a, wq=wq, aq=aq, out=out,
                        wdq=oracle.dequantize_q4_0(wq, K, "ref"))

    # ---- ggml Q6_K x Q8_K (output.weight of llama.cpp "Q4_0" GGUF files): quantisers, dequantiser, mul_mat
    rng6 = np.random.default_rng(4321)
    N6, K6, M6 = 40, 1024, 3
    w6 = rng6.normal(0, 0.02, (N6, K6)).astype(np.float32)
    w6[2, :256] = 0.0
    w6[3, 32:48] = 0.0
    a6 = rng6.normal(0, 1.0, (M6, K6)).astype(np.float32)
    a6[1, 256:512] = 0.0
    wq6 = oracle.quantize_q6_K(w6, "ref")
    np.savez_compressed(os.path.join(HERE, "ggml_q6_K.npz"), w=w6, a=a6, wq=wq6, aq=oracle.quantize_q8_K(a6, "ref"),
                        wdq=oracle.dequantize_q6_K(wq6, K6, "ref"), out=oracle.mul_mat_q6_K_f32(wq6, a6, "ref", nth=1))

    # ---- BesTLA: RTN quantiser, activation quantisers, NF4 (kernel_ref.h)
    K2, N2, M2 = 256, 48, 4
    w2 = rng.uniform(-0.5, 0.5, (K2, N2)).astype(np.float32)
    w2[:, 1] = np.abs(w2[:, 1])
    a2 = rng.uniform(-0.5, 0.5, (M2, K2)).astype(np.float32)
    d = dict(w=w2, a=a2)
    for g in (32, 128):
        for asym in (False, True):
            q, sc, zp = oracle.btla_quantize(w2, g, 4, asym, "ref")
            d[f"s4_g{g}_{'asym' if asym else 'sym'}_q"] = q
            d[f"s4_g{g}_{'asym' if asym else 'sym'}_sc"] = sc
            if asym:
                d[f"s4_g{g}_asym_zp"] = zp
        q8, sc8, _ = oracle.btla_quantize(w2, g, 8, False, "ref")
        d[f"s8_g{g}_q"], d[f"s8_g{g}_sc"] = q8, sc8
        qn, scn = oracle.btla_quantize_nf4(w2, g, "ref")
        d[f"nf4_g{g}_q"], d[f"nf4_g{g}_sc"] = qn, scn
        au, asu, azu = oracle.btla_quantize_act_u8(a2, g, "ref")
        d[f"act_u8_g{g}_q"], d[f"act_u8_g{g}_sc"], d[f"act_u8_g{g}_zp"] = au, asu, azu
        as8, ass = oracle.btla_quantize_act_s8(a2, g, "ref")
        d[f"act_s8_g{g}_q"], d[f"act_s8_g{g}_sc"] = as8, ass
    np.savez_compressed(os.path.join(HERE, "btla_quant.npz"), **d)

    # ---- element-wise ops of the Llama eval graph, from the reference's own engine (oracle/_ref/libref_ne.so = core/ne_layers.c)
    import ctypes as C
    ne = oracle.ref_ne()
    assert ne is not None
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rl = np.random.default_rng(777)
    g = {}
    hd, H, T, n_past = 64, 3, 2, 9
    x = rl.normal(0, 1, (T, H, hd)).astype(np.float32)
    y = x.copy()
    ne.ref_ne_rope(vp(y), hd, H, T, n_past, 10000.0, 1.0)
    g["rope_x"], g["rope_y"], g["rope_cfg"] = x, y, np.array([hd, H, T, n_past], np.int32)
    s = rl.normal(0, 3, (3, 77)).astype(np.float32)
    sm = s.copy()
    ne.ref_ne_soft_max(vp(sm), 77, 3)
    g["softmax_x"], g["softmax_y"] = s, sm
    xr = rl.normal(0, 2, (2, 512)).astype(np.float32)
    yr = np.zeros_like(xr)
    ne.ref_ne_rms_norm(vp(xr), vp(yr), 512, 2, 1e-5)
    g["rms_x"], g["rms_y"] = xr, yr
    Ha, hda, ln = 2, 128, 45
    q = rl.normal(0, 1, (Ha, hda)).astype(np.float32)
    kc = rl.normal(0, 1, (Ha, ln, hda)).astype(np.float16)
    vc = rl.normal(0, 1, (Ha, ln, hda)).astype(np.float16)
    out = np.zeros((Ha, hda), np.float32)
    ne.ref_ne_attn_1tok(vp(q), vp(kc), vp(np.ascontiguousarray(vc.transpose(0, 2, 1))), vp(out), hda, Ha, ln,
                        float(np.float32(1.0) / np.float32(np.sqrt(np.float32(hda)))))
    g["attn_q"], g["attn_k"], g["attn_v"], g["attn_out"] = q, kc, vc, out
    np.savez_compressed(os.path.join(HERE, "llama_ops.npz"), **g)

    # ---- a tiny Llama evaluated by the reference's engine (graph of models/llama/llama.cpp; oracle.RefNeLlama): logits of a
    # 4-token prompt and of three single-token steps
    rm = np.random.default_rng(4242)
    hp = dict(n_vocab=96, n_embd=128, n_head=2, n_head_kv=2, n_layer=2, n_ff=192, n_ctx=16, norm_eps=1e-5, rope_theta=10000.0,
              rope_scale=1.0)
    E, FF, V = hp["n_embd"], hp["n_ff"], hp["n_vocab"]
    qw = lambda n, k: oracle.quantize_q4_0(rm.normal(0, 1.0 / np.sqrt(k), (n, k)).astype(np.float32), "ref")
    mdl = dict(tok=rm.normal(0, 1, (V, E)).astype(np.float32), out_norm=rm.uniform(0.5, 1.5, E).astype(np.float32), output=qw(V, E))
    layers = []
    for il in range(hp["n_layer"]):
        L = dict(attn_norm=rm.uniform(0.5, 1.5, E).astype(np.float32), ffn_norm=rm.uniform(0.5, 1.5, E).astype(np.float32),
                 wq=qw(E, E), wk=qw(E, E), wv=qw(E, E), wo=qw(E, E), w1=qw(FF, E), w2=qw(E, FF), w3=qw(FF, E))
        layers.append(L)
        for k2, v2 in L.items():
            mdl[f"l{il}.{k2}"] = v2
    ref = oracle.RefNeLlama(hp, mdl["tok"], mdl["out_norm"], mdl["output"], layers)
    steps = [[1, 40, 7, 91], [13], [55], [2]]
    pos = 0
    for i, t in enumerate(steps):
        mdl[f"logits{i}"] = ref.eval(t, pos)
        mdl[f"tokens{i}"] = np.array(t, np.int32)
        pos += len(t)
    ref.close()
    mdl["hp"] = np.array([hp[k2] for k2 in ("n_vocab", "n_embd", "n_head", "n_head_kv", "n_layer", "n_ff", "n_ctx")], np.int32)
    np.savez_compressed(os.path.join(HERE, "llama_tiny.npz"), **mdl)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
