"""The oracle restatement must reproduce the committed golden fixtures (generated from the reference's own code by
tests/golden/make_golden.py) bit-for-bit -- this also runs on boxes where oracle/_ref is absent."""
import os

import numpy as np
import pytest

import oracle

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_ggml_q4_0_golden():
    z = np.load(os.path.join(G, "ggml_q4_0.npz"))
    assert np.array_equal(oracle.quantize_q4_0(z["w"]), z["wq"])
    assert np.array_equal(oracle.quantize_q8_0(z["a"]), z["aq"])
    assert np.array_equal(oracle.dequantize_q4_0(z["wq"], z["w"].shape[1]), z["wdq"])
    assert np.array_equal(oracle.mul_mat_q4_0_f32(z["wq"], z["a"]), z["out"])


def test_ggml_q6_K_golden():
    z = np.load(os.path.join(G, "ggml_q6_K.npz"))
    k = z["w"].shape[1]
    assert np.array_equal(oracle.quantize_q6_K(z["w"]), z["wq"])
    assert np.array_equal(oracle.quantize_q8_K(z["a"]), z["aq"])
    assert np.array_equal(oracle.dequantize_q6_K(z["wq"], k), z["wdq"])
    assert np.array_equal(oracle.mul_mat_q6_K_f32(z["wq"], z["a"]), z["out"])


@pytest.mark.parametrize("g", [32, 128])
def test_btla_quant_golden(g):
    z = np.load(os.path.join(G, "btla_quant.npz"))
    w, a = z["w"], z["a"]
    for asym in (False, True):
        tag = f"s4_g{g}_{'asym' if asym else 'sym'}"
        q, sc, zp = oracle.btla_quantize(w, g, 4, asym)
        assert np.array_equal(q, z[tag + "_q"]) and np.array_equal(sc, z[tag + "_sc"])
        if asym:
            assert np.array_equal(zp, z[tag + "_zp"])
    q, sc, _ = oracle.btla_quantize(w, g, 8, False)
    assert np.array_equal(q, z[f"s8_g{g}_q"]) and np.array_equal(sc, z[f"s8_g{g}_sc"])
    q, sc = oracle.btla_quantize_nf4(w, g)
    assert np.array_equal(q, z[f"nf4_g{g}_q"]) and np.array_equal(sc, z[f"nf4_g{g}_sc"])
    q, sc, zp = oracle.btla_quantize_act_u8(a, g)
    assert np.array_equal(q, z[f"act_u8_g{g}_q"]) and np.array_equal(sc, z[f"act_u8_g{g}_sc"]) and np.array_equal(zp, z[f"act_u8_g{g}_zp"])
    q, sc = oracle.btla_quantize_act_s8(a, g)
    assert np.array_equal(q, z[f"act_s8_g{g}_q"]) and np.array_equal(sc, z[f"act_s8_g{g}_sc"])


def test_llama_elementwise_ops_golden():
    """rope / soft_max / rms_norm / single-token attention of the Llama eval graph: the numpy restatement in
    oracle/llama_model.py against outputs of the reference's own engine (tests/golden/llama_ops.npz)"""
    from oracle import llama_model as lm
    z = np.load(os.path.join(G, "llama_ops.npz"))
    hd, H, T, n_past = (int(v) for v in z["rope_cfg"])
    got = np.stack([lm.rope_mode0(z["rope_x"][t], n_past + t, hd) for t in range(T)])
    assert np.array_equal(got, z["rope_y"])
    assert np.array_equal(np.stack([lm.soft_max_f16table(r) for r in z["softmax_x"]]), z["softmax_y"])
    assert np.array_equal(lm.rms_norm(z["rms_x"], 1e-5), z["rms_y"])
    q, kc, vc = z["attn_q"], z["attn_k"], z["attn_v"]
    scale = np.float32(1.0) / np.float32(np.sqrt(np.float32(q.shape[1])))
    for h in range(q.shape[0]):
        s = lm.vec_dot_f16_rows(kc[h].astype(np.float32), lm._f16(q[h])) * scale
        p = lm.soft_max_f16table(s)
        assert np.array_equal(lm.vec_dot_f16_rows(np.ascontiguousarray(vc[h].astype(np.float32).T), lm._f16(p)), z["attn_out"][h])


def test_llama_tiny_model_golden():
    """whole eval graph: logits produced by the reference's engine for a tiny Q4_0 Llama (prompt + single-token steps)"""
    from oracle.llama_model import OracleLlama
    z = np.load(os.path.join(G, "llama_tiny.npz"))
    keys = ("n_vocab", "n_embd", "n_head", "n_head_kv", "n_layer", "n_ff", "n_ctx")
    hp = dict(zip(keys, (int(v) for v in z["hp"])), norm_eps=1e-5, rope_theta=10000.0, rope_scale=1.0)
    layers = [{k: z[f"l{il}.{k}"] for k in ("attn_norm", "ffn_norm", "wq", "wk", "wv", "wo", "w1", "w2", "w3")}
              for il in range(hp["n_layer"])]
    orc = OracleLlama(hp, z["tok"], z["out_norm"], z["output"], layers)
    pos = 0
    for i in range(4):
        t = list(z[f"tokens{i}"])
        assert np.array_equal(orc.eval(t, pos), z[f"logits{i}"]), i
        pos += len(t)
