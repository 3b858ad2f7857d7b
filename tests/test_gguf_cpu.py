"""GGUF (llama) reader for the eval step: write a tiny model with the gguf package's own writer, read it back."""
import numpy as np
import pytest

import oracle
from neural_speed_b200 import gguf_loader

gguf = pytest.importorskip("gguf")


def _write(path, tie_embeddings=False, out_q6k=True, embd_type="q4_0"):
    rng = np.random.default_rng(2)
    V, E, H, HK, NL, FF = 64, 256, 4, 2, 2, 512
    kvd = E // H * HK
    w = gguf.GGUFWriter(path, "llama")
    w.add_context_length(128)
    w.add_embedding_length(E)
    w.add_block_count(NL)
    w.add_feed_forward_length(FF)
    w.add_head_count(H)
    w.add_head_count_kv(HK)
    w.add_layer_norm_rms_eps(1e-5)
    w.add_rope_freq_base(10000.0)
    T = gguf.GGMLQuantizationType
    ref = {}

    def q4(name, n, k):
        rows = oracle.quantize_q4_0(rng.normal(0, 0.05, (n, k)).astype(np.float32))
        ref[name] = rows
        w.add_tensor(name, rows, raw_dtype=T.Q4_0)

    tok_f = rng.normal(0, 1, (V, E)).astype(np.float32)
    if embd_type == "q4_0":
        rows = oracle.quantize_q4_0(tok_f)
        ref["token_embd.weight"] = oracle.dequantize_q4_0(rows, E)
        w.add_tensor("token_embd.weight", rows, raw_dtype=T.Q4_0)
    else:
        ref["token_embd.weight"] = tok_f.astype(np.float16).astype(np.float32)
        w.add_tensor("token_embd.weight", tok_f.astype(np.float16))
    ref["output_norm.weight"] = rng.uniform(0.5, 1.5, E).astype(np.float32)
    w.add_tensor("output_norm.weight", ref["output_norm.weight"])
    if not tie_embeddings:
        if out_q6k:
            rows = oracle.quantize_q6_K(rng.normal(0, 0.05, (V, E)).astype(np.float32))
            ref["output.weight"] = rows
            w.add_tensor("output.weight", rows, raw_dtype=T.Q6_K)
        else:
            q4("output.weight", V, E)
    for il in range(NL):
        for nm in ("attn_norm", "ffn_norm"):
            ref[f"blk.{il}.{nm}.weight"] = rng.uniform(0.5, 1.5, E).astype(np.float32)
            w.add_tensor(f"blk.{il}.{nm}.weight", ref[f"blk.{il}.{nm}.weight"])
        for nm, (n, k) in dict(attn_q=(E, E), attn_k=(kvd, E), attn_v=(kvd, E), attn_output=(E, E), ffn_gate=(FF, E),
                               ffn_down=(E, FF), ffn_up=(FF, E)).items():
            q4(f"blk.{il}.{nm}.weight", n, k)
    w.write_header_to_file()
    w.write_kv_data_to_file()
    w.write_tensors_to_file()
    w.close()
    return ref, dict(n_vocab=V, n_embd=E, n_head=H, n_head_kv=HK, n_layer=NL, n_ff=FF, n_ctx=128)


def test_parse_llama_gguf_q4_0_with_q6_K_output(tmp_path):
    path = str(tmp_path / "tiny.gguf")
    ref, hp = _write(path)
    m = gguf_loader.parse(path)
    for k, v in hp.items():
        assert m.hparams[k] == v, k
    assert abs(m.hparams["norm_eps"] - 1e-5) < 1e-12 and m.hparams["rope_theta"] == 10000.0
    assert np.array_equal(m.tok_embd, ref["token_embd.weight"])           # dequantised exactly as the CPU path does
    assert np.array_equal(m.out_norm, ref["output_norm.weight"])
    assert m.output[0] == "q6_K" and np.array_equal(m.output[1], ref["output.weight"])
    names = dict(wq="attn_q", wk="attn_k", wv="attn_v", wo="attn_output", w1="ffn_gate", w2="ffn_down", w3="ffn_up")
    for il, L in enumerate(m.layers):
        assert np.array_equal(L["attn_norm"], ref[f"blk.{il}.attn_norm.weight"])
        assert np.array_equal(L["ffn_norm"], ref[f"blk.{il}.ffn_norm.weight"])
        for ours, g in names.items():
            assert L[ours][0] == "q4_0" and np.array_equal(L[ours][1], ref[f"blk.{il}.{g}.weight"])
    # the parsed model runs through the CPU graph oracle (GQA: 4 heads over 2 KV heads, Q6_K head)
    from oracle.llama_model import OracleLlama
    orc = OracleLlama(m.hparams, m.tok_embd, m.out_norm, m.output[1],
                      [{k: (v[1] if isinstance(v, tuple) else v) for k, v in L.items()} for L in m.layers], fmt="q6_K")
    logits = orc.eval([1, 5, 9], 0)
    assert logits.shape == (hp["n_vocab"],) and np.isfinite(logits).all()


def test_parse_f16_embeddings_and_tied_output(tmp_path):
    path = str(tmp_path / "tied.gguf")
    ref, hp = _write(path, tie_embeddings=True, embd_type="q4_0")
    m = gguf_loader.parse(path)
    assert m.output[0] == "q4_0" and m.output[1].shape == (hp["n_vocab"], hp["n_embd"] // 32 * 18)   # output = token_embd rows
    path2 = str(tmp_path / "f16.gguf")
    ref2, _ = _write(path2, out_q6k=False, embd_type="f16")
    m2 = gguf_loader.parse(path2)
    assert np.array_equal(m2.tok_embd, ref2["token_embd.weight"]) and m2.output[0] == "q4_0"


def test_other_architectures_are_refused(tmp_path):
    path = str(tmp_path / "x.gguf")
    w = gguf.GGUFWriter(path, "gptj")
    w.add_block_count(1)
    w.add_tensor("token_embd.weight", np.zeros((4, 32), np.float32))
    w.write_header_to_file()
    w.write_kv_data_to_file()
    w.write_tensors_to_file()
    w.close()
    with pytest.raises(ValueError, match="llama"):
        gguf_loader.parse(path)
