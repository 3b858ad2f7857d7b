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
