"""GPTQ/AWQ ingest: the numpy converter against vectors produced by the reference's own
convert/common.py (tests/golden/make_golden_gptq.py), and the blob it emits against the HF dequantisation formula."""
import os

import numpy as np
import pytest

import neural_speed_b200 as ns
from neural_speed_b200 import convert

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "gptq_awq.npz"))
CASES = ["gptq4_asym", "gptq4_sym", "gptq4_desc", "gptq8_asym", "gptq8_sym", "awq4"]


def _case(name):
    bits, g, sym, desc = (int(v) for v in GOLD[f"{name}.cfg"])
    cfg = dict(quant_method=str(GOLD[f"{name}.method"]), bits=bits, group_size=g, sym=bool(sym), desc_act=bool(desc))
    return GOLD[f"{name}.qweight"], GOLD[f"{name}.scales"], GOLD[f"{name}.qzeros"], GOLD[f"{name}.g_idx"], cfg


@pytest.mark.parametrize("name", CASES)
def test_canonical_tensors_match_reference_converter(name):
    qw, sc, qz, gi, cfg = _case(name)
    c = convert.to_canonical(qw, sc, qz, gi, **cfg)
    assert np.array_equal(c["q"], GOLD[f"{name}.ref_q"])
    assert np.array_equal(c["scales"], GOLD[f"{name}.ref_scales"])
    if cfg["sym"]:
        assert c["zp"] is None
    else:
        assert np.array_equal(c["zp"], GOLD[f"{name}.ref_zp"])
    assert (c["g_idx"] is not None) == cfg["desc_act"]


def _hf_dequant(name):
    """W[k, n] from the checkpoint's own definition: scale[g_idx[k], n] * (w[k, n] - (zero[g_idx[k], n] + 1)) for GPTQ,
    scale * (w - zero) for AWQ (AutoGPTQ / AutoAWQ dequantisation)."""
    qw, sc, qz, gi, cfg = _case(name)
    bits = cfg["bits"]
    if cfg["quant_method"] == "awq":
        w, s, z = convert.unpack_awq(qw, sc, qz)
        zf = z.astype(np.float32)
        wf = w.astype(np.float32)
    else:
        per, mask = 32 // bits, (1 << bits) - 1
        u = qw.view(np.uint32)
        wf = np.stack([(u >> np.uint32(bits * i)) & mask for i in range(per)], 1).reshape(-1, u.shape[1]).astype(np.float32)
        uz = qz.view(np.uint32)
        zf = np.stack([(uz >> np.uint32(bits * i)) & mask for i in range(per)], 2).reshape(uz.shape[0], -1).astype(np.float32)
        zf = (zf + 1) % (1 << bits) if bits == 8 else zf + 1
        s = sc.astype(np.float32)
    if cfg["sym"]:
        zf = np.full_like(zf, 1 << (bits - 1))
    g = gi if cfg["desc_act"] else np.arange(wf.shape[0]) // cfg["group_size"]
    return s.astype(np.float32)[g] * (wf - zf[g])


@pytest.mark.parametrize("name", CASES)
def test_blob_dequantises_to_the_checkpoint_weights(name):
    qw, sc, qz, gi, cfg = _case(name)
    blob = convert.to_blob(qw, sc, qz, gi, compute_dtype="fp32", **cfg)
    k, n = GOLD[f"{name}.ref_q"].shape
    w = ns.unpack_blob(blob, n, k)                       # [K, N] in the blob's (regrouped) row order
    want = _hf_dequant(name)
    if cfg["desc_act"]:
        order = np.argsort(gi, kind="stable")            # blob row j holds original row order[j]
        want = want[order]
    assert np.array_equal(w, want.astype(np.float32))


def test_regroup_matches_the_reference_loop_on_random_group_maps():
    rng = np.random.default_rng(5)
    k, g = 192, 32
    w = rng.integers(-8, 8, (k, 7)).astype(np.int8)
    gi = rng.permutation(np.repeat(np.arange(k // g), g))
    out = convert.regroup_by_g_idx(w, gi, g)
    seen = {}
    want = w.copy()
    for i, grp in enumerate(gi):                         # the loop of common.py:671-682
        seen[grp] = seen.get(grp, -1) + 1
        want[grp * g + seen[grp]] = w[i]
    assert np.array_equal(out, want)


def test_permute_llama_is_the_rotary_pair_interleave():
    n_head, hd = 4, 8
    w = np.arange(n_head * hd * 3).reshape(n_head * hd, 3)
    p = convert.permute_llama(w, n_head)
    # within a head, HF rows [0..hd/2) and [hd/2..hd) become interleaved pairs
    for h in range(n_head):
        for i in range(hd // 2):
            assert np.array_equal(p[h * hd + 2 * i], w[h * hd + i])
            assert np.array_equal(p[h * hd + 2 * i + 1], w[h * hd + hd // 2 + i])
