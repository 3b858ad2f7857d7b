"""Golden vectors for the GPTQ/AWQ ingest path, produced by IMPORTING the reference converter
(/root/reference/neural_speed/convert/common.py: unpack_weight + the desc_act regrouping and -8 recentring of
convert_q4_bestla_tensor :649-714).  Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden_gptq.py   ->  tests/golden/gptq_awq.npz
"""
import importlib.util
import os

import numpy as np
import torch

spec = importlib.util.spec_from_file_location("ref_common", "/root/reference/neural_speed/convert/common.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)


def pack_rows(vals, bits):      # [K, N] -> int32 [K*bits/32, N], LSB-first along K (AutoGPTQ qweight)
    per = 32 // bits
    v = vals.astype(np.uint32).reshape(-1, per, vals.shape[1])
    out = np.zeros((v.shape[0], v.shape[2]), np.uint32)
    for i in range(per):
        out |= v[:, i, :] << np.uint32(bits * i)
    return out.view(np.int32)


def pack_cols(vals, bits, order=None):   # [R, N] -> int32 [R, N*bits/32] along N (qzeros; AWQ qweight with its nibble order)
    per = 32 // bits
    v = vals.astype(np.uint32).reshape(vals.shape[0], -1, per)
    out = np.zeros(v.shape[:2], np.uint32)
    for i in range(per):
        out |= v[:, :, i] << np.uint32(bits * (order[i] if order else i))
    return out.view(np.int32)


def reference_canonical(qweight, scales, qzeros, g_idx, q_config):
    """The tensor-level body of convert_q4_bestla_tensor (common.py:656-701), minus file I/O."""
    int_weight, gptq_scales, gptq_zeros = ref.unpack_weight(torch.from_numpy(qweight), torch.from_numpy(scales),
                                                            torch.from_numpy(qzeros), q_config)
    int_weight = int_weight.view(-1, int_weight.shape[-1])
    if q_config.get("desc_act"):
        # the act-order step of convert_q4_bestla_tensor (common.py:667-683), driven row by row exactly as the reference does:
        # row i goes to the next free slot of its group g_idx[i]
        group_size = q_config["group_size"]
        filled = {}
        regrouped = int_weight.clone()
        for i, grp in enumerate(g_idx.tolist()):
            slot = filled.get(grp, 0)
            regrouped[grp * group_size + slot] = int_weight[i]
            filled[grp] = slot + 1
        int_weight = regrouped
    if q_config["bits"] == 4:
        int_weight = int_weight - 8
        gptq_zeros = gptq_zeros - 8
    return (np.ascontiguousarray(int_weight.numpy()).astype(np.int8), gptq_scales.float().numpy(),
            np.ascontiguousarray(gptq_zeros.numpy()).astype(np.int8))


def main():
    rng = np.random.default_rng(20240923)
    out = {}
    cases = [("gptq4_asym", "gptq", 4, 128, False, False), ("gptq4_sym", "gptq", 4, 32, True, False),
             ("gptq4_desc", "gptq", 4, 64, False, True), ("gptq8_asym", "gptq", 8, 64, False, False),
             ("gptq8_sym", "gptq", 8, 64, True, False), ("awq4", "awq", 4, 128, False, False)]
    for name, method, bits, g, sym, desc in cases:
        K, N = 256, 64
        hi = 1 << bits
        w = rng.integers(0, hi, (K, N))
        z = rng.integers(0, hi - 1, (K // g, N)) if not sym else np.full((K // g, N), hi // 2 - 1)
        scales = rng.uniform(0.005, 0.02, (K // g, N)).astype(np.float16)
        g_idx = np.repeat(np.arange(K // g), g).astype(np.int32)
        if desc:
            g_idx = rng.permutation(g_idx).astype(np.int32)
        if method == "awq":
            qweight, qzeros = pack_cols(w, 4, ref_order()), pack_cols(z, 4, ref_order())
        else:
            qweight, qzeros = pack_rows(w, bits), pack_cols(z, bits)
        q_config = dict(quant_method=method, bits=bits, group_size=g, sym=sym, desc_act=desc)
        qi, sc, zp = reference_canonical(qweight, scales, qzeros, g_idx, q_config)
        out[f"{name}.qweight"], out[f"{name}.qzeros"], out[f"{name}.scales"], out[f"{name}.g_idx"] = qweight, qzeros, scales, g_idx
        out[f"{name}.ref_q"], out[f"{name}.ref_scales"], out[f"{name}.ref_zp"] = qi, sc, zp
        out[f"{name}.cfg"] = np.array([bits, g, int(sym), int(desc)], np.int32)
        out[f"{name}.method"] = np.array(method)
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "gptq_awq.npz"), **out)
    print("wrote", len(cases), "cases")


def ref_order():
    return [0, 4, 1, 5, 2, 6, 3, 7]


if __name__ == "__main__":
    main()
