"""Generate tests/golden/moe_q4_0.npz from the REFERENCE's own engine: ne_mul_mat_id on Q4_0 experts through ne_graph_compute
(oracle/_ref/libref_ne.so = /root/reference/neural_speed/core/ne_layers.c compiled in place).

Run in the authoring container (needs /root/reference):  python tests/golden/make_golden_moe.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    L = oracle.ref_ne()
    assert L is not None and oracle.ref_ggml() is not None, "needs oracle/_ref (build with /root/reference)"
    rng = np.random.default_rng(1234)
    n_as, n, k, n_tok, n_used = 4, 96, 256, 7, 2
    w = rng.normal(0, 0.02, (n_as, n, k)).astype(np.float32)
    rows = [oracle.quantize_q4_0(w[e], "ref") for e in range(n_as)]
    a = rng.normal(0, 1.0, (n_tok, k)).astype(np.float32)
    a[2, :32] = 0.0
    ids = np.stack([rng.permutation(n_as)[:n_used] for _ in range(n_tok)]).astype(np.int32)  # top-2 of 4, distinct per token
    outs = [oracle.ref_mul_mat_id(L, rows, oracle.NE_TYPE_Q4_0, n, k, ids, s, a, n_threads=1) for s in range(n_used)]
    np.savez_compressed(os.path.join(HERE, "moe_q4_0.npz"), rows=np.stack(rows), a=a, ids=ids, out=np.stack(outs))
    print("wrote moe_q4_0.npz", ids.tolist())


if __name__ == "__main__":
    main()
