"""The one stated deviation of the tensor-core prompt attention and of the split-context decode attention from the reference's
soft_max: the reference rounds p = e / sum to fp16 BEFORE the V product (ne_compute_forward_soft_max_f32,
core/ne_layers.c:8887-8954, then mul_mat(V, P) with P converted to fp16, :6943-7083); the kernels accumulate sum e V with the exact
fp16 e and divide once at the end -- and, with several context ranges, merge per-range {max, sum e, sum e V} with exp(max_s - max)
weights.  This numpy model bounds what that costs."""
import numpy as np


def f16(x):
    return np.asarray(x, np.float32).astype(np.float16).astype(np.float32)


def reference_order(s, v):
    mx = s.max()
    e = f16(np.exp(f16(s - mx)))
    p = f16(e * np.float32(1.0 / e.sum(dtype=np.float32)))
    return (p[:, None] * v).sum(axis=0, dtype=np.float32)


def kernel_order(s, v, ranges=1):
    parts = []
    for idx in np.array_split(np.arange(len(s)), ranges):
        mx = s[idx].max()
        e = f16(np.exp(f16(s[idx] - mx)))
        parts.append((mx, e.sum(dtype=np.float32), (e[:, None] * v[idx]).sum(axis=0, dtype=np.float32)))
    gm = max(p[0] for p in parts)
    num = sum(np.exp(np.float32(p[0] - gm)) * p[2] for p in parts)
    den = sum(np.exp(np.float32(p[0] - gm)) * p[1] for p in parts)
    return (num / den).astype(np.float32)


def test_normalising_after_the_v_product_stays_within_1e3_of_the_reference_order():
    rng = np.random.default_rng(0)
    worst = 0.0
    for length, hd, ranges in ((40, 64, 1), (300, 128, 1), (300, 128, 2), (2100, 128, 9)):
        for _ in range(4):
            s = rng.normal(0, 2.0, length).astype(np.float32)
            v = f16(rng.normal(0, 1.0, (length, hd)))
            a, b = reference_order(s, v), kernel_order(s, v, ranges)
            worst = max(worst, float(np.abs(a - b).max() / np.abs(a).max()))
    assert worst <= 1e-3, worst
