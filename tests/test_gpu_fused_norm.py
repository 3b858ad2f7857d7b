"""GPU tests of the RMSNorm folded into the decode GEMV launch (ns_rmsnorm_mul_mat / _mul_qkv / _ffn_silu).

Reference graph: cur = ne_rms_norm(x); cur = ne_mul(cur, norm_w); ne_mul_mat / ne_mul_qkv / ne_ffn_silu(W, cur)
(models/llama/llama.cpp:205-215, :601-612, :703-712).  The fused launch must equal (a) the same nodes issued one by one on the
device (separate norm, then the plain entry) and (b) the CPU oracle on the numpy-normalised row.  Only the fp32 order of the
sum of squares differs between the two device forms: a last-bit difference in 1/rms can move an activation code across a
rounding boundary (1/127 of a block maximum on one of k terms), hence the 2e-3 bar rather than bit equality.
"""
import ctypes as C

import numpy as np
import pytest

import oracle
import neural_speed_b200 as ns

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

EPS = 1e-6


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    ns.lib().bestla_init()
    yield


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def sync():
    torch.cuda.synchronize()
    ns.lib().bestla_device_sync(None)


def close(got, want, rtol):
    scale = float(np.abs(want).max()) + 1e-30
    np.testing.assert_allclose(got, want, rtol=rtol, atol=rtol * scale)


def rmsnorm_np(x, w, eps=EPS):
    x = x.astype(np.float32)
    ss = (x.astype(np.float64) ** 2).sum(axis=1, keepdims=True) / x.shape[1]
    inv = (1.0 / np.sqrt(ss + eps)).astype(np.float32)
    return (x * inv * w[None, :]).astype(np.float32)


@pytest.mark.parametrize("m,k", [(1, 4096), (2, 4096), (1, 1024), (1, 8192), (2, 11008), (1, 4000)])
def test_q4_0_rmsnorm_mul_mat_matches_separate_nodes_and_oracle(m, k):
    rng = np.random.default_rng(100 + m + k)
    n = 384
    if k % 32:
        pytest.skip("ggml Q4_0 rows are whole 32-blocks")
    w = rng.normal(0, 0.02, (n, k)).astype(np.float32)
    x = (rng.normal(0, 1, (m, k)) * rng.uniform(0.2, 5.0, (m, 1))).astype(np.float32)
    nw = rng.uniform(0.5, 1.5, k).astype(np.float32)
    res = rng.normal(0, 1, (m, n)).astype(np.float32)
    rows = oracle.quantize_q4_0(w)
    wd = ns.Weight.from_q4_0_host(rows, n, k)
    assert ns.rmsnorm_fusable([wd], m)
    xd, nwd, rd = dev(x), dev(nw), dev(res)
    out = torch.full((m, n), float("nan"), device="cuda")
    torch.cuda.synchronize()
    ns.rmsnorm_mul_mat(wd, xd.data_ptr(), k, nwd.data_ptr(), EPS, out.data_ptr(), n, m, rd.data_ptr())
    sync()
    got = out.cpu().numpy()
    # (a) separate nodes on the device
    xn = xd * torch.rsqrt((xd * xd).mean(dim=1, keepdim=True) + EPS) * nwd
    out2 = torch.full((m, n), float("nan"), device="cuda")
    torch.cuda.synchronize()
    ns.mul_mat(wd, xn.data_ptr(), k, out2.data_ptr(), n, m, None, rd.data_ptr())
    sync()
    close(got, out2.cpu().numpy(), 2e-3)
    # (b) CPU oracle
    want = oracle.mul_mat_q4_0_f32(rows, rmsnorm_np(x, nw)) + res
    close(got, want, 2e-3)


def test_three_rows_are_not_fusable_and_fail_loudly():
    rng = np.random.default_rng(5)
    n, k = 128, 1024
    wd = ns.Weight.from_q4_0_host(oracle.quantize_q4_0(rng.normal(0, 0.02, (n, k)).astype(np.float32)), n, k)
    assert not ns.rmsnorm_fusable([wd], 3)
    x, nw = dev(rng.normal(0, 1, (3, k)).astype(np.float32)), dev(np.ones(k, np.float32))
    out = torch.zeros(3, n, device="cuda")
    rc = ns.lib().ns_rmsnorm_mul_mat(wd.h, C.c_void_p(x.data_ptr()), k, C.c_void_p(nw.data_ptr()), EPS, C.c_void_p(out.data_ptr()), n, 3,
                                     None, None, None)
    assert rc != 0 and "RMSNorm" in ns.last_error()


@pytest.mark.parametrize("alg,comp", [("sym", "int8"), ("asym", "int8")])
def test_btla_int4_g128_rmsnorm_qkv_and_ffn(alg, comp):
    """BesTLA int4 g128 blobs with u8 activations (config 2/3 decode): fused QKV and fused FFN with the norm inside."""
    rng = np.random.default_rng(77)
    m, k, n, fmid, g = 1, 1024, 512, 1408, 128
    x = rng.normal(0, 1.5, (m, k)).astype(np.float32)
    nw = rng.uniform(0.5, 1.5, k).astype(np.float32)
    mk = lambda r, c: ns.Weight.from_blob(ns.np_bestla_quantize(rng.uniform(-0.5, 0.5, (r, c)).astype(np.float32), "int4", g, alg, "fp32", comp))
    wq, wk, wv = mk(n, k), mk(n, k), mk(n, k)
    w1, w3, w2 = mk(fmid, k), mk(fmid, k), mk(k, fmid)
    assert ns.rmsnorm_fusable([wq, wk, wv], m) and ns.rmsnorm_fusable([w1, w3], m)
    xd, nwd = dev(x), dev(nw)
    xn = xd * torch.rsqrt((xd * xd).mean(dim=1, keepdim=True) + EPS) * nwd
    qkv = torch.full((3, m, n), float("nan"), device="cuda")
    qkv2 = torch.full((3, m, n), float("nan"), device="cuda")
    torch.cuda.synchronize()
    ns.rmsnorm_mul_qkv(wq, wk, wv, xd.data_ptr(), k, nwd.data_ptr(), EPS, qkv.data_ptr(), n, m)
    ns.mul_qkv(wq, wk, wv, xn.data_ptr(), k, qkv2.data_ptr(), n, m)
    sync()
    close(qkv.cpu().numpy(), qkv2.cpu().numpy(), 2e-3)
    # FFN with the residual folded into the down projection: dst = x + W2 (silu(W1 xn) * W3 xn)
    tmp = torch.zeros(2 * m * fmid, device="cuda")
    tmp2 = torch.zeros(2 * m * fmid, device="cuda")
    ffn = torch.full((m, k), float("nan"), device="cuda")
    ffn2 = torch.full((m, k), float("nan"), device="cuda")
    torch.cuda.synchronize()
    ns.rmsnorm_ffn_silu(w1, w2, w3, xd.data_ptr(), k, nwd.data_ptr(), EPS, tmp.data_ptr(), ffn.data_ptr(), k, m, xd.data_ptr())
    ns.ffn_silu(w1, w2, w3, xn.data_ptr(), k, tmp2.data_ptr(), ffn2.data_ptr(), k, m)
    sync()
    close(ffn.cpu().numpy(), (ffn2 + xd).cpu().numpy(), 2e-3)
