"""GPU parity of the persistent multi-op decode kernel (ns_program_*): a dependency chain of fused matmuls executed by one
cooperative launch must give exactly what the per-op kernels give (same arithmetic), and match the CPU oracle."""
import ctypes as C

import numpy as np
import pytest

import oracle
import neural_speed_b200 as ns

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    ns.lib().bestla_init()
    yield


def sync():
    torch.cuda.synchronize()
    ns.lib().bestla_device_sync(None)


def q4_weight(rng, n, k):
    rows = oracle.quantize_q4_0(rng.normal(0, 0.05, (n, k)).astype(np.float32))
    return rows, ns.Weight.from_q4_0_host(rows, n, k)


@pytest.mark.parametrize("m", [1, 2, 3])
def test_program_chain_matches_per_op_and_oracle(m):
    rng = np.random.default_rng(40 + m)
    N, F = 1024, 8192  # F/2 = 4096 pairs x ... enough units per CTA to wrap the ring several times at K = 1024
    (rq, wq), (rk, wk), (rv, wv) = q4_weight(rng, N, N), q4_weight(rng, N, N), q4_weight(rng, N, N)
    ro, wo = q4_weight(rng, N, N)
    (r1, w1), (r3, w3) = q4_weight(rng, F, N), q4_weight(rng, F, N)
    r2, w2 = q4_weight(rng, N, F)
    x = torch.from_numpy(rng.normal(0, 1, (m, N)).astype(np.float32)).cuda()
    qkv = torch.zeros(m, 3 * N, device="cuda")
    o = torch.zeros(m, N, device="cuda")
    h = torch.zeros(m, F, device="cuda")
    y = torch.zeros(m, N, device="cuda")
    prog = ns.Program(m)
    prog.add([wq, wk, wv], ns.Program.CONCAT, x.data_ptr(), N, qkv.data_ptr(), 3 * N, barrier_before=0)
    prog.add([wo], ns.Program.PLAIN, qkv.data_ptr(), 3 * N, o.data_ptr(), N)            # input = the q slice of qkv
    prog.add([w1, w3], ns.Program.GATE_UP_SILU, o.data_ptr(), N, h.data_ptr(), F)
    prog.add([w2], ns.Program.PLAIN, h.data_ptr(), F, y.data_ptr(), N, residual_ptr=x.data_ptr())
    prog.finalize()
    torch.cuda.synchronize()
    for _ in range(3):  # several launches: the epoch / barrier counters must carry over correctly
        prog.run()
    sync()
    got = {k: v.cpu().numpy() for k, v in dict(qkv=qkv, o=o, h=h, y=y).items()}

    # per-op kernels (same arithmetic -> bit-identical)
    qkv2 = torch.zeros(3, m, N, device="cuda")
    o2 = torch.zeros(m, N, device="cuda")
    h2 = torch.zeros(m, F, device="cuda")
    y2 = torch.zeros(m, N, device="cuda")
    torch.cuda.synchronize()
    ns.mul_qkv(wq, wk, wv, x.data_ptr(), N, qkv2.data_ptr(), N, m)
    sync()
    qin = qkv2[0].contiguous()
    ns.mul_mat(wo, qin.data_ptr(), N, o2.data_ptr(), N, m)
    sync()
    ns.ffn_silu(w1, w2, w3, o2.data_ptr(), N, h2.data_ptr(), y2.data_ptr(), N, m)
    sync()
    if m <= 2:  # from 3 rows on the per-op API runs the integer tensor-core kernel: same block sums, another fp32 summation order
        assert np.array_equal(got["qkv"].reshape(m, 3, N).transpose(1, 0, 2), qkv2.cpu().numpy())
        assert np.array_equal(got["o"], o2.cpu().numpy())
        assert np.array_equal(got["h"], h2.cpu().numpy())
        np.testing.assert_array_equal(got["y"], (y2 + x).cpu().numpy())
    else:
        np.testing.assert_allclose(got["qkv"].reshape(m, 3, N).transpose(1, 0, 2), qkv2.cpu().numpy(), rtol=0, atol=2e-5)

    # CPU oracle, stage by stage on the GPU's own intermediate inputs (isolates each op)
    xn = x.cpu().numpy()
    def close(a, b, rtol=1e-4):
        np.testing.assert_allclose(a, b, rtol=rtol, atol=rtol * (np.abs(b).max() + 1e-30))
    want_qkv = np.concatenate([oracle.mul_mat_q4_0_f32(r, xn) for r in (rq, rk, rv)], axis=1)
    close(got["qkv"], want_qkv)
    close(got["o"], oracle.mul_mat_q4_0_f32(ro, np.ascontiguousarray(got["qkv"][:, :N])))
    g = oracle.mul_mat_q4_0_f32(r1, got["o"])
    u = oracle.mul_mat_q4_0_f32(r3, got["o"])
    silu = np.vectorize(lambda t: oracle.lib().orc_silu(float(t)))(g).astype(np.float32)
    close(got["h"], silu * u, 2e-5 * 10)
    close(got["y"], oracle.mul_mat_q4_0_f32(r2, got["h"]) + xn)


def test_program_btla_int8_asym():
    rng = np.random.default_rng(77)
    m, N, K, g = 2, 512, 1024, 128
    w = rng.uniform(-0.5, 0.5, (K, N)).astype(np.float32)
    a = rng.uniform(-0.5, 0.5, (m, K)).astype(np.float32)
    q, sc, zp = oracle.btla_quantize(w, g, 4, True)
    wd = ns.Weight.from_unpacked(q, sc, zp, g, ns.W_S4, ns.S_BF16, ns.COMP_INT8)
    x = torch.from_numpy(a).cuda()
    y = torch.zeros(m, N, device="cuda")
    prog = ns.Program(m)
    prog.add([wd], ns.Program.PLAIN, x.data_ptr(), K, y.data_ptr(), N, barrier_before=0)
    prog.finalize()
    torch.cuda.synchronize()
    prog.run()
    sync()
    a8, asc, azp = oracle.btla_quantize_act_u8(a, g)
    sc_b = oracle.bf16_bits_to_f32(oracle.f32_to_bf16_bits(sc))
    want = oracle.btla_gemv_u8s8(a8, asc, azp, q, sc_b, zp, g)
    np.testing.assert_allclose(y.cpu().numpy(), want, rtol=1e-4, atol=1e-4 * np.abs(want).max())
