"""Tensor-parallel host logic on CPU: blob / Q4_0 shard splitting and the world_size-2 all-reduce wiring over
`gloo`.  The per-rank matmuls are computed with the CPU oracle here (this is the test's checker; on GPUs the same shards go
through libns_b200 -- tests/test_gpu_tp.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import neural_speed_b200 as ns
import oracle
from neural_speed_b200 import tp


def _blob(seed, n, k, g=128, alg="sym", sdt="fp32", cdt="int8"):
    w = np.random.default_rng(seed).uniform(-0.5, 0.5, (n, k)).astype(np.float32)
    return ns.np_bestla_quantize(w, "int4", g, alg, sdt, cdt)


@pytest.mark.parametrize("alg,sdt,cdt,g", [("sym", "fp32", "int8", 128), ("asym", "bf16", "int8", 32), ("sym", "fp32", "fp32", 64)])
@pytest.mark.parametrize("split", [tp.SPLIT_N, tp.SPLIT_K])
def test_split_blob_is_unpack_slice_requantise(alg, sdt, cdt, g, split):
    """bestla_split_weight (model_files.h:1538-1562): the shard is the RTN re-quantisation of the dequantised slice with the
    source blob's attributes."""
    n, k, world = 192, 512, 2
    blob = _blob(3, n, k, g, alg, sdt, cdt)
    full = ns.unpack_blob(blob, n, k)  # [K, N]
    for rank in range(world):
        shard = tp.split_blob(blob, n, k, world, rank, split)
        sl = full[:, rank * n // world:(rank + 1) * n // world] if split == tp.SPLIT_N else full[rank * k // world:(rank + 1) * k // world]
        want = ns.np_bestla_quantize(np.ascontiguousarray(sl.T), "int4", g, alg, sdt, cdt)
        assert shard.size == want.size and np.array_equal(shard, want)
    with pytest.raises(ValueError):
        tp.split_blob(blob, n + 48, k, world, 0, split)  # header mismatch is refused, not sliced blindly


def test_split_blob_qkv_fusion_takes_a_third_of_each_projection():
    n, k, world = 3 * 96, 256, 2
    blob = _blob(5, n, k)
    full = ns.unpack_blob(blob, n, k)
    for rank in range(world):
        shard = tp.split_blob(blob, n, k, world, rank, tp.SPLIT_N, qkv_fusion=True)
        per = n // 3 // world
        sl = np.concatenate([full[:, j * n // 3 + rank * per: j * n // 3 + (rank + 1) * per] for j in range(3)], axis=1)
        want = ns.np_bestla_quantize(np.ascontiguousarray(sl.T), "int4", 128, "sym", "fp32", "int8")
        assert np.array_equal(shard, want)


def test_shard_plan_llama2_70b_and_constraints():
    for w in (2, 4, 8):
        p = tp.LlamaShardPlan(w, 8192, 28672, 64, 8, 128)
        sh = p.shapes()
        assert sh["wq"] == (tp.SPLIT_N, 8192 // w, 8192) and sh["wk"] == (tp.SPLIT_N, 1024 // w, 8192)
        assert sh["wo"] == (tp.SPLIT_K, 8192, 8192 // w) and sh["w2"] == (tp.SPLIT_K, 8192, 28672 // w)
        assert sum(n * k for _, n, k in sh.values()) * w == 8192 * 8192 * 2 + 2 * 1024 * 8192 + 3 * 8192 * 28672
    with pytest.raises(ValueError):
        tp.LlamaShardPlan(16, 8192, 28672, 64, 8, 128)       # n_head_kv % W != 0 (llama.cpp:121-124)
    with pytest.raises(ValueError):
        tp.LlamaShardPlan(8, 4096, 11008, 32, 32, 128)       # 11008/8 = 1376 is not a multiple of the group


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    try:
        ctx = tp.TPContext(backend="gloo")
        assert (ctx.rank, ctx.world) == (rank, world)
        n_embd, n_ff, n_head, g, m = 256, 512, 4, 128, 2
        plan = tp.LlamaShardPlan(world, n_embd, n_ff, n_head, n_head, g)
        rng = np.random.default_rng(11)                    # identical on every rank
        x = rng.uniform(-0.5, 0.5, (m, n_embd)).astype(np.float32)
        full = {name: _blob(20 + i, *{"wq": (n_embd, n_embd), "wo": (n_embd, n_embd), "w1": (n_ff, n_embd), "w3": (n_ff, n_embd),
                                     "w2": (n_embd, n_ff)}[name]) for i, name in enumerate(("wq", "wo", "w1", "w3", "w2"))}
        dims = {"wq": (n_embd, n_embd), "wo": (n_embd, n_embd), "w1": (n_ff, n_embd), "w3": (n_ff, n_embd), "w2": (n_embd, n_ff)}
        sh = plan.shapes()

        def shard_w(name, r):                               # dequantised [k_local, n_local] of rank r's shard
            nn, kk = dims[name]
            b = tp.split_blob(full[name], nn, kk, world, r, sh[name][0])
            return ns.unpack_blob(b, sh[name][1], sh[name][2])

        # N-split (q, gate, up): local columns, no communication
        q_loc = oracle.gemm_f64acc(x, shard_w("wq", rank))
        # K-split (o): partial over this rank's slice of the attention output, then sum all-reduce (llama.cpp:592)
        attn_all = [oracle.gemm_f64acc(x, shard_w("wq", r)) for r in range(world)]  # identity attention core
        part = torch.from_numpy(oracle.gemm_f64acc(q_loc, shard_w("wo", rank)))
        ctx.all_reduce(part)
        want_o = sum(oracle.gemm_f64acc(attn_all[r], shard_w("wo", r)).astype(np.float64) for r in range(world))
        assert np.allclose(part.numpy(), want_o, rtol=1e-5, atol=1e-5)
        # FFN: silu(x W1_r) * (x W3_r) stays local, down is K-split + all-reduce (llama.cpp:693)
        h = x + part.numpy()

        def ffn_mid(r):
            gte = oracle.gemm_f64acc(h, shard_w("w1", r))
            return (gte / (1 + np.exp(-gte)) * oracle.gemm_f64acc(h, shard_w("w3", r))).astype(np.float32)

        dn = torch.from_numpy(oracle.gemm_f64acc(ffn_mid(rank), shard_w("w2", rank)))
        ctx.all_reduce(dn)
        want_dn = sum(oracle.gemm_f64acc(ffn_mid(r), shard_w("w2", r)).astype(np.float64) for r in range(world))
        assert np.allclose(dn.numpy(), want_dn, rtol=1e-5, atol=1e-5)
        # every rank ends with the same full activation
        gathered = [torch.empty_like(dn) for _ in range(world)]
        dist.all_gather(gathered, dn)
        assert all(torch.equal(gathered[0], t) for t in gathered)

        # ggml Q4_0 K-split: Q8_0 activation blocks are 32 wide, so slicing K at block boundaries commutes with quantisation
        wrows = oracle.quantize_q4_0(rng.normal(0, 0.02, (64, 512)).astype(np.float32))
        a = rng.normal(0, 1, (1, 512)).astype(np.float32)
        mine = tp.split_q4_0_rows(wrows, 512, world, rank, tp.SPLIT_K)
        kk = 512 // world
        p4 = torch.from_numpy(oracle.mul_mat_q4_0_f32(mine, a[:, rank * kk:(rank + 1) * kk]))
        ctx.all_reduce(p4)
        full4 = oracle.mul_mat_q4_0_f32(wrows, a)
        assert np.allclose(p4.numpy(), full4, rtol=1e-5, atol=1e-6)
        nrows = tp.split_q4_0_rows(wrows, 512, world, rank, tp.SPLIT_N)
        assert np.array_equal(oracle.mul_mat_q4_0_f32(nrows, a), full4[:, rank * 32:(rank + 1) * 32])
        ctx.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
        raise e


def test_two_rank_gloo_tensor_parallel_layer():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
