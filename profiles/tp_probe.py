"""Tensor-parallel decode probe: Llama-2-70B-shaped layers (n_embd 8192, n_ff 28672, 64 heads,
8 KV heads), Q4_0 shards created directly per rank, M = 1.  Times `n_layer` decoder layers' matmul path per token with
(a) NCCL all-reduce, (b) the one-shot NVLink all-reduce (ns_comm_*), each replayed from one CUDA graph per token.
Run on W GPUs:  python profiles/tp_probe.py W [n_layer]   (spawns W processes; writes gpurun_out/tp_probe_W.json)"""
import json
import os
import socket
import sys

import torch


def worker(rank, world, port, n_layer, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import ctypes as C
    import torch.distributed as dist
    import neural_speed_b200 as ns
    from neural_speed_b200 import tp
    torch.cuda.set_device(rank)
    L = ns.lib()
    L.bestla_init()
    ctx = tp.TPContext(backend="nccl")
    plan = tp.LlamaShardPlan(world, 8192, 28672, 64, 8, 32)
    shapes = plan.shapes()

    def mk(n, k):
        w = torch.randn(n, k, device="cuda") * 0.02
        rows = torch.empty(n * (k // 32) * 18, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        assert L.ns_device_quantize_q4_0(C.c_void_p(w.data_ptr()), C.c_void_p(rows.data_ptr()), n, k, None) == 0
        L.bestla_device_sync(None)
        return ns.Weight.from_q4_0_device(rows.data_ptr(), n, k, (k // 32) * 18)

    torch.manual_seed(1234 + rank)
    layers = [{name: mk(n, k) for name, (_, n, k) in shapes.items()} for _ in range(n_layer)]
    bytes_per_token = sum(w.algorithmic_bytes for lay in layers for w in lay.values())
    eng = tp.TPLlamaMatmuls(plan, layers, ctx)
    x0 = torch.randn(1, 8192, device="cuda")
    res = {}
    for mode in ("nccl", "p2p"):
        if mode == "p2p":
            ctx.enable_p2p(8192)
        stream = torch.cuda.Stream()
        xa, xb = x0.clone(), torch.empty_like(x0)

        def token():
            cur, nxt = xa, xb
            for li in range(n_layer):
                out = eng.layer(li, cur, out=nxt)
                cur, nxt = out, cur
            return cur

        with torch.cuda.stream(stream):
            for _ in range(3):
                token()
            torch.cuda.synchronize()
            dist.barrier()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                y = token()
            for _ in range(5):
                g.replay()
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            steps = 50
            e0.record(stream)
            for _ in range(steps):
                g.replay()
            e1.record(stream)
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        assert bool(torch.isfinite(y).all())
        res[mode] = {"ms_per_token_layers": ms, "us_per_layer": ms * 1e3 / n_layer,
                     "per_gpu_GBps": bytes_per_token / (ms * 1e-3) / 1e9}
    if rank == 0:
        full = {"world": world, "n_layer_measured": n_layer, "model_layers": 80, "per_rank_packed_bytes_per_layer": bytes_per_token // n_layer,
                "all_reduces_per_layer": 2, "message_bytes": 8192 * 4, **res,
                "extrapolated_80_layers_tokens_per_s": {m: 1000.0 / (res[m]["us_per_layer"] * 80 / 1e3) for m in res},
                "note": "matmul path + all-reduces only (identity attention core); extrapolation = 80 x the measured per-layer time, "
                        "lm_head excluded"}
        os.makedirs(os.path.dirname(out_path), exist_ok=True)
        json.dump(full, open(out_path, "w"), indent=1)
        print(json.dumps(full))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    n_layer = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    import torch.multiprocessing as mp
    mp.spawn(worker, args=(world, port, n_layer, os.path.join(os.getcwd(), "gpurun_out", f"tp_probe_{world}.json")), nprocs=world)
