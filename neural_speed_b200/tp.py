"""Tensor parallelism for the weight-only matmul path -- one process per GPU, torch.distributed.

The reference's scheme (docs/tensor_parallelism.md:17-27; models/llama/llama.cpp:114-125,592,693):
  * q/k/v and gate(w1)/up(w3) are split along N ("TP_1D_ROW", model_files.h:146-163): every rank owns n_head/W heads and
    n_ff/W hidden units; no communication, attention and SiLU*mul stay local.
  * o-proj and down(w2) are split along K ("TP_1D_COLUMN", model_files.h:171-185): every rank produces a partial
    [M, n_embd]; one sum all-reduce after each (ne_all_reduce, ne_layers.c:5466; reduce_add, parallel_context.cpp:47).
  * a BesTLA blob is split by unpacking to fp32 [K][N], slicing and RE-QUANTISING the slice with the blob's own
    attributes (bestla_split_weight, model_files.h:1538-1562 -> bestla_unpackweight_fp32 + bestla_packweight_copyattr).
    split_blob() below is that function; ggml Q4_0 rows are split by rows (N) or by 18-byte blocks (K) without
    re-quantisation (model_files.h:1619-1631,1650-1672 memcpy branches).

Host logic only: the collectives are torch.distributed (NCCL on GPUs, gloo in the CPU tests).  Fits-one-GPU models
(Llama-2-7B) are never sharded -- bench.py --gpus N runs replicas for those.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

SPLIT_N = "n"   # reference TP_1D_ROW
SPLIT_K = "k"   # reference TP_1D_COLUMN


@dataclass(frozen=True)
class LlamaShardPlan:
    """Per-rank shapes of one decoder layer's matmuls under W-way tensor parallelism."""
    world: int
    n_embd: int
    n_ff: int
    n_head: int
    n_head_kv: int
    group: int

    def __post_init__(self):
        w = self.world
        if self.n_head % w or self.n_head_kv % w:
            raise ValueError(f"n_head={self.n_head}/n_head_kv={self.n_head_kv} not divisible by world={w} (llama.cpp:121-124)")
        if self.n_ff % w or self.n_embd % w:
            raise ValueError("n_embd and n_ff must be divisible by the world size")
        if (self.n_embd // w) % self.group or (self.n_ff // w) % self.group:
            raise ValueError("K-split of o/down must cut at quantisation-group boundaries")

    @property
    def head_dim(self):
        return self.n_embd // self.n_head

    def shapes(self):
        """name -> (split, n_local, k_local)"""
        w, hd = self.world, self.head_dim
        return {
            "wq": (SPLIT_N, self.n_head // w * hd, self.n_embd), "wk": (SPLIT_N, self.n_head_kv // w * hd, self.n_embd),
            "wv": (SPLIT_N, self.n_head_kv // w * hd, self.n_embd), "wo": (SPLIT_K, self.n_embd, self.n_embd // w),
            "w1": (SPLIT_N, self.n_ff // w, self.n_embd), "w3": (SPLIT_N, self.n_ff // w, self.n_embd),
            "w2": (SPLIT_K, self.n_embd, self.n_ff // w),
        }


def split_blob(blob: np.ndarray, n: int, k: int, world: int, rank: int, split: str, qkv_fusion: bool = False) -> np.ndarray:
    """bestla_split_weight (model_files.h:1538-1562) through the C-ABI (ns_split_weight): unpack -> slice -> re-quantise
    with the source blob's attributes.  Returns the rank's blob (uint8, 64-byte aligned like the packer's output)."""
    from . import lib, _np_ptr
    L = lib()
    if split == SPLIT_N:
        dst_n, dst_k, n_rank, k_rank = n // world, k, rank, 0
    elif split == SPLIT_K:
        dst_n, dst_k, n_rank, k_rank = n, k // world, 0, rank
    else:
        raise ValueError(split)
    src = np.ascontiguousarray(blob, np.uint8)
    size = L.ns_split_weight_size(_np_ptr(src), dst_n, dst_k)
    if size == 0:
        raise ValueError("not a BesTLA k-block blob, or shard shape unsupported by the packer")
    raw = np.zeros(size + 64, np.uint8)
    dst = raw[(-raw.ctypes.data) % 64:][:size]
    if not L.ns_split_weight(_np_ptr(src), _np_ptr(dst), n, k, dst_n, dst_k, n_rank, k_rank, qkv_fusion):
        raise ValueError("ns_split_weight failed (shape mismatch with the blob header?)")
    return dst


def split_q4_0_rows(rows: np.ndarray, k: int, world: int, rank: int, split: str) -> np.ndarray:
    """ggml Q4_0 rows [N, K/32*18]: N-split = a contiguous block of rows (model_files.h:1628-1630); K-split = the rank's
    K/32/W blocks of every row (model_files.h:1664-1670).  No re-quantisation."""
    rows = np.ascontiguousarray(rows, np.uint8)
    n = rows.shape[0]
    if split == SPLIT_N:
        per = n // world
        return np.ascontiguousarray(rows[rank * per:(rank + 1) * per])
    per_row = rows.shape[1] // world
    if per_row % 18:
        raise ValueError("K-split must cut at Q4_0 block boundaries")
    return np.ascontiguousarray(rows[:, rank * per_row:(rank + 1) * per_row])


class TPContext:
    """init_parallel_context / get_tp_size / get_tp_rank / reduce_add of core/parallel_context.cpp on torch.distributed."""

    def __init__(self, backend: str | None = None, init: bool = True):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        if init and not dist.is_initialized():
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            dist.init_process_group(backend=backend)
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1

    def all_reduce(self, t):
        """in-place sum over ranks (reduce_add(sendBuf == recvBuf), ne_layers.c:5474)"""
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()


class TPLlamaMatmuls:
    """The matmul nodes of Llama decoder layers under tensor parallelism, device-resident (torch tensors for buffers,
    libns_b200 kernels for the matmuls, torch.distributed NCCL for the two all-reduces per layer).

    `layers` is a list of dicts name -> neural_speed_b200.Weight holding THIS RANK's shards (shapes per LlamaShardPlan).
    forward() maps the layer input x [M, n_embd] to the layer output contribution the way llama.cpp does around the
    attention core, which is supplied as `attn_fn(q, k, v) -> [M, n_head_local*head_dim]` (identity on q by default)."""

    def __init__(self, plan: LlamaShardPlan, layers, ctx: TPContext, stream=None):
        import torch
        self.plan, self.layers, self.ctx, self.torch = plan, layers, ctx, torch
        self.stream = stream if stream is not None else torch.cuda.current_stream()
        self.queue = C.c_void_p(self.stream.cuda_stream)

    def layer(self, li: int, x, attn_fn=None):
        from . import mul_mat, ffn_silu
        torch, p, lay = self.torch, self.plan, self.layers[li]
        m = x.shape[0]
        hd, w = p.head_dim, p.world
        nq, nkv = p.n_head // w * hd, p.n_head_kv // w * hd
        q = torch.empty(m, nq, device=x.device)
        k = torch.empty(m, nkv, device=x.device)
        v = torch.empty(m, nkv, device=x.device)
        for wt, out in ((lay["wq"], q), (lay["wk"], k), (lay["wv"], v)):   # GQA: n differs, so three plain matmuls
            mul_mat(wt, x.data_ptr(), p.n_embd, out.data_ptr(), out.shape[1], m, queue=self.queue)
        a = attn_fn(q, k, v) if attn_fn is not None else q
        o = torch.empty(m, p.n_embd, device=x.device)
        mul_mat(lay["wo"], a.data_ptr(), nq, o.data_ptr(), p.n_embd, m, queue=self.queue)
        self.ctx.all_reduce(o)                                            # llama.cpp:592
        h = x + o
        ff = p.n_ff // w
        tmp = torch.empty(2 if m > 4 else 1, m, ff, device=x.device)
        dn = torch.empty(m, p.n_embd, device=x.device)
        ffn_silu(lay["w1"], lay["w2"], lay["w3"], h.data_ptr(), p.n_embd, tmp.data_ptr(), dn.data_ptr(), p.n_embd, m, self.queue)
        self.ctx.all_reduce(dn)                                           # llama.cpp:693
        return h + dn
