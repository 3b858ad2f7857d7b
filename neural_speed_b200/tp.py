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

def current_queue(torch):
    """torch's current CUDA stream as a libns_b200 queue.  The legacy default stream has handle 0, which the C ABI reads as
    "use the library's own stream" -- work would then race with torch ops; cudaStreamLegacy (0x1) names it explicitly."""
    h = torch.cuda.current_stream().cuda_stream
    return C.c_void_p(h if h else 1)


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

    def enable_p2p(self, max_elems: int, queue=None) -> bool:
        """Set up the one-shot NVLink all-reduce (ns_comm_*): exchange the cudaIpc handles through torch.distributed and map
        every peer's buffer.  Returns False (and keeps NCCL) when the world is 1 or the tensors are not on CUDA."""
        torch, dist = self.torch, self.dist
        if self.world == 1 or not torch.cuda.is_available():
            return False
        from . import lib, last_error
        L = lib()
        comm = L.ns_comm_create(self.rank, self.world, max_elems, queue)
        if not comm:
            raise RuntimeError("ns_comm_create failed: " + last_error())
        hb = int(L.ns_comm_handle_bytes())
        mine = np.zeros(hb, np.uint8)
        if L.ns_comm_get_handle(C.c_void_p(comm), mine.ctypes.data_as(C.c_void_p)) != 0:
            raise RuntimeError("ns_comm_get_handle failed: " + last_error())
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t_mine = torch.from_numpy(mine).to(dev)
        gathered = [torch.empty_like(t_mine) for _ in range(self.world)]
        dist.all_gather(gathered, t_mine)
        allh = np.concatenate([g.cpu().numpy() for g in gathered])
        if L.ns_comm_open_peers(C.c_void_p(comm), allh.ctypes.data_as(C.c_void_p)) != 0:
            raise RuntimeError("ns_comm_open_peers failed: " + last_error())
        dist.barrier()
        self._comm, self._comm_max, self._comm_queue = C.c_void_p(comm), max_elems, queue
        return True

    def all_reduce(self, t, residual=None):
        """in-place sum over ranks (reduce_add(sendBuf == recvBuf), ne_layers.c:5474); `residual` (same shape) is added
        once after the reduction.  Uses the one-shot NVLink kernel when enable_p2p() was called and the tensor qualifies."""
        comm = getattr(self, "_comm", None)
        if (comm is not None and t.is_cuda and t.dtype == self.torch.float32 and t.is_contiguous() and t.numel() <= self._comm_max
                and t.numel() % 4 == 0):
            from . import lib, last_error
            q = self._comm_queue if self._comm_queue is not None else current_queue(self.torch)
            rc = lib().ns_comm_all_reduce_f32(comm, C.c_void_p(t.data_ptr()), t.numel(),
                                              C.c_void_p(residual.data_ptr()) if residual is not None else None, q)
            if rc != 0:
                raise RuntimeError("ns_comm_all_reduce_f32 failed: " + last_error())
            return t
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        if residual is not None:
            t += residual
        return t

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()


class TPLlamaMatmuls:
    """The matmul nodes of Llama decoder layers under tensor parallelism, device-resident (torch tensors for buffers,
    libns_b200 kernels for the matmuls, one sum all-reduce after o-proj and after down-proj with the residual add folded in).

    `layers` is a list of dicts name -> neural_speed_b200.Weight holding THIS RANK's shards (shapes per LlamaShardPlan).
    layer() maps the layer input x [M, n_embd] to the layer output the way llama.cpp does around the attention core, which
    is supplied as `attn_fn(q, k, v) -> [M, n_head_local*head_dim]` (identity on q by default).  Buffers are allocated once
    per M, kernels go to torch's CURRENT stream, so a whole token can be captured in a torch.cuda.CUDAGraph."""

    def __init__(self, plan: LlamaShardPlan, layers, ctx: TPContext):
        import torch
        self.plan, self.layers, self.ctx, self.torch = plan, layers, ctx, torch
        self._bufs = {}

    def _buffers(self, m, device):
        b = self._bufs.get(m)
        if b is None:
            torch, p = self.torch, self.plan
            hd, w = p.head_dim, p.world
            nq, nkv, ff = p.n_head // w * hd, p.n_head_kv // w * hd, p.n_ff // w
            b = dict(q=torch.empty(m, nq, device=device), k=torch.empty(m, nkv, device=device), v=torch.empty(m, nkv, device=device),
                     o=torch.empty(m, p.n_embd, device=device), h=torch.empty(m, p.n_embd, device=device),
                     tmp=torch.empty(2 if m > 4 else 1, m, ff, device=device), dn=torch.empty(m, p.n_embd, device=device))
            self._bufs[m] = b
        return b

    def layer(self, li: int, x, attn_fn=None, out=None):
        from . import mul_mat, ffn_silu
        torch, p, lay = self.torch, self.plan, self.layers[li]
        m = x.shape[0]
        b = self._buffers(m, x.device)
        queue = current_queue(torch)
        q, k, v = b["q"], b["k"], b["v"]
        for wt, dst in ((lay["wq"], q), (lay["wk"], k), (lay["wv"], v)):   # GQA: n differs, so three plain matmuls
            mul_mat(wt, x.data_ptr(), p.n_embd, dst.data_ptr(), dst.shape[1], m, queue=queue)
        a = attn_fn(q, k, v) if attn_fn is not None else q
        o = b["o"]
        mul_mat(lay["wo"], a.data_ptr(), a.shape[1], o.data_ptr(), p.n_embd, m, queue=queue)
        h = self.ctx.all_reduce(o, residual=x)                            # llama.cpp:592 + inpFF = cur + inpSA (:598)
        dn = out if out is not None else b["dn"]
        ffn_silu(lay["w1"], lay["w2"], lay["w3"], h.data_ptr(), p.n_embd, b["tmp"].data_ptr(), dn.data_ptr(), p.n_embd, m, queue)
        return self.ctx.all_reduce(dn, residual=h)                        # llama.cpp:693 + cur = cur + inpFF (:698)
