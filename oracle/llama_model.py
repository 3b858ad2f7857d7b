"""oracle/llama_model.py -- TEST INFRASTRUCTURE ONLY.  CPU restatement (numpy + the pinned C oracle for the matmuls) of
the reference's Llama eval graph, models/llama/llama.cpp:190-720, ggml (non-fused-attention) path:

  get_rows -> per layer [rms_norm * w (kernel_ref.h:2199-2225) -> Q/K/V matmuls -> rope mode 0 (ne_layers.c:9380-9396,
  theta_base *= theta_scale iteratively in fp32) -> fp16 KV cache -> K.Q with Q rounded to fp16
  (ne_compute_forward_mul_mat_f16_f32) -> scale -> causal mask -> soft_max with the fp16 exp table
  (ne_layers.c:8923-8945) -> V.P with P rounded to fp16 -> wo + residual -> rms_norm * w -> silu(w1 x) * (w3 x) -> w2 +
  residual] -> rms_norm * w -> output matmul.  Greedy pick: lowest index among maxima (model_utils.cpp:2963-2985).

Parity status: the MATMULS are the pinned oracle (bit-exact with oracle/_ref); the element-wise ops around them are
restated from the files above and are NOT pinned against a build of the reference (its rms_norm/mul/add run through
BesTLA kernels that need xbyak) -- "parity unpinned" for those; the GPU engine is held to the north-star tolerance
(1e-2 on logits, greedy ids equal wherever the top-2 margin exceeds that tolerance).
"""
import numpy as np

import oracle


def _f16(x):
    return np.asarray(x, np.float32).astype(np.float16).astype(np.float32)


class OracleLlama:
    def __init__(self, hp: dict, tok_embd, out_norm, output_rows, layers, fmt="q4_0"):
        """layers: list of dicts with attn_norm, ffn_norm (f32 [E]) and wq, wk, wv, wo, w1, w2, w3 as Q4_0 row arrays
        (uint8 [N, K/32*18]); output_rows likewise (or Q6_K rows when fmt_out == 'q6_K')."""
        self.hp = dict(hp)
        self.tok_embd = np.asarray(tok_embd, np.float32)
        self.out_norm = np.asarray(out_norm, np.float32)
        self.output_rows = output_rows
        self.layers = layers
        self.out_fmt = fmt
        E, H, HK = hp["n_embd"], hp["n_head"], hp["n_head_kv"]
        self.hd = E // H
        self.kc = np.zeros((hp["n_layer"], HK, hp["n_ctx"], self.hd), np.float16)
        self.vc = np.zeros_like(self.kc)

    @staticmethod
    def _mm(rows, a):
        return oracle.mul_mat_q4_0_f32(rows, np.ascontiguousarray(a, np.float32))

    def _rms(self, x, w):
        eps = np.float32(self.hp.get("norm_eps", 1e-6))
        out = np.empty_like(x)
        for r in range(x.shape[0]):
            ms = np.float32(0)
            for v in x[r]:                                   # T mean_square += src*src, sequential fp32 (kernel_ref.h:2204-2207)
                ms = np.float32(ms + np.float32(v * v))
            rms = np.float32(np.sqrt(np.float32(ms / np.float32(x.shape[1]) + eps)))
            inv = np.float32(1.0) / rms
            out[r] = (x[r] * inv) * w
        return out

    def _rope(self, x, pos):
        """x [n_head, hd] fp32, in place semantics of ne_rope_inplace mode 0"""
        hd = self.hd
        theta_scale = np.float32(np.power(np.float32(self.hp.get("rope_theta", 10000.0)), np.float32(-2.0) / np.float32(hd)))
        freq_scale = np.float32(1.0) / np.float32(self.hp.get("rope_scale", 1.0))
        out = x.copy()
        theta = np.float32(pos)
        for i0 in range(0, hd, 2):
            th = np.float32(freq_scale * theta)
            c, s = np.float32(np.cos(th)), np.float32(np.sin(th))
            theta = np.float32(theta * theta_scale)
            x0, x1 = x[:, i0], x[:, i0 + 1]
            out[:, i0] = x0 * c - x1 * s
            out[:, i0 + 1] = x0 * s + x1 * c
        return out

    def eval(self, tokens, n_past):
        hp = self.hp
        E, H, HK, hd = hp["n_embd"], hp["n_head"], hp["n_head_kv"], self.hd
        n = len(tokens)
        x = self.tok_embd[np.asarray(tokens)].astype(np.float32)
        scale = np.float32(1.0) / np.float32(np.sqrt(np.float32(hd)))
        for il, L in enumerate(self.layers):
            cur = self._rms(x, L["attn_norm"])
            q = self._mm(L["wq"], cur).reshape(n, H, hd)
            k = self._mm(L["wk"], cur).reshape(n, HK, hd)
            v = self._mm(L["wv"], cur).reshape(n, HK, hd)
            attn = np.zeros((n, H, hd), np.float32)
            for t in range(n):
                pos = n_past + t
                q[t] = self._rope(q[t], pos)
                self.kc[il, :, pos, :] = self._rope(k[t], pos).astype(np.float16)
                self.vc[il, :, pos, :] = v[t].astype(np.float16)
            for t in range(n):
                ln = n_past + t + 1
                for h in range(H):
                    hk = h // (H // HK)
                    kk = self.kc[il, hk, :ln].astype(np.float32)              # [ln, hd]
                    s = (kk @ _f16(q[t, h])).astype(np.float32) * scale       # fp16 K x fp16(Q), fp32 accumulate
                    mx = s.max()
                    e = _f16(np.exp(_f16(s - mx)))                            # table_exp_f16
                    p = e * np.float32(1.0 / np.float64(e.astype(np.float64).sum()))
                    attn[t, h] = (_f16(p)[None, :] @ self.vc[il, hk, :ln].astype(np.float32))[0]
            inp_ff = self._mm(L["wo"], attn.reshape(n, E)) + x
            cur = self._rms(inp_ff, L["ffn_norm"])
            g = self._mm(L["w1"], cur)
            silu = np.array([[oracle.lib().orc_silu(float(z)) for z in row] for row in g], np.float32)
            mid = silu * self._mm(L["w3"], cur)
            x = self._mm(L["w2"], mid) + inp_ff
        last = self._rms(x[-1:], self.out_norm)
        if self.out_fmt == "q6_K":
            return oracle.mul_mat_q6_K_f32(self.output_rows, last)[0]
        return self._mm(self.output_rows, last)[0]


def greedy(logits):
    return int(oracle.argmax(logits))
