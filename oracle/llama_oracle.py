"""CPU oracle for the dnet pipelined-ring shard forward (Llama family).

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this module.  The
product path (dnet_b200/) never does; it fails loudly when the CUDA library is
missing.

PARITY UNPINNED.  The reference's arithmetic for this path lives in a
third-party dependency that is absent from /root/reference: ``mlx-lm==0.28.2``
on ``mlx`` (reference pyproject.toml:22,44-46).  The reference's own tests hold
no golden logits / token ids for any model block,
and mlx cannot be imported here.  This file therefore restates the published
algorithm of ``mlx_lm.models.llama`` (Attention / MLP / TransformerBlock),
``mx.fast.rms_norm``, ``mx.fast.rope`` (non-traditional), ``mx.fast.
scaled_dot_product_attention`` and ``mlx_lm.models.cache.KVCache`` and anchors
it on the reference's own call sites:

  * src/dnet/core/models/llama.py:56-66   embed / normalize / lm_project
  * src/dnet/core/models/llama.py:76-102  apply_single_layer (mask None at T=1,
    causal otherwise, per-local-layer cache entry)
  * src/dnet/shard/policies/fit_in_memory.py:34-209  the per-message loop
    (tokens -> embed -> cast to wire dtype; per-layer apply + cast to wire
    dtype; end shard: normalize + lm_project + Sampler.sample)
  * src/dnet/core/decoding/sampler.py:15-65  temp==0 -> argmax; logprob =
    v - logsumexp(v); top-k logprobs by full argsort (descending)

and is cross-checked in tests/test_oracle.py against an independent
implementation of the same architecture (HF transformers LlamaForCausalLM,
fp32) so that the *structure* (rotate-half RoPE, GQA grouping, scale, causal
mask, SwiGLU, pre-norm residuals) is pinned even though mlx bit patterns are
not.

Storage-dtype rounding points (what MLX does with bf16 arrays):
  rms_norm : fp32 math, y = T(T(x * rsqrt(mean(x^2)+eps)) * w)   (two roundings,
             same as the mx.fast.rms_norm fallback and HF LlamaRMSNorm)
  Linear   : fp32 accumulate, output rounded to T
  RoPE     : fp32 rotation of T inputs, one rounding to T; theta = pos * inv_freq
  SDPA     : fp32 scores/softmax/PV on T inputs, one rounding to T
  SwiGLU   : mx.compile'd  silu(g) * u  with per-primitive temporaries in T:
             s = T(sigmoid(g)); a = T(g * s); m = T(a * u)
  residual : T(x + r)
  policy   : cast to wire dtype after every layer (no-op when wire == T)
  lm_head  : Linear -> logits are T (bf16!) ; argmax takes the first maximal index
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


@dataclass
class OracleConfig:
    hidden_size: int
    num_attention_heads: int
    num_key_value_heads: int
    head_dim: int
    intermediate_size: int
    vocab_size: int
    num_hidden_layers: int
    rms_norm_eps: float = 1e-5
    rope_theta: float = 500000.0
    rope_scaling: Optional[dict] = None
    tie_word_embeddings: bool = False
    attention_bias: bool = False  # qwen2-style QKV bias
    num_local_experts: int = 0    # mixtral: experts per MoE layer (0 = dense MLP)
    num_experts_per_tok: int = 2

    @classmethod
    def from_dict(cls, d: dict) -> "OracleConfig":
        hd = d.get("head_dim") or d["hidden_size"] // d["num_attention_heads"]
        return cls(
            hidden_size=d["hidden_size"],
            num_attention_heads=d["num_attention_heads"],
            num_key_value_heads=d.get("num_key_value_heads", d["num_attention_heads"]),
            head_dim=hd,
            intermediate_size=d["intermediate_size"],
            vocab_size=d["vocab_size"],
            num_hidden_layers=d["num_hidden_layers"],
            rms_norm_eps=d.get("rms_norm_eps", 1e-5),
            rope_theta=d.get("rope_theta", 10000.0),
            rope_scaling=d.get("rope_scaling"),
            tie_word_embeddings=bool(d.get("tie_word_embeddings", False)),
            attention_bias=bool(d.get("attention_bias", False)),
            num_local_experts=int(d.get("num_local_experts", 0) or 0),
            num_experts_per_tok=int(d.get("num_experts_per_tok", 2) or 2),
        )


def rope_inv_freq(cfg: OracleConfig) -> torch.Tensor:
    """fp32 inverse frequencies, one per rotated pair (head_dim/2).

    mx.fast.rope derives inv_freq = base^(-2i/dims) in fp32; mlx_lm's Llama3RoPE
    (rope_scaling type "llama3") precomputes rescaled freqs and passes them in.
    Computed in float64 and rounded once to fp32 here; the CUDA side consumes the
    same table (dnet_b200.core.models.llama builds it with this formula).
    """
    half = cfg.head_dim // 2
    i = torch.arange(half, dtype=torch.float64)
    inv = cfg.rope_theta ** (-(2.0 * i) / cfg.head_dim)
    rs = cfg.rope_scaling
    if rs and rs.get("rope_type", rs.get("type")) == "llama3":
        factor = float(rs["factor"])
        low = float(rs.get("low_freq_factor", 1.0))
        high = float(rs.get("high_freq_factor", 4.0))
        old = float(rs["original_max_position_embeddings"])
        wavelen = 2.0 * math.pi / inv
        smooth = (old / wavelen - low) / (high - low)
        mid = (1.0 - smooth) * inv / factor + smooth * inv
        inv = torch.where(wavelen > old / low, inv / factor, torch.where(wavelen < old / high, inv, mid))
    return inv.to(torch.float32)


class OracleKV:
    """Contiguous growing KV for one layer (mlx_lm.models.cache.KVCache semantics:
    offset = tokens already stored; update_and_fetch appends then returns all)."""

    def __init__(self) -> None:
        self.k: Optional[torch.Tensor] = None  # [n_kv, n, hd] storage dtype
        self.v: Optional[torch.Tensor] = None
        self.offset = 0

    def update_and_fetch(self, k: torch.Tensor, v: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        if self.k is None:
            self.k, self.v = k.clone(), v.clone()
        else:
            self.k = torch.cat([self.k, k], dim=1)
            self.v = torch.cat([self.v, v], dim=1)
        self.offset = self.k.shape[1]
        return self.k, self.v


def mlx_affine_quantize(w: torch.Tensor, bits: int, group_size: int = 64):
    """Restatement of mlx.core.quantize (affine mode) along the last axis, as used by
    mlx_lm.models.cache.QuantizedKVCache (reference utils/model.py:505-554 selects it for
    kv_bits 4/8).  Restated from the published algorithm of mlx's quantize kernels, NOT verifiable in
    this tree (mlx is not vendored): PARITY UNPINNED.
        per group: edge = the bound with the larger magnitude; scale = max((max - min) / (2^bits - 1), 1e-7),
        signed so that edge / scale >= 0, then snapped so that edge is exactly representable
        (scale = edge / round(edge / scale)); bias = edge (0 when round(edge / scale) == 0);
        code = clip(round((w - bias) / scale), 0, 2^bits - 1)   -- with the fp32 scale / bias;
        the STORED scale / bias are rounded to w.dtype afterwards (what dequantisation then uses).
    Returns (codes uint8 [..., n], scales [..., n/group], biases [..., n/group]); dequantised value =
    scale * code + bias.  Rounding is round-half-to-even (rint), as torch.round."""
    n_bins = float((1 << bits) - 1)
    shape = w.shape
    g = w.to(torch.float32).reshape(*shape[:-1], shape[-1] // group_size, group_size)
    w_max, w_min = g.amax(-1, keepdim=True), g.amin(-1, keepdim=True)
    mask = w_min.abs() > w_max.abs()
    scales = torch.clamp((w_max - w_min) / n_bins, min=1e-7)
    scales = torch.where(mask, scales, -scales)
    edge = torch.where(mask, w_min, w_max)
    q0 = torch.round(edge / scales)
    scales = torch.where(q0 != 0, edge / q0, scales)
    biases = torch.where(q0 == 0, torch.zeros_like(edge), edge)
    codes = torch.clamp(torch.round((g - biases) / scales), 0, n_bins).to(torch.uint8)
    return codes.reshape(shape), scales.squeeze(-1).to(w.dtype), biases.squeeze(-1).to(w.dtype)


def mlx_affine_dequantize(codes: torch.Tensor, scales: torch.Tensor, biases: torch.Tensor, group_size: int = 64) -> torch.Tensor:
    shape = codes.shape
    g = codes.to(torch.float32).reshape(*shape[:-1], shape[-1] // group_size, group_size)
    return (g * scales.to(torch.float32).unsqueeze(-1) + biases.to(torch.float32).unsqueeze(-1)).reshape(shape)


class OracleQuantKV(OracleKV):
    """mlx_lm.models.cache.QuantizedKVCache semantics: every appended K / V row is quantised per group
    of 64 along head_dim (codes + scale + bias in the cache dtype); ``update_and_fetch`` returns the
    whole quantised cache, and attention runs on it with ``LlamaOracle.sdpa_quantized``.  Here the
    dequantised fp32 view (scale * code + bias, exactly what quantized_matmul multiplies by) is kept
    next to the packed parts."""

    quantized = True

    def __init__(self, bits: int, group_size: int = 64) -> None:
        super().__init__()
        self.bits, self.group_size = int(bits), int(group_size)
        self.parts: List[Tuple[Tuple[torch.Tensor, ...], Tuple[torch.Tensor, ...]]] = []

    def update_and_fetch(self, k: torch.Tensor, v: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        kq = mlx_affine_quantize(k, self.bits, self.group_size)
        vq = mlx_affine_quantize(v, self.bits, self.group_size)
        self.parts.append((kq, vq))
        kd = mlx_affine_dequantize(*kq, self.group_size)
        vd = mlx_affine_dequantize(*vq, self.group_size)
        self.k = kd if self.k is None else torch.cat([self.k, kd], dim=1)     # fp32 dequantised view
        self.v = vd if self.v is None else torch.cat([self.v, vd], dim=1)
        self.offset = self.k.shape[1]
        return self.k, self.v


class LlamaOracle:
    """Restates BaseRingModel's operator API (reference core/models/base.py:20-73)
    on torch-CPU.  ``dtype`` is the storage dtype T (bfloat16 for the parity
    configs; float32 turns every rounding point into the identity, used for the
    HF structural cross-check)."""

    def __init__(self, cfg: OracleConfig, weights: Dict[str, torch.Tensor], dtype=torch.bfloat16,
                 exact_linear: bool = False, f64_linear: bool = False):
        self.cfg = cfg
        self.dtype = dtype
        self.w = weights  # HF-style names, storage dtype
        self.inv_freq = rope_inv_freq(cfg)
        self.exact_linear = exact_linear
        # f64_linear: accumulate the dot products in float64 (a second, equally valid summation
        # order).  Used only to measure the oracle's own order-sensitivity ("noise floor"):
        # every bf16 rounding point can flip by one ulp when the accumulation order changes,
        # and downstream layers amplify a flip, so two correct implementations of this bf16
        # pipeline do not agree to 1e-3 end to end (DESIGN.md section "parity").
        self.f64_linear = f64_linear

    # -- primitives ---------------------------------------------------------
    def T(self, x: torch.Tensor) -> torch.Tensor:
        return x.to(self.dtype)

    def linear(self, x: torch.Tensor, name: str, bias: Optional[str] = None) -> torch.Tensor:
        W = self.w[name]
        if self.f64_linear:
            y = (x.to(torch.float64) @ W.to(torch.float64).T)
            if bias is not None and bias in self.w:
                y = y + self.w[bias].to(torch.float64)
            return self.T(y.to(torch.float32) if self.dtype == torch.float32 else y)
        rows = x.shape[0] if x.dim() > 1 else 1
        if self.exact_linear or self.dtype == torch.float32 or rows > 1:
            # fp32 accumulate of exact bf16 products.  (oneDNN's bf16 GEMM is used only for
            # single-row GEMV, where it was checked to be one correctly rounded fp32 sum; for
            # M > 1 it rounds differently, so prefill always takes this path.)
            y = x.to(torch.float32) @ W.to(torch.float32).T
        else:
            # oneDNN/ATen bf16 GEMM: fp32 accumulate, single rounding to bf16
            # (checked against an exact float64 product in tests/test_oracle.py)
            y = F.linear(x.to(self.dtype), W).to(torch.float32)
        if bias is not None and bias in self.w:
            y = y + self.w[bias].to(torch.float32)
        return self.T(y)

    def rms_norm(self, x: torch.Tensor, name: str) -> torch.Tensor:
        xf = x.to(torch.float32)
        inv = torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + self.cfg.rms_norm_eps)
        y = self.T(xf * inv)
        return self.T(y.to(torch.float32) * self.w[name].to(torch.float32))

    def rope(self, x: torch.Tensor, offset: int) -> torch.Tensor:
        """x: [heads, T, hd] storage dtype.  Non-traditional (rotate-half) RoPE,
        positions offset..offset+T-1 (mlx_lm llama Attention: rope(q, offset=cache.offset))."""
        Tn = x.shape[1]
        half = self.cfg.head_dim // 2
        pos = torch.arange(offset, offset + Tn, dtype=torch.float32)
        theta = pos[:, None] * self.inv_freq[None, :]  # fp32 product, like the kernel
        cos = torch.cos(theta.to(torch.float64)).to(torch.float32)
        sin = torch.sin(theta.to(torch.float64)).to(torch.float32)
        xf = x.to(torch.float32)
        x1, x2 = xf[..., :half], xf[..., half:]
        r1 = x1 * cos - x2 * sin
        r2 = x1 * sin + x2 * cos
        return self.T(torch.cat([r1, r2], dim=-1))

    def sdpa(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, offset: int) -> torch.Tensor:
        """q [H, T, hd]; k, v [n_kv, n, hd] with n = offset + T.  Causal,
        bottom-right aligned (query t sees keys 0..offset+t)."""
        Hq, Tn, hd = q.shape
        n_kv, n, _ = k.shape
        g = Hq // n_kv
        scale = hd ** -0.5
        qf = q.to(torch.float32) * scale
        kf = k.to(torch.float32).repeat_interleave(g, dim=0)
        vf = v.to(torch.float32).repeat_interleave(g, dim=0)
        s = qf @ kf.transpose(1, 2)  # [H, T, n]
        if Tn > 1:
            qpos = torch.arange(offset, offset + Tn)[:, None]
            kpos = torch.arange(n)[None, :]
            s = s.masked_fill(kpos > qpos, float("-inf"))
        p = torch.softmax(s, dim=-1)
        return self.T(p @ vf)

    def sdpa_quantized(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, offset: int) -> torch.Tensor:
        """mlx_lm.models.base.quantized_scaled_dot_product_attention, which mlx_lm's
        scaled_dot_product_attention dispatches to when the cache is a QuantizedKVCache:
            queries *= scale                                   (in T: one rounding)
            scores = quantized_matmul(queries, K^T)            (fp32 accumulate, output in T)
            scores = where(causal mask, scores, finfo.min)
            scores = softmax(scores, precise=True)             (fp32 math, output in T)
            out    = quantized_matmul(scores, V)               (fp32 accumulate, output in T)
        k, v are the dequantised fp32 views (scale * code + bias) of the quantised cache."""
        Hq, Tn, hd = q.shape
        n_kv, n, _ = k.shape
        g = Hq // n_kv
        qs = self.T(q.to(torch.float32) * (hd ** -0.5)).to(torch.float32)
        kf = k.to(torch.float32).repeat_interleave(g, dim=0)
        vf = v.to(torch.float32).repeat_interleave(g, dim=0)
        s = self.T(qs @ kf.transpose(1, 2)).to(torch.float32)
        if Tn > 1:
            qpos = torch.arange(offset, offset + Tn)[:, None]
            kpos = torch.arange(n)[None, :]
            s = s.masked_fill(kpos > qpos, torch.finfo(self.dtype).min)
        p = self.T(torch.softmax(s, dim=-1)).to(torch.float32)
        return self.T(p @ vf)

    # -- BaseRingModel operator API ----------------------------------------
    def embed(self, ids: torch.Tensor) -> torch.Tensor:
        return self.w["model.embed_tokens.weight"][ids.long()]

    def normalize(self, x: torch.Tensor) -> torch.Tensor:
        return self.rms_norm(x, "model.norm.weight")

    def lm_project(self, x: torch.Tensor, return_fp32: bool = False) -> torch.Tensor:
        name = "model.embed_tokens.weight" if self.cfg.tie_word_embeddings or "lm_head.weight" not in self.w else "lm_head.weight"
        if return_fp32:
            W = self.w[name]
            if W.numel() > (1 << 26):
                out = torch.empty(x.shape[0], W.shape[0], dtype=torch.float32)
                step = 1 << 14
                for r in range(0, W.shape[0], step):
                    out[:, r:r + step] = x.to(torch.float32) @ W[r:r + step].to(torch.float32).T
                return out
            return x.to(torch.float32) @ W.to(torch.float32).T
        return self.linear(x, name)

    def apply_single_layer(self, layer_idx: int, x: torch.Tensor, cache: OracleKV) -> torch.Tensor:
        """x: [T, hidden] storage dtype (the reference carries a leading batch of 1)."""
        c = self.cfg
        p = f"model.layers.{layer_idx}."
        Tn = x.shape[0]
        xn = self.rms_norm(x, p + "input_layernorm.weight")
        q = self.linear(xn, p + "self_attn.q_proj.weight", p + "self_attn.q_proj.bias")
        k = self.linear(xn, p + "self_attn.k_proj.weight", p + "self_attn.k_proj.bias")
        v = self.linear(xn, p + "self_attn.v_proj.weight", p + "self_attn.v_proj.bias")
        q = q.view(Tn, c.num_attention_heads, c.head_dim).transpose(0, 1)
        k = k.view(Tn, c.num_key_value_heads, c.head_dim).transpose(0, 1)
        v = v.view(Tn, c.num_key_value_heads, c.head_dim).transpose(0, 1)
        offset = cache.offset
        q = self.rope(q, offset)
        k = self.rope(k, offset)
        kk, vv = cache.update_and_fetch(k, v)
        a = (self.sdpa_quantized if getattr(cache, "quantized", False) else self.sdpa)(q, kk, vv, offset)  # [H, T, hd]
        a = a.transpose(0, 1).reshape(Tn, -1)
        r = self.linear(a, p + "self_attn.o_proj.weight")
        h = self.T(x.to(torch.float32) + r.to(torch.float32))
        hn = self.rms_norm(h, p + "post_attention_layernorm.weight")
        if c.num_local_experts:
            d = self.moe_block(hn, p + "block_sparse_moe.")
            return self.T(h.to(torch.float32) + d.to(torch.float32))
        g = self.linear(hn, p + "mlp.gate_proj.weight").to(torch.float32)
        u = self.linear(hn, p + "mlp.up_proj.weight").to(torch.float32)
        s = self.T(torch.sigmoid(g)).to(torch.float32)
        act = self.T(g * s).to(torch.float32)
        m = self.T(act * u)
        d = self.linear(m, p + "mlp.down_proj.weight")
        return self.T(h.to(torch.float32) + d.to(torch.float32))


    def moe_block(self, hn: torch.Tensor, p: str) -> torch.Tensor:
        """Sparse MoE FFN restated from mlx_lm.models.mixtral.MixtralSparseMoeBlock + switch_layers.SwitchGLU
        (mlx-lm 0.28.2; not in /root/reference -- PARITY UNPINNED like the rest of this file).  The reference hosts
        MoE families through the same BaseRingModel operator API (core/models/gpt_oss.py, deepseek_v2.py); mixtral is
        BASELINE.json configs[4].

          gates  = gate(x)                                   Linear, logits rounded to T
          inds   = argpartition(-gates, k-1)[..., :k]        the k largest logits (ties: lowest index first here)
          scores = softmax(gates[inds], precise=True)        fp32 math over the k selected logits, rounded to T
          y_e    = down_e(silu(gate_e(x)) * up_e(x))         per selected expert, SwiGLU temporaries in T as in the dense MLP
          y      = sum_e T(y_e * score_e)                    elementwise product rounded to T, then the sum over k rounded
                                                             to T (k = 2: one addition; selection order = descending logit)
        """
        c = self.cfg
        E, k = c.num_local_experts, c.num_experts_per_tok
        gates = self.linear(hn, p + "gate.weight").to(torch.float32)                 # [T, E]
        top = torch.sort(gates, dim=-1, descending=True, stable=True)
        inds = top.indices[:, :k]
        sel = top.values[:, :k]
        scores = self.T(torch.softmax(sel, dim=-1)).to(torch.float32)                # [T, k]
        out = torch.zeros(hn.shape[0], c.hidden_size, dtype=torch.float32)
        for t in range(hn.shape[0]):
            acc = None
            for j in range(k):
                e = int(inds[t, j])
                x1 = hn[t:t + 1]
                g = self.linear(x1, f"{p}experts.{e}.w1.weight").to(torch.float32)
                u = self.linear(x1, f"{p}experts.{e}.w3.weight").to(torch.float32)
                s_ = self.T(torch.sigmoid(g)).to(torch.float32)
                act = self.T(g * s_).to(torch.float32)
                m = self.T(act * u)
                y = self.linear(m, f"{p}experts.{e}.w2.weight").to(torch.float32)[0]
                term = self.T(y * scores[t, j]).to(torch.float32)
                acc = term if acc is None else self.T(acc + term).to(torch.float32)
            out[t] = acc
        return self.T(out)


@dataclass
class TokenResult:
    token_id: int
    logprob: float = 0.0
    top_logprobs: Dict[int, float] = field(default_factory=dict)


def sample_greedy(logits_T: torch.Tensor, req_logprobs: bool = False, req_top_logprobs: int = 0) -> TokenResult:
    """reference core/decoding/sampler.py:33-65 with temperature == 0 (argmax).

    ``logits_T`` is the last-position logits vector in the storage dtype (bf16 in
    the parity configs).  argmax returns the first maximal index; logsumexp
    accumulates in fp32 and rounds to T; log_probs = T(v - lse).
    """
    v = logits_T
    token_id = int(torch.argmax(v.to(torch.float32)).item())
    res = TokenResult(token_id=token_id)
    if req_logprobs or req_top_logprobs > 0:
        lse = torch.logsumexp(v.to(torch.float32), dim=-1).to(v.dtype)
        lp = (v.to(torch.float32) - lse.to(torch.float32)).to(v.dtype)
        if req_logprobs:
            res.logprob = float(lp[token_id].item())
        if req_top_logprobs > 0:
            order = torch.argsort(v.to(torch.float32), stable=True).flip(0)[:req_top_logprobs]
            for i in order.tolist():
                res.top_logprobs[int(i)] = float(lp[int(i)].item())
    return res


class OracleShard:
    """The FitInMemoryPolicy.process loop for one shard, restated
    (reference shard/policies/fit_in_memory.py:34-209).  Holds per-nonce KV like
    ShardRuntime.get_or_make_kv (reference shard/runtime.py:374-396)."""

    def __init__(self, model: LlamaOracle, assigned_layers: Sequence[int], wire_dtype=torch.bfloat16):
        self.model = model
        self.assigned = sorted(assigned_layers)
        self.assigned_set = set(self.assigned)
        self.wire = wire_dtype
        self.kv: Dict[str, Dict[int, OracleKV]] = {}

    def get_or_make_kv(self, nonce: str) -> Dict[int, OracleKV]:
        if nonce not in self.kv:
            self.kv[nonce] = {l: OracleKV() for l in self.assigned}
        return self.kv[nonce]

    def process(self, nonce: str, payload: torch.Tensor, dtype: str, layer_id: int,
                req_logprobs: bool = False, req_top_logprobs: int = 0):
        """Returns ("activation", x, last_layer) or ("final", TokenResult, last_layer)."""
        kv = self.get_or_make_kv(nonce)
        if dtype == "tokens":
            x = self.model.embed(payload.to(torch.int32)).to(self.wire)
        else:
            x = payload.to(self.wire)
        cur = layer_id + 1
        last = cur
        while True:
            if cur not in self.assigned_set:
                raise RuntimeError(f"layer {cur} not hosted")
            x = self.model.apply_single_layer(cur, x, kv[cur]).to(self.wire)
            last = cur
            if cur + 1 in self.assigned_set:
                cur += 1
                continue
            break
        nxt = last + 1
        if nxt >= self.model.cfg.num_hidden_layers:
            y = self.model.normalize(x)
            logits = self.model.lm_project(y[-1:])  # last position only: result-identical
            return "final", sample_greedy(logits[0], req_logprobs, req_top_logprobs), last
        return "activation", x, last


def make_weights(cfg: OracleConfig, seed: int, layers: Optional[Sequence[int]] = None,
                 dtype=torch.bfloat16, std: float = 0.02, with_api: bool = True) -> Dict[str, torch.Tensor]:
    """Deterministic synthetic checkpoint (HF names).  Uses numpy PCG64 streams
    keyed by (seed, tensor tag) so any subset of layers can be generated alone
    and bit-identically on any machine."""
    import numpy as np

    def rnd(tag: int, shape, s=std):
        g = np.random.Generator(np.random.PCG64([seed, tag]))
        a = g.standard_normal(size=shape, dtype=np.float32) * np.float32(s)
        return torch.from_numpy(a).to(dtype)

    H, F_, V = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size
    qd = cfg.num_attention_heads * cfg.head_dim
    kd = cfg.num_key_value_heads * cfg.head_dim
    w: Dict[str, torch.Tensor] = {}
    for l in (range(cfg.num_hidden_layers) if layers is None else layers):
        p = f"model.layers.{l}."
        b = 100 + 16 * l
        w[p + "self_attn.q_proj.weight"] = rnd(b + 0, (qd, H))
        w[p + "self_attn.k_proj.weight"] = rnd(b + 1, (kd, H))
        w[p + "self_attn.v_proj.weight"] = rnd(b + 2, (kd, H))
        w[p + "self_attn.o_proj.weight"] = rnd(b + 3, (H, qd))
        if cfg.num_local_experts:
            w[p + "block_sparse_moe.gate.weight"] = rnd(b + 12, (cfg.num_local_experts, H), 1.0)   # wide logits: clear top-k
            for e in range(cfg.num_local_experts):
                eb = 100000 + 4096 * l + 8 * e
                w[p + f"block_sparse_moe.experts.{e}.w1.weight"] = rnd(eb + 0, (F_, H))
                w[p + f"block_sparse_moe.experts.{e}.w3.weight"] = rnd(eb + 1, (F_, H))
                w[p + f"block_sparse_moe.experts.{e}.w2.weight"] = rnd(eb + 2, (H, F_))
        else:
            w[p + "mlp.gate_proj.weight"] = rnd(b + 4, (F_, H))
            w[p + "mlp.up_proj.weight"] = rnd(b + 5, (F_, H))
            w[p + "mlp.down_proj.weight"] = rnd(b + 6, (H, F_))
        # norm weights near 1 but not exactly 1 so the second rounding is exercised
        w[p + "input_layernorm.weight"] = (1.0 + rnd(b + 7, (H,), 0.1).to(torch.float32)).to(dtype)
        w[p + "post_attention_layernorm.weight"] = (1.0 + rnd(b + 8, (H,), 0.1).to(torch.float32)).to(dtype)
        if cfg.attention_bias:
            w[p + "self_attn.q_proj.bias"] = rnd(b + 9, (qd,))
            w[p + "self_attn.k_proj.bias"] = rnd(b + 10, (kd,))
            w[p + "self_attn.v_proj.bias"] = rnd(b + 11, (kd,))
    if with_api:
        w["model.embed_tokens.weight"] = rnd(1, (V, H), 1.0)
        w["model.norm.weight"] = (1.0 + rnd(2, (H,), 0.1).to(torch.float32)).to(dtype)
        if not cfg.tie_word_embeddings:
            w["lm_head.weight"] = rnd(3, (V, H))
    return w


def greedy_generate(cfg: OracleConfig, weights, prompt: Sequence[int], steps: int,
                    splits: Optional[List[List[int]]] = None, dtype=torch.bfloat16,
                    exact_linear: bool = False, collect=None):
    """Drive the ring like InferenceManager.generate_stream does
    (reference api/inference.py:135-212): prompt as one "tokens" message, then one
    token per step; each shard forwards its activation to the next."""
    model = LlamaOracle(cfg, weights, dtype, exact_linear=exact_linear)
    L = cfg.num_hidden_layers
    splits = splits or [list(range(L))]
    shards = [OracleShard(model, s, dtype) for s in splits]
    ids = list(prompt)
    out: List[TokenResult] = []
    y = torch.tensor(ids, dtype=torch.int32)
    for _ in range(steps):
        kind, payload, last = "activation", y, -1
        dt = "tokens"
        for sh in shards:
            kind, payload, last = sh.process("n0", payload, dt, last, True, 0)
            dt = "bfloat16"
            if collect is not None and kind == "activation":
                collect.append((last, payload.clone()))
        assert kind == "final"
        out.append(payload)
        y = torch.tensor([payload.token_id], dtype=torch.int32)
    return out
