"""Llama ring model (reference src/dnet/core/models/llama.py:11-137) on libdnet_b200."""
from __future__ import annotations

import math
from typing import Any, Dict, List, Optional, Tuple

import torch

from .base import BaseRingModel, KVHandle, _stream_ptr
from dnet_b200 import _cabi


def rope_inv_freq(cfg: dict) -> torch.Tensor:
    """fp32 inverse frequencies per rotated pair.  mx.fast.rope derives base^(-2i/d);
    mlx_lm's Llama3RoPE (rope_scaling type "llama3") rescales them."""
    hd = cfg.get("head_dim") or cfg["hidden_size"] // cfg["num_attention_heads"]
    half = hd // 2
    i = torch.arange(half, dtype=torch.float64)
    inv = float(cfg.get("rope_theta", 10000.0)) ** (-(2.0 * i) / hd)
    rs = cfg.get("rope_scaling")
    if rs and rs.get("rope_type", rs.get("type")) == "llama3":
        factor = float(rs["factor"])
        low = float(rs.get("low_freq_factor", 1.0))
        high = float(rs.get("high_freq_factor", 4.0))
        old = float(rs["original_max_position_embeddings"])
        wavelen = 2.0 * math.pi / inv
        smooth = (old / wavelen - low) / (high - low)
        mid = (1.0 - smooth) * inv / factor + smooth * inv
        inv = torch.where(wavelen > old / low, inv / factor, torch.where(wavelen < old / high, inv, mid))
    elif rs and rs.get("rope_type", rs.get("type")) not in (None, "default", "linear"):
        raise NotImplementedError(f"rope_scaling {rs} unsupported")
    elif rs and rs.get("rope_type", rs.get("type")) == "linear":
        inv = inv / float(rs["factor"])
    return inv.to(torch.float32)


class LlamaRingModel(BaseRingModel):
    """Constructs only the locally assigned decoder blocks and exposes layer-wise
    application.  Activations are torch bf16 CUDA tensors shaped (1, T, H) or (T, H)."""

    model_type = "llama"

    def __init__(self, model_config: Any, assigned_layers: Optional[List[int]] = None, is_api_layer: bool = False,
                 kv_pool_pages: int = 0, wire_dtype: str = "bfloat16", kv_bits: int = 0, kv_group: int = 64):
        if is_api_layer and assigned_layers:
            raise RuntimeError("API layer doesn't handle layers")
        self.model_config = model_config
        self.is_api_layer = is_api_layer
        self.config = dict(model_config)
        self.tie_word_embeddings = bool(self.config.get("tie_word_embeddings", False))
        if self.config.get("attention_bias", False) and self.model_type == "llama":
            pass  # biases are bound when present in the checkpoint
        if self.config.get("mlp_bias", False):
            raise NotImplementedError("mlp_bias is not supported")
        self.layers = sorted(assigned_layers or [])
        self._create(self.config, self.layers, rope_inv_freq(self.config), kv_pool_pages, wire_dtype, kv_bits, kv_group)

    # -- operator API ----------------------------------------------------------------------
    def embed(self, x: torch.Tensor, stream=None) -> torch.Tensor:
        """x: int32 token ids, shape (T,) or (1, T) on the device -> (1, T, H) bf16."""
        ids = x.reshape(-1).to(torch.int32)
        if not ids.is_cuda:
            raise ValueError("embed expects device token ids")
        out = torch.empty(ids.numel(), self.hidden_size, dtype=torch.bfloat16, device=ids.device)
        _cabi.check(self._lib.dn_embed(self._h, ids.data_ptr(), ids.numel(), out.data_ptr(), _stream_ptr(stream)))
        return out.view(1, ids.numel(), self.hidden_size) if x.dim() == 2 else out

    def normalize(self, x: torch.Tensor) -> torch.Tensor:
        """Final norm alone is not exposed by the C ABI (it is fused into the head kernel);
        normalize() returns x tagged so lm_project() runs the fused norm+head."""
        return _Normed(x)

    def lm_project(self, x) -> torch.Tensor:
        """fused final-norm + lm_head on the last position -> (1, 1, V) bf16 logits.
        (The reference projects all T positions and the sampler keeps the last one;
        last-position-only is result-identical, SURVEY.md Appendix C.)"""
        if not isinstance(x, _Normed):
            raise ValueError("lm_project expects the output of normalize()")
        h = x.x.reshape(-1, self.hidden_size)
        _, b16 = self.head_logits(h, want_f32=False, want_bf16=True)
        return b16.view(1, 1, -1)

    def apply_single_layer(self, layer_idx: int, x: torch.Tensor, cache: Optional[KVHandle] = None,
                           stream=None) -> torch.Tensor:
        if layer_idx not in self.abs_to_local:
            raise RuntimeError(f"Layer {layer_idx} not hosted on this model instance")
        if cache is None:
            raise ValueError("apply_single_layer needs the nonce's KV handle")
        shp = x.shape
        h = x.reshape(-1, self.hidden_size)
        if not h.is_contiguous():
            h = h.contiguous()
        self.window_forward([layer_idx], h, cache, stream)
        return h.view(shp)

    def sanitize(self, weights):
        weights = {k: v for k, v in weights.items() if "self_attn.rotary_emb.inv_freq" not in k}
        if self.tie_word_embeddings:
            weights.pop("lm_head.weight", None)
        return weights

    @property
    def decoding_layers(self):
        return self.layers

    @property
    def head_dim(self) -> Tuple[int, int]:
        hd = self.config.get("head_dim") or self.config["hidden_size"] // self.config["num_attention_heads"]
        return (hd, hd)

    @property
    def n_kv_heads(self) -> int:
        return self.config.get("num_key_value_heads", self.config["num_attention_heads"])

    @property
    def num_layers(self) -> int:
        return len(self.layers)


class _Normed:
    __slots__ = ("x",)

    def __init__(self, x):
        self.x = x


class Qwen2RingModel(LlamaRingModel):
    """qwen2 = llama block + q/k/v bias (mlx_lm.models.qwen2); the reference has no qwen2
    wrapper, BASELINE config 3 needs one."""

    model_type = "qwen2"


class MixtralRingModel(LlamaRingModel):
    """mixtral = llama attention + a sparse MoE FFN (mlx_lm.models.mixtral: router -> top-k experts -> softmax over
    the selected logits -> SwiGLU experts -> weighted sum).  The reference has no mixtral wrapper; it hosts its MoE
    families (gpt_oss, deepseek_v2) through this same operator API, and BASELINE.json configs[4] names
    Mixtral-8x7B.  Expert selection happens on the device, so MoE layers run on the per-op path
    (dn_window_forward; CUDA-graph capturable), not in the persistent step kernel."""

    model_type = "mixtral"
    step_kernel_ok = False
