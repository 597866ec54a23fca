"""Base class for ring topology models (reference src/dnet/core/models/base.py:10-486).

Same operator API -- embed / normalize / lm_project / apply_single_layer /
load_weights / unload_layers -- but a model instance is a handle on a ``dn_model`` in
libdnet_b200.so: the arithmetic is hand-written sm_100a CUDA, tensors are torch CUDA
tensors used purely as device memory, and weights are BORROWED from the WeightCache
(binding is by pointer, exactly like the reference rebinding by reference).
"""
from __future__ import annotations

import ctypes as C
from abc import ABC, abstractmethod
from typing import Any, Dict, List, Optional, Tuple

import torch

from dnet_b200 import _cabi
from dnet_b200._cabi import DN_W_COUNT, SUFFIX_TO_SLOT


class KVHandle:
    """Per-nonce paged KV + device step state (replaces the per-layer list of mlx_lm
    KVCache objects built by make_cache, reference utils/model.py:470-555)."""

    def __init__(self, model: "BaseRingModel", max_tokens: int):
        self.model = model
        self._lib = _cabi.load()
        p = C.c_void_p()
        _cabi.check(self._lib.dn_kv_create(model._h, int(max_tokens), C.byref(p)))
        self._h = p.value
        self.max_tokens = int(max_tokens)
        self.last_used = 0.0

    @property
    def offset(self) -> int:
        return int(self._lib.dn_kv_offset(self._h))

    def reset(self, stream: int = 0) -> None:
        _cabi.check(self._lib.dn_kv_reset(self._h, stream))

    def advance(self, T: int, stream: int = 0) -> None:
        _cabi.check(self._lib.dn_kv_advance(self._h, int(T), stream))

    def seek(self, pos: int, stream: int = 0) -> None:
        _cabi.check(self._lib.dn_kv_seek(self._h, int(pos), stream))

    def note_advance(self, T: int) -> None:
        _cabi.check(self._lib.dn_kv_note_advance(self._h, int(T)))

    def set_token(self, token: int, stream: int = 0) -> None:
        _cabi.check(self._lib.dn_kv_set_token(self._h, int(token), stream))

    @property
    def token_ptr(self) -> int:
        return int(self._lib.dn_kv_token_ptr(self._h))

    def free(self) -> None:
        if self._h:
            self._lib.dn_kv_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _stream_ptr(stream) -> int:
    if stream is None:
        return int(torch.cuda.current_stream().cuda_stream)
    if isinstance(stream, int):
        return stream
    return int(stream.cuda_stream)


def _decompose(T: int, tmax: int, big: int = 0) -> List[int]:
    """Split T into chunk sizes the kernels are instantiated for: chunks of 16..``big`` tokens go
    to the tensor-core prefill GEMMs, the remainder to the (4, 2, 1)-token GEMV kernels."""
    out: List[int] = []
    while big >= 16 and T >= 16:
        c = min(T, big)
        out.append(c)
        T -= c
    c = 4
    while c > tmax:
        c >>= 1
    while T > 0:
        while c > T:
            c >>= 1
        out.append(c)
        T -= c
    return out


class BaseRingModel(ABC):
    """Base class for models used in ring topology."""

    model_type: Optional[str] = None

    # -- construction shared by subclasses --------------------------------------------
    def _create(self, cfg: dict, assigned_layers: List[int], inv_freq: torch.Tensor, kv_pool_pages: int,
                wire_dtype: str = "bfloat16", kv_bits: int = 0, kv_group: int = 64) -> None:
        self._lib = _cabi.load()
        dev = torch.cuda.current_device() if torch.cuda.is_available() else 0
        _cabi.init(dev)
        hd = cfg.get("head_dim") or cfg["hidden_size"] // cfg["num_attention_heads"]
        dt = str(cfg.get("torch_dtype", cfg.get("dtype", "bfloat16")))
        if "bfloat16" not in dt and "bf16" not in dt:
            raise ValueError(f"dnet_b200 supports bfloat16 checkpoints only (got {dt})")
        if wire_dtype not in ("bfloat16", "bf16"):
            raise ValueError(
                "wire dtype must equal the model dtype: set DNET_TRANSPORT_WIRE_DTYPE=bf16 for bf16 models "
                "(the reference promotes fp16 x bf16 to fp32 compute; that mixed path is not rebuilt)")
        mc = _cabi.ModelCfg(
            hidden=cfg["hidden_size"], n_heads=cfg["num_attention_heads"],
            n_kv_heads=cfg.get("num_key_value_heads", cfg["num_attention_heads"]), head_dim=hd,
            ffn=cfg["intermediate_size"], vocab=cfg["vocab_size"], n_layers_total=cfg["num_hidden_layers"],
            rms_eps=float(cfg.get("rms_norm_eps", 1e-5)), tie_embeddings=int(bool(cfg.get("tie_word_embeddings", False))),
            dtype=0, wire_dtype=0, kv_page_tokens=64, kv_pool_pages=int(kv_pool_pages), kv_bits=int(kv_bits),
            kv_group=int(kv_group), n_experts=int(cfg.get("num_local_experts", 0) or 0),
            top_k=int(cfg.get("num_experts_per_tok", 0) or 0))
        self.kv_bits = int(kv_bits)
        self.n_experts = int(mc.n_experts)
        layers = sorted(assigned_layers or [])
        arr = (C.c_int32 * max(1, len(layers)))(*layers)
        inv = inv_freq.to(torch.float32).contiguous().cpu()
        invp = (C.c_float * inv.numel())(*inv.tolist())
        h = C.c_void_p()
        _cabi.check(self._lib.dn_model_create(C.byref(mc), arr, len(layers), invp, C.byref(h)))
        self._h = h.value
        self._cfg_struct = mc
        self.abs_to_local: Dict[int, int] = {l: i for i, l in enumerate(layers)}
        self._bound: Dict[int, Dict[str, torch.Tensor]] = {}
        self._api: Dict[str, torch.Tensor] = {}
        self.hidden_size = cfg["hidden_size"]
        self.vocab_size = cfg["vocab_size"]
        self.max_chunk = int(self._lib.dn_model_max_chunk(self._h))
        self.max_prefill_chunk = int(self._lib.dn_model_max_prefill_chunk(self._h))

    # -- abstract operator API -----------------------------------------------------------
    @abstractmethod
    def embed(self, x): ...

    @abstractmethod
    def normalize(self, x): ...

    @abstractmethod
    def lm_project(self, x): ...

    @abstractmethod
    def apply_single_layer(self, layer_idx: int, x, cache: Optional[Any] = None): ...

    @property
    @abstractmethod
    def decoding_layers(self) -> Any: ...

    @property
    @abstractmethod
    def head_dim(self) -> Tuple[int, int]: ...

    @property
    @abstractmethod
    def n_kv_heads(self) -> int: ...

    @property
    @abstractmethod
    def num_layers(self) -> int: ...

    # -- weights -------------------------------------------------------------------------
    def load_weights(self, file_or_weights, strict: bool = False):
        """Bind weights for this shard (reference base.py:111-195): accepts
        ``model.layers.N.*`` / ``layers.N.*`` (absolute N) and ``(model.)embed_tokens.*``,
        ``(model.)norm.*``, ``lm_head.*``; non-hosted layers are skipped."""
        if isinstance(file_or_weights, dict):
            wdict = dict(file_or_weights)
        elif isinstance(file_or_weights, (list, tuple)):
            wdict = {(k.decode("utf-8") if isinstance(k, (bytes, bytearray)) else str(k)): v for k, v in file_or_weights}
        else:
            raise TypeError("load_weights expects a dict or a list of (name, tensor)")
        if hasattr(self, "sanitize"):
            wdict = self.sanitize(wdict)
        per_layer: Dict[int, Dict[str, torch.Tensor]] = {}
        tie = bool(getattr(self, "tie_word_embeddings", False))
        api_changed = False
        for key, value in wdict.items():
            if key.startswith("_"):
                continue
            if key.startswith("model.layers.") or key.startswith("layers."):
                parts = key.split(".")
                idx_pos = 2 if parts[0] == "model" else 1
                try:
                    abs_idx = int(parts[idx_pos])
                except Exception:
                    continue
                if abs_idx not in self.abs_to_local:
                    continue
                per_layer.setdefault(abs_idx, {})[".".join(parts[idx_pos + 1:])] = value
                continue
            bare = key[6:] if key.startswith("model.") else key
            if bare.startswith("embed_tokens.") or bare.startswith("norm.") or (bare.startswith("lm_head.") and not tie):
                self._api[bare] = value
                api_changed = True
        for abs_idx, tensors in per_layer.items():
            ptrs = (C.c_void_p * DN_W_COUNT)()
            for suffix, t in tensors.items():
                slot = SUFFIX_TO_SLOT.get(suffix)
                if slot is None:
                    if strict:
                        raise ValueError(f"unexpected tensor {suffix} for layer {abs_idx}")
                    continue
                if not t.is_cuda or t.dtype != torch.bfloat16 or not t.is_contiguous():
                    raise ValueError(f"layer {abs_idx} {suffix}: expected a contiguous bf16 CUDA tensor")
                ptrs[slot] = t.data_ptr()
            _cabi.check(self._lib.dn_bind_layer(self._h, abs_idx, ptrs))
            if getattr(self, "n_experts", 0):
                self._bind_experts(abs_idx, tensors)
            self._bound[abs_idx] = tensors  # keep the borrowed views alive
        if api_changed:
            e, n, hh = self._api.get("embed_tokens.weight"), self._api.get("norm.weight"), self._api.get("lm_head.weight")
            for t in (e, n, hh):
                if t is not None and (not t.is_cuda or t.dtype != torch.bfloat16):
                    raise ValueError("API-layer tensors must be bf16 CUDA tensors")
            _cabi.check(self._lib.dn_bind_api(self._h, e.data_ptr() if e is not None else None,
                                              n.data_ptr() if n is not None else None,
                                              hh.data_ptr() if hh is not None else None))
        return self

    # sparse MoE layers: HF / mlx_lm checkpoint names of the router and the experts (mixtral)
    EXPERT_ROUTER = "block_sparse_moe.gate.weight"
    EXPERT_FMT = "block_sparse_moe.experts.{e}.{w}.weight"       # w1 = gate, w3 = up, w2 = down

    def _bind_experts(self, abs_idx: int, tensors: Dict[str, torch.Tensor]) -> None:
        E = self.n_experts
        router = tensors.get(self.EXPERT_ROUTER)
        if router is None:
            raise ValueError(f"layer {abs_idx}: MoE model without {self.EXPERT_ROUTER}")
        tabs = []
        for w in ("w1", "w3", "w2"):
            arr = (C.c_void_p * E)()
            for e in range(E):
                t = tensors.get(self.EXPERT_FMT.format(e=e, w=w))
                if t is None:
                    raise ValueError(f"layer {abs_idx}: expert {e} has no {w}")
                if not t.is_cuda or t.dtype != torch.bfloat16 or not t.is_contiguous():
                    raise ValueError(f"layer {abs_idx} expert {e} {w}: expected a contiguous bf16 CUDA tensor")
                arr[e] = t.data_ptr()
            tabs.append(arr)
        if not router.is_cuda or router.dtype != torch.bfloat16 or not router.is_contiguous():
            raise ValueError(f"layer {abs_idx} router: expected a contiguous bf16 CUDA tensor")
        _cabi.check(self._lib.dn_bind_layer_experts(self._h, abs_idx, router.data_ptr(), tabs[0], tabs[1], tabs[2], E))

    def unload_layers(self, abs_layers: List[int]) -> None:
        """Drop the binding of the given absolute layers (reference base.py:474-486 shrinks
        the params to 1-element zeros; here the pointers are simply forgotten)."""
        for abs_idx in abs_layers:
            if abs_idx in self.abs_to_local:
                self._lib.dn_unbind_layer(self._h, int(abs_idx))
                self._bound.pop(abs_idx, None)

    def is_bound(self, abs_layer: int) -> bool:
        return bool(self._lib.dn_layer_is_bound(self._h, int(abs_layer)))

    def apply_quantization_from_config(self, model_config: Any, model_metadata: Any) -> bool:
        """Reference base.py:227-419 converts modules to mlx QuantizedLinear when the checkpoint holds
        `<path>.scales`.  Here MLX affine group-quantised tensors were already expanded to bf16 by the
        loader (utils/model.py `_fold_quantised`), so there is nothing to convert; anything else
        (a quantisation section whose tensors were NOT folded, e.g. another scheme) is refused."""
        q = (model_config or {}).get("quantization") or (model_config or {}).get("quantization_config")
        if not q:
            return False
        mode = str(q.get("mode", q.get("quant_method", "affine")) or "affine").strip().lower()
        if mode not in ("affine", "mlx", ""):
            raise NotImplementedError(f"quantisation mode {mode!r} is not supported (MLX affine group quantisation only)")
        leftovers = [k for tmap in list(model_metadata.weight_info.values()) + [model_metadata.embed_tokens, model_metadata.lm_head]
                     for k in tmap if k.endswith(("scales", "biases"))]
        if leftovers:
            raise NotImplementedError(f"quantised tensors were not expanded by the loader: {leftovers[:3]}")
        return True

    def make_cache(self, max_tokens: int = 4096) -> KVHandle:
        return KVHandle(self, max_tokens)

    # -- fused paths used by the policies ---------------------------------------------------
    def window_forward(self, layers: List[int], x: torch.Tensor, cache: KVHandle, stream=None) -> torch.Tensor:
        """x: [T, H] bf16 CUDA, updated in place, positions cache.offset.. ; does NOT advance."""
        s = _stream_ptr(stream)
        T = x.shape[0]
        arr = (C.c_int32 * len(layers))(*layers)
        base = cache.offset
        t0 = 0
        chunks = _decompose(T, self.max_chunk, self.max_prefill_chunk)
        for c in chunks:
            if len(chunks) > 1:
                cache.seek(base + t0, s)
            _cabi.check(self._lib.dn_window_forward(self._h, arr, len(layers), x[t0:t0 + c].data_ptr(), c, cache._h, s))
            t0 += c
        if len(chunks) > 1:
            cache.seek(base, s)
        return x

    def head_sample_greedy(self, x: torch.Tensor, cache: Optional[KVHandle], token_out_ptr: int, logprob_out_ptr: int,
                           stream=None) -> None:
        _cabi.check(self._lib.dn_head_sample_greedy(self._h, x.data_ptr(), x.shape[0], cache._h if cache else None,
                                                    token_out_ptr, logprob_out_ptr, _stream_ptr(stream)))

    def head_logits(self, x: torch.Tensor, want_f32: bool = True, want_bf16: bool = True, stream=None):
        f32 = torch.empty(self.vocab_size, dtype=torch.float32, device=x.device) if want_f32 else None
        b16 = torch.empty(self.vocab_size, dtype=torch.bfloat16, device=x.device) if want_bf16 else None
        _cabi.check(self._lib.dn_head_logits(self._h, x.data_ptr(), x.shape[0], f32.data_ptr() if want_f32 else None,
                                             b16.data_ptr() if want_bf16 else None, _stream_ptr(stream)))
        return f32, b16

    def destroy(self) -> None:
        if getattr(self, "_h", None):
            self._lib.dn_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
