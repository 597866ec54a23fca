"""safetensors header parser + ModelMetadata (reference src/dnet/utils/model.py:27-155,388-467).

Three weight sources share the ModelMetadata/TensorInfo contract:
  * a local HF-style directory of *.safetensors + config.json (the reference path);
  * ``HostDictSource``: an in-memory {name: tensor} checkpoint (parity tests feed the
    oracle's synthetic weights through it);
  * ``SyntheticSource``: shape-only random-init weights generated directly in HBM
    (bench.py: there is no network for real checkpoints).
"""
from __future__ import annotations

import glob
import json
import mmap
import re
import struct
from collections import defaultdict
from dataclasses import dataclass
from functools import cached_property
from pathlib import Path
from typing import Any, Dict, Optional, Tuple

import torch

EMBED_TOKENS_RE = re.compile(r"^model\.embed_tokens\.(.+)$")
LAYERS_RE = re.compile(r"^model\.layers\.(\d+)\.(.+)$")
LM_HEAD_RE = re.compile(r"^lm_head\.(.+)$")
NORM_RE = re.compile(r"^model\.norm\.(.+)$")


def get_model_layer_name(layer_idx: int, name: str) -> str:
    return f"layers.{layer_idx}.{name}"


def get_model_embed_tokens_name(name: str) -> str:
    return f"embed_tokens.{name}"


def get_lm_head_name(name: str) -> str:
    return f"lm_head.{name}"


def get_model_norm_name(name: str) -> str:
    return f"norm.{name}"


@dataclass(slots=True, frozen=True)
class TensorInfo:
    """The tensor information stored inside safetensor file."""

    dtype: str
    shape: Tuple[int, ...]
    size_bytes: int
    offset: int
    filename: str


@dataclass(frozen=True)
class ModelMetadata:
    """LLM model metadata"""

    path: Any
    weight_info: Dict[int, Dict[str, TensorInfo]]
    embed_tokens: Dict[str, TensorInfo]
    lm_head: Dict[str, TensorInfo]
    norm: Dict[str, TensorInfo]
    config: Any
    source: Any = None  # HostDictSource / SyntheticSource, None for files

    @cached_property
    def embedding_size(self) -> int:
        embedding_size = self.model_config.get("embedding_size")
        if embedding_size is None:
            if self.embed_tokens and "weight" in self.embed_tokens:
                embedding_size = self.embed_tokens["weight"].shape[1]
            else:
                embedding_size = self.model_config.get("hidden_size")
        if embedding_size is None:
            raise ValueError("Could not find embedding_size or hidden_size in model metadata")
        return embedding_size

    @cached_property
    def num_layers(self) -> int:
        try:
            n = int(self.config.get("num_hidden_layers"))
            if n > 0:
                return n
        except Exception:
            pass
        return max(self.weight_info.keys()) + 1

    @cached_property
    def model_type(self) -> str:
        return self.config["model_type"]

    @property
    def model_config(self) -> Any:
        return self.config


def get_safetensor_details(path) -> Dict[str, TensorInfo]:
    """8-byte LE header length + JSON header; offsets are absolute file offsets."""
    with open(path, "rb") as f:
        mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
        header_len = struct.unpack("<Q", mm[:8])[0]
        header = json.loads(mm[8:8 + header_len].decode("utf-8"))
        data_base = 8 + header_len
        details = {}
        for name, info in header.items():
            if name == "__metadata__":
                continue
            start, end = info["data_offsets"]
            details[name] = TensorInfo(dtype=info["dtype"], shape=tuple(info["shape"]), size_bytes=end - start,
                                       offset=data_base + start, filename=str(path))
        mm.close()
        return details


def _classify(details: Dict[str, TensorInfo], weight_info, embed_tokens, lm_head, norm) -> None:
    for key, val in details.items():
        if m := EMBED_TOKENS_RE.match(key):
            embed_tokens[m.group(1)] = val
        elif m := LM_HEAD_RE.match(key):
            lm_head[m.group(1)] = val
        elif m := NORM_RE.match(key):
            norm[m.group(1)] = val
        elif m := LAYERS_RE.match(key):
            layer_idx, suffix = m.groups()
            weight_info[int(layer_idx)][suffix] = val
        else:
            raise RuntimeError(f"Unexpected key {key}")


def _validate_layers(config, weight_info) -> None:
    try:
        cfg_layers = int(config.get("num_hidden_layers", -1))
    except Exception:
        cfg_layers = -1
    if cfg_layers > 0:
        bad = [i for i in weight_info if i < 0 or i >= cfg_layers]
        if bad:
            raise RuntimeError(
                f"Layer indices out of range for model (num_hidden_layers={cfg_layers}): {sorted(set(bad))}")


_ST_NAME = {torch.bfloat16: "BF16", torch.float16: "F16", torch.float32: "F32", torch.int32: "I32",
            torch.int64: "I64", torch.uint8: "U8", torch.int8: "I8"}


class HostDictSource:
    """An in-memory checkpoint {HF name: host tensor}."""

    def __init__(self, tensors: Dict[str, torch.Tensor], config: dict):
        self.tensors = tensors
        self.config = config

    def details(self) -> Dict[str, TensorInfo]:
        return {k: TensorInfo(_ST_NAME[v.dtype], tuple(v.shape), v.numel() * v.element_size(), 0, f"mem://{k}")
                for k, v in self.tensors.items()}

    def host_tensor(self, info: TensorInfo) -> torch.Tensor:
        return self.tensors[info.filename[len("mem://"):]]


class SyntheticSource:
    """Random-init weights of a given architecture, generated on the device.

    normal(0, std) from a per-tensor seeded torch.cuda generator; norm weights are 1.
    ``host_tensor`` is unavailable: tensors are materialised straight into HBM slots.
    """

    def __init__(self, config: dict, seed: int = 0, std: float = 0.02, layers=None, share_layers: bool = False):
        self.config = config
        self.seed = seed
        self.std = std
        # share_layers: every layer gets the values of the first one (host copies are generated once per tensor kind);
        # for benchmarks of data movement at sizes where generating 70B distinct parameters would dominate the run
        self.share_layers = bool(share_layers)
        self._host_cache: Dict[str, torch.Tensor] = {}
        c = config
        H, F_, V = c["hidden_size"], c["intermediate_size"], c["vocab_size"]
        hd = c.get("head_dim") or H // c["num_attention_heads"]
        qd, kd = c["num_attention_heads"] * hd, c.get("num_key_value_heads", c["num_attention_heads"]) * hd
        self._shapes: Dict[str, Tuple[int, ...]] = {}
        for l in (range(c["num_hidden_layers"]) if layers is None else layers):
            p = f"model.layers.{l}."
            self._shapes[p + "self_attn.q_proj.weight"] = (qd, H)
            self._shapes[p + "self_attn.k_proj.weight"] = (kd, H)
            self._shapes[p + "self_attn.v_proj.weight"] = (kd, H)
            self._shapes[p + "self_attn.o_proj.weight"] = (H, qd)
            if int(c.get("num_local_experts", 0) or 0) > 0:       # sparse MoE FFN (mixtral checkpoint names)
                self._shapes[p + "block_sparse_moe.gate.weight"] = (int(c["num_local_experts"]), H)
                for e in range(int(c["num_local_experts"])):
                    self._shapes[p + f"block_sparse_moe.experts.{e}.w1.weight"] = (F_, H)
                    self._shapes[p + f"block_sparse_moe.experts.{e}.w3.weight"] = (F_, H)
                    self._shapes[p + f"block_sparse_moe.experts.{e}.w2.weight"] = (H, F_)
            else:
                self._shapes[p + "mlp.gate_proj.weight"] = (F_, H)
                self._shapes[p + "mlp.up_proj.weight"] = (F_, H)
                self._shapes[p + "mlp.down_proj.weight"] = (H, F_)
            self._shapes[p + "input_layernorm.weight"] = (H,)
            self._shapes[p + "post_attention_layernorm.weight"] = (H,)
        self._shapes["model.embed_tokens.weight"] = (V, H)
        self._shapes["model.norm.weight"] = (H,)
        if not c.get("tie_word_embeddings", False):
            self._shapes["lm_head.weight"] = (V, H)
        self._names = list(self._shapes)

    def details(self) -> Dict[str, TensorInfo]:
        out = {}
        for k, shp in self._shapes.items():
            n = 1
            for s in shp:
                n *= s
            out[k] = TensorInfo("BF16", shp, n * 2, 0, f"syn://{k}")
        return out

    def fill_device(self, info: TensorInfo, dst: torch.Tensor) -> None:
        """Generate the tensor directly into ``dst`` (a bf16 cuda view of the right shape)."""
        name = info.filename[len("syn://"):]
        if name.endswith("layernorm.weight") or name == "model.norm.weight":
            dst.fill_(1.0)
            return
        g = torch.Generator(device=dst.device)
        g.manual_seed(self.seed * 1000003 + self._names.index(name))
        std = 1.0 if name == "model.embed_tokens.weight" or name.endswith("block_sparse_moe.gate.weight") else self.std
        # generate in fp32 chunks to bound temporary memory
        flat = dst.view(-1)
        step = 1 << 26
        for i in range(0, flat.numel(), step):
            n = min(step, flat.numel() - i)
            flat[i:i + n] = (torch.randn(n, generator=g, device=dst.device, dtype=torch.float32) * std).to(dst.dtype)

    def host_tensor(self, info: TensorInfo) -> torch.Tensor:
        name = info.filename[len("syn://"):]
        key = name.split(".", 3)[3] if (self.share_layers and name.startswith("model.layers.")) else None
        if key is not None and key in self._host_cache:
            return self._host_cache[key]
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        t = torch.empty(info.shape, dtype=torch.bfloat16, device=dev)
        self.fill_device(info, t)
        t = t.cpu()
        if key is not None:
            self._host_cache[key] = t
        return t


# ---------------------------------------------------------------------------------
# MLX group-quantised checkpoints (reference core/models/base.py:227-419: a Linear / Embedding is
# quantised iff `<path>.scales` exists; bits / group_size come from config["quantization"] or
# "quantization_config").  The reference turns those modules into mlx QuantizedLinear; here the
# packed words are expanded on the host, once, while the layer record is packed:
#     w[r, j] = scales[r, j // group] * q[r, j] + biases[r, j // group]      (fp32, one rounding to bf16)
# with MLX's little-endian packing (element j of a row sits in word j // (32/bits) at bit
# (j % (32/bits)) * bits).  The kernels then stream bf16 (no int8 / int4 kernels yet: the
# checkpoint loads and runs, it does not yet run faster than bf16).
# ---------------------------------------------------------------------------------
_DEQUANT: Dict[str, Tuple[TensorInfo, TensorInfo, TensorInfo, int, int]] = {}


def _quant_params(config: dict) -> Tuple[int, int]:
    q = (config or {}).get("quantization") or (config or {}).get("quantization_config") or {}
    bits = int(q.get("bits", 0) or 0)
    group = int(q.get("group_size", 0) or 0)
    return bits, group


def _fold_quantised(tmap: Dict[str, TensorInfo], config: dict, tag: str) -> None:
    """Replace every (`X.weight` packed, `X.scales`, `X.biases`) triple by one virtual bf16 `X.weight`."""
    bits_cfg, group_cfg = _quant_params(config)
    for key in [k for k in list(tmap) if k.endswith("scales")]:
        base = key[: -len("scales")]                      # "" for embed_tokens / lm_head maps, "self_attn.q_proj." for layers
        wk, bk = base + "weight", base + "biases"
        if wk not in tmap or bk not in tmap:
            raise ValueError(f"quantised tensor {tag}{base}: found scales without weight/biases")
        w, sc, bi = tmap[wk], tmap[key], tmap[bk]
        if w.dtype not in ("U32", "I32"):
            raise ValueError(f"quantised tensor {tag}{wk}: packed dtype {w.dtype} unsupported")
        # bits / group_size: the global config values when they fit the shapes, else whatever does
        # (per-path overrides, reference base.py:283-310, show up as different shapes)
        found = None
        for g in dict.fromkeys([group_cfg or 64, 64, 32, 128]):
            inf = sc.shape[-1] * g
            if inf > 0 and (32 * w.shape[-1]) % inf == 0 and (32 * w.shape[-1]) // inf in (2, 4, 8):
                cand = ((32 * w.shape[-1]) // inf, g)
                if found is None or cand[0] == bits_cfg:
                    found = cand
                if cand[0] == bits_cfg:
                    break
        if found is None:
            raise ValueError(f"quantised tensor {tag}{wk}: shapes {w.shape} / {sc.shape} fit no 2/4/8-bit packing "
                             f"with group size 32/64/128 (3/5/6-bit packing is not supported)")
        bits, group = found
        in_features = sc.shape[-1] * group
        vname = f"deq://{tag}{wk}#{len(_DEQUANT)}"
        shape = tuple(w.shape[:-1]) + (in_features,)
        n = 1
        for d in shape:
            n *= d
        _DEQUANT[vname] = (w, sc, bi, bits, group)
        tmap[wk] = TensorInfo("BF16", shape, n * 2, 0, vname)
        del tmap[key], tmap[bk]


def _fold_all_quantised(weight_info, embed_tokens, lm_head, config) -> None:
    for lid, tmap in weight_info.items():
        _fold_quantised(tmap, config, f"model.layers.{lid}.")
    _fold_quantised(embed_tokens, config, "model.embed_tokens.")
    _fold_quantised(lm_head, config, "lm_head.")


def _dequantise(vname: str, mapped_files, source) -> torch.Tensor:
    import numpy as np
    w, sc, bi, bits, group = _DEQUANT[vname]
    wq = load_weight(w, mapped_files, source).contiguous().view(torch.int32).numpy().view(np.uint32)
    s = load_weight(sc, mapped_files, source).to(torch.float32).numpy()
    b = load_weight(bi, mapped_files, source).to(torch.float32).numpy()
    per = 32 // bits
    shifts = (np.arange(per, dtype=np.uint32) * bits)[None, None, :]
    q = ((wq[..., :, None] >> shifts) & np.uint32((1 << bits) - 1)).reshape(*wq.shape[:-1], wq.shape[-1] * per)
    rows, cols = q.shape[-2], q.shape[-1]
    out = q.reshape(-1, rows, cols // group, group).astype(np.float32) * s.reshape(-1, rows, cols // group, 1) \
        + b.reshape(-1, rows, cols // group, 1)
    return torch.from_numpy(out.reshape(*q.shape)).to(torch.bfloat16)


def get_model_metadata(model_path) -> ModelMetadata:
    """Accepts a local directory (reference behaviour, minus the HF download), a
    HostDictSource or a SyntheticSource."""
    weight_info: Dict[int, Dict[str, Any]] = defaultdict(dict)
    embed_tokens, lm_head, norm = {}, {}, {}
    if isinstance(model_path, (HostDictSource, SyntheticSource)):
        src = model_path
        _classify(src.details(), weight_info, embed_tokens, lm_head, norm)
        _fold_all_quantised(weight_info, embed_tokens, lm_head, src.config)
        _validate_layers(src.config, weight_info)
        return ModelMetadata(Path("."), dict(weight_info), embed_tokens, lm_head, norm, src.config, src)
    path = Path(model_path)
    if not (path.exists() and path.is_dir()):
        raise FileNotFoundError(f"model path {model_path!r} is not a local directory (no network: HF download unsupported)")
    with open(path / "config.json", "r") as f:
        config = json.load(f)
    for weight in sorted(glob.glob(str(path / "*.safetensors"))):
        _classify(get_safetensor_details(weight), weight_info, embed_tokens, lm_head, norm)
    _fold_all_quantised(weight_info, embed_tokens, lm_head, config)
    _validate_layers(config, weight_info)
    return ModelMetadata(path, dict(weight_info), embed_tokens, lm_head, norm, config)


class MappedFile:
    """Read-only mmap of a weight file (reference utils/model.py:215-239)."""

    def __init__(self, file_path: str):
        self.file_path = file_path
        self.file = open(file_path, "rb")
        self.mmap = mmap.mmap(self.file.fileno(), 0, access=mmap.ACCESS_READ)

    def close(self):
        try:
            self.mmap.close()
        finally:
            self.file.close()


def load_weight(wt: TensorInfo, mapped_files: Dict[str, MappedFile], source=None) -> torch.Tensor:
    """Byte-exact host tensor for one checkpoint entry (reference utils/model.py:242-260).
    BF16 stays BF16 bit-for-bit (the reference round-trips uint16<<16 -> fp32 -> bf16,
    which is the identity on the bit pattern)."""
    from .serialization import safetensor_torch_dtype

    if wt.filename.startswith("deq://"):
        return _dequantise(wt.filename, mapped_files, source)
    if wt.filename.startswith(("mem://", "syn://")):
        if source is None:
            raise ValueError("in-memory tensor needs its source")
        return source.host_tensor(wt)
    if wt.filename not in mapped_files:
        mapped_files[wt.filename] = MappedFile(wt.filename)
    mv = memoryview(mapped_files[wt.filename].mmap)[wt.offset:wt.offset + wt.size_bytes]
    td = safetensor_torch_dtype[wt.dtype]
    return torch.frombuffer(mv, dtype=torch.uint8).view(td).reshape(tuple(wt.shape))
