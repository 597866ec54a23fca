"""Per-layer repacking of a checkpoint (row a17 / N1; on-disk format of the reference's
src/dnet/utils/repack.py:42-217, which its offload / sliding_fit policies load layer by layer).

Directory layout and manifest are the reference's, so buckets written by either side are usable by
the other:

    <DNET_REPACK_DIR | ~/.dria/dnet/repacked_layers>/<sanitised model id>/<sha1("l0,l1,...")[:10]>/
        layer_0007.safetensors      keys  model.layers.7.<suffix>
        api_layers.safetensors      keys  model.embed_tokens.* / model.norm.* / lm_head.*
        repack-manifest.json        {version, model_id, source_path, assigned_layers, layers_hash,
                                     num_layers, created_at, api_layers_file, files}
        + the checkpoint's tokenizer / config files (everything that is not a weight file)

What is specific here: the safetensors files are written by a small writer of our own (the reference
calls mx.save_safetensors) with the tensors of a layer in the order -- and, when every tensor size is
a multiple of 256 bytes, at exactly the offsets -- of the pinned *layer record* the LayerManager
stages to HBM, so a swapped-in layer is ONE sequential ``readinto`` from disk into pinned memory
followed by ONE cudaMemcpyAsync (``read_layer_record``); any other safetensors file of the same layer
(written by mx.save_safetensors, different order) is read tensor by tensor.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import struct
import time
from pathlib import Path
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch

from .model import MappedFile, ModelMetadata, TensorInfo, get_model_metadata, get_safetensor_details, load_weight
from .serialization import safetensor_torch_dtype

_TORCH_TO_ST = {torch.float32: "F32", torch.float16: "F16", torch.bfloat16: "BF16", torch.int32: "I32", torch.int64: "I64",
                torch.uint8: "U8", torch.int8: "I8"}
_WEIGHT_EXTS = {".safetensors", ".bin", ".pt", ".pth", ".ckpt", ".npz", ".gguf", ".onnx"}
MANIFEST = "repack-manifest.json"


# ---------------------------------------------------------------------------- naming
def _get_repack_base_dir() -> Path:
    """DNET_REPACK_DIR wins (so tests and operators can redirect it), else the reference's default."""
    env = os.getenv("DNET_REPACK_DIR")
    return Path(env).expanduser() if env else Path.home() / ".dria" / "dnet" / "repacked_layers"


def _sanitize_model_id(model_id: str) -> str:
    cleaned = "".join(ch if (ch.isalnum() or ch in "-.") else "_" for ch in str(model_id).strip().replace("\\", "/"))
    return cleaned.strip("_") or "model"


def _hash_layers(layers: Iterable[int]) -> str:
    return hashlib.sha1("".join(f"{v}," for v in sorted(int(i) for i in layers)).encode("utf-8")).hexdigest()[:10]


def layer_file_name(layer: int) -> str:
    return f"layer_{int(layer):04d}.safetensors"


# ---------------------------------------------------------------------------- safetensors writer
def save_safetensors(path: Path, tensors: Sequence[Tuple[str, torch.Tensor, Optional[str]]], metadata: Optional[dict] = None) -> None:
    """``tensors``: (name, host tensor, safetensors dtype tag or None) in file order.  Layout per the
    safetensors spec: u64 LE header length, JSON header (padded with spaces to 8 bytes), then the
    tensors back to back in the given order."""
    header: Dict[str, dict] = {}
    if metadata:
        header["__metadata__"] = {str(k): str(v) for k, v in metadata.items()}
    off = 0
    blobs = []
    for name, t, tag in tensors:
        t = t.detach().contiguous()
        raw = t.view(torch.uint8).reshape(-1) if t.dtype != torch.uint8 else t.reshape(-1)
        header[name] = {"dtype": tag or _TORCH_TO_ST[t.dtype], "shape": list(t.shape), "data_offsets": [off, off + raw.numel()]}
        off += raw.numel()
        blobs.append(raw)
    hjson = json.dumps(header, separators=(",", ":")).encode("utf-8")
    hjson += b" " * ((-len(hjson)) % 8)
    tmp = Path(str(path) + ".partial")
    with open(tmp, "wb") as f:
        f.write(struct.pack("<Q", len(hjson)))
        f.write(hjson)
        for raw in blobs:
            f.write(memoryview(raw.numpy()))
    os.replace(tmp, path)      # a reader never sees a half-written layer


# ---------------------------------------------------------------------------- repack
def _copy_non_weight_artifacts(src_root: Path, dst_root: Path) -> None:
    dst_root.mkdir(parents=True, exist_ok=True)
    try:
        for entry in src_root.iterdir():
            if entry.is_file() and entry.suffix.lower() not in _WEIGHT_EXTS and not (dst_root / entry.name).exists():
                try:
                    shutil.copy2(entry, dst_root / entry.name)
                except OSError:
                    pass
    except OSError:
        pass


def _manifest(model_id, source_path, assigned, num_layers, out_root: Path, api_file: Optional[str]) -> dict:
    layers = sorted(set(int(i) for i in assigned))
    return {"version": 1, "model_id": str(model_id), "source_path": str(source_path), "assigned_layers": layers,
            "layers_hash": _hash_layers(layers), "num_layers": int(num_layers), "created_at": int(time.time()),
            "api_layers_file": api_file,
            "files": [layer_file_name(i) for i in layers if (out_root / layer_file_name(i)).exists()]}


def repack_per_layer(model_path, assigned_layers: List[int], out_root: Path, md: Optional[ModelMetadata] = None) -> None:
    """Write one safetensors file per assigned layer (+ api_layers.safetensors + manifest) under out_root.
    Existing files are kept (the reference's idempotence)."""
    md = md or get_model_metadata(model_path)
    out_root = Path(out_root)
    out_root.mkdir(parents=True, exist_ok=True)
    if isinstance(md.path, Path) and md.path.is_dir() and md.source is None:
        _copy_non_weight_artifacts(md.path, out_root)
    mapped: Dict[str, MappedFile] = {}
    try:
        for lid in sorted(set(int(i) for i in assigned_layers)):
            info = md.weight_info.get(lid, {})
            target = out_root / layer_file_name(lid)
            if not info or target.exists():
                continue
            # record order = sorted suffixes (LayerManager._build_layout), so the file IS the layer record
            save_safetensors(target, [(f"model.layers.{lid}.{sfx}", load_weight(info[sfx], mapped, md.source), _tag(info[sfx]))
                                      for sfx in sorted(info)], {"format": "pt", "dnet_layer": lid})
        api = []
        for prefix, group in (("model.embed_tokens.", md.embed_tokens), ("model.norm.", md.norm), ("lm_head.", md.lm_head)):
            for k in sorted(group):
                try:
                    api.append((prefix + k, load_weight(group[k], mapped, md.source), _tag(group[k])))
                except Exception:
                    api = []
                    break
        api_file = None
        if api:
            api_file = "api_layers.safetensors"
            if not (out_root / api_file).exists():
                save_safetensors(out_root / api_file, api, {"format": "pt"})
        (out_root / MANIFEST).write_text(json.dumps(
            _manifest(model_path if isinstance(model_path, (str, Path)) else "in-memory", md.path, assigned_layers,
                      md.num_layers, out_root, api_file), indent=2))
    finally:
        for mf in mapped.values():
            try:
                mf.close()
            except Exception:
                pass


def _tag(wt: TensorInfo) -> Optional[str]:
    return wt.dtype if wt.dtype in safetensor_torch_dtype and not wt.filename.startswith("deq://") else None


def ensure_repacked_for_layers(model_id, assigned_layers: List[int], md: Optional[ModelMetadata] = None) -> Tuple[Path, bool]:
    """Deterministic bucket for (model, assignment); repack on first use.  Returns (bucket, did_repack)."""
    layers = sorted(set(int(i) for i in assigned_layers))
    out_root = _get_repack_base_dir() / _sanitize_model_id(model_id if isinstance(model_id, (str, Path)) else "in-memory") / _hash_layers(layers)
    did = False
    if not (out_root / layer_file_name(layers[0])).exists():       # the reference's quick existence check
        repack_per_layer(model_id, layers, out_root, md=md)
        did = True
    man = out_root / MANIFEST
    if not man.exists():
        try:
            md = md or get_model_metadata(model_id)
            api = "api_layers.safetensors" if (out_root / "api_layers.safetensors").exists() else None
            m = _manifest(model_id, md.path, layers, md.num_layers, out_root, api)
            m["files"] = [p.name for p in sorted(out_root.glob("layer_*.safetensors"))]
            man.write_text(json.dumps(m, indent=2))
        except Exception:
            pass
    return out_root, did


def delete_repacked_layers(*, model_id: Optional[str] = None, all_flag: bool = False, base_dir=None,
                           current_model_path: Optional[str] = None) -> List[str]:
    """Remove repack buckets: everything (all_flag), one model's bucket (model_id), or the bucket that
    ``current_model_path`` belongs to -- resolved through its manifest's model_id, else by being inside the
    base directory, else by sanitising the path string (the reference's three cases)."""
    base = Path(base_dir).expanduser() if base_dir is not None else _get_repack_base_dir()
    removed: List[str] = []

    def rm(target: Path) -> bool:
        if target.exists():
            shutil.rmtree(target, ignore_errors=True)
            removed.append(str(target))
            return True
        return False

    if all_flag:
        rm(base)
        return removed
    if model_id:
        rm(base / _sanitize_model_id(str(model_id)))
        return removed
    if not current_model_path:
        return removed
    cur = Path(current_model_path)
    try:
        man = cur / MANIFEST
        if man.exists():
            src_id = json.loads(man.read_text()).get("model_id")
            if src_id and rm(base / _sanitize_model_id(str(src_id))):
                return removed
    except Exception:
        pass
    try:
        parts = cur.resolve().relative_to(base.resolve()).parts
        if parts and rm(base.resolve() / parts[0]):
            return removed
    except Exception:
        pass
    rm(base / _sanitize_model_id(str(current_model_path)))
    return removed


# ---------------------------------------------------------------------------- loader fast path
def read_layer_record(path: Path, layout, record: torch.Tensor) -> str:
    """Fill the pinned layer ``record`` (uint8, laid out by ``layout`` = LayerManager's PackedEntry list)
    from a per-layer safetensors file.  Returns how it was read:
      "sequential"  the file's data region has the record's exact order and offsets: one readinto
      "per-tensor"  anything else (e.g. written by mx.save_safetensors): one pread per tensor"""
    details = get_safetensor_details(path)
    by_suffix = {k.split(".", 3)[3] if k.startswith("model.layers.") else k: v for k, v in details.items()}
    missing = [e.suffix for e in layout if e.suffix not in by_suffix]
    if missing:
        raise KeyError(f"{path} lacks {missing}")
    base = min(v.offset for v in details.values())
    exact = all(by_suffix[e.suffix].offset - base == e.offset and by_suffix[e.suffix].size_bytes == e.nbytes for e in layout)
    buf = memoryview(record.numpy())
    with open(path, "rb", buffering=0) as f:
        if exact:
            end = max(e.offset + e.nbytes for e in layout)
            f.seek(base)
            got = 0
            while got < end:
                n = f.readinto(buf[got:end])
                if not n:
                    raise EOFError(f"{path} ended after {got} of {end} data bytes")
                got += n
            return "sequential"
        for e in layout:
            wt = by_suffix[e.suffix]
            if wt.size_bytes != e.nbytes:
                raise ValueError(f"{path}: {e.suffix} has {wt.size_bytes} bytes, the record expects {e.nbytes}")
            got = 0
            while got < e.nbytes:
                n = os.preadv(f.fileno(), [buf[e.offset + got:e.offset + e.nbytes]], wt.offset + got)
                if not n:
                    raise EOFError(f"{path}: short read of {e.suffix}")
                got += n
    return "per-tensor"
