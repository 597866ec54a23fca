"""Per-layer repack on disk (reference src/dnet/utils/repack.py:42-217; its tests/test_utils_repack_ops.py):
directory naming, manifest, idempotence, bucket deletion -- and what this rebuild adds: a layer file whose
data region IS the LayerManager's pinned record (one sequential read), with a per-tensor fallback for files
written in another order (e.g. by mx.save_safetensors)."""
import hashlib
import json
import struct

import numpy as np
import pytest
import torch

from dnet_b200.utils import repack as rp
from dnet_b200.utils.layer_manager import LayerManager
from dnet_b200.utils.model import get_model_metadata, get_safetensor_details


def _write_safetensors(path, tensors):
    header, blob, off = {}, b"", 0
    for name, (dtype, shape, raw) in tensors.items():
        header[name] = {"dtype": dtype, "shape": list(shape), "data_offsets": [off, off + len(raw)]}
        blob += raw
        off += len(raw)
    hj = json.dumps(header).encode()
    path.write_bytes(struct.pack("<Q", len(hj)) + hj + blob)


def _checkpoint(root, layers=3, hidden=128):
    rng = np.random.default_rng(0)
    root.mkdir(parents=True, exist_ok=True)
    (root / "config.json").write_text(json.dumps({"model_type": "llama", "num_hidden_layers": layers, "hidden_size": hidden}))
    (root / "tokenizer.json").write_text("{}")
    shards = [{}, {}]
    raw = {}
    for l in range(layers):
        for sfx, shape in (("self_attn.q_proj.weight", (hidden, hidden)), ("mlp.up_proj.weight", (2 * hidden, hidden)),
                           ("input_layernorm.weight", (hidden,))):
            words = rng.integers(0, 65536, size=int(np.prod(shape)), dtype=np.uint16).tobytes()
            key = f"model.layers.{l}.{sfx}"
            raw[key] = words
            shards[l % 2][key] = ("BF16", shape, words)       # a layer's tensors are spread over two shard files
    for name, shape in (("model.embed_tokens.weight", (16, hidden)), ("model.norm.weight", (hidden,)), ("lm_head.weight", (16, hidden))):
        words = rng.integers(0, 65536, size=int(np.prod(shape)), dtype=np.uint16).tobytes()
        raw[name] = words
        shards[0][name] = ("BF16", shape, words)
    for i, s in enumerate(shards):
        _write_safetensors(root / f"model-0000{i}.safetensors", s)
    return raw


def test_naming_matches_the_reference():
    assert rp._sanitize_model_id("Qwen/Qwen2.5-32B Instruct!") == "Qwen_Qwen2.5-32B_Instruct"
    assert rp._sanitize_model_id("  ///  ") == "model"
    want = hashlib.sha1(b"2,5,7,").hexdigest()[:10]
    assert rp._hash_layers([7, 2, 5]) == want == rp._hash_layers(["5", 2, 7])
    assert rp.layer_file_name(7) == "layer_0007.safetensors"


def test_repack_layout_manifest_idempotence_and_byte_exactness(tmp_path, monkeypatch):
    src = tmp_path / "ckpt"
    raw = _checkpoint(src)
    monkeypatch.setenv("DNET_REPACK_DIR", str(tmp_path / "repacked"))
    out, did = rp.ensure_repacked_for_layers(str(src), [2, 0])
    assert did and out == tmp_path / "repacked" / rp._sanitize_model_id(str(src)) / rp._hash_layers([0, 2])
    assert sorted(p.name for p in out.glob("*.safetensors")) == ["api_layers.safetensors", "layer_0000.safetensors", "layer_0002.safetensors"]
    assert (out / "config.json").exists() and (out / "tokenizer.json").exists()          # non-weight artifacts copied
    man = json.loads((out / rp.MANIFEST).read_text())
    assert man["version"] == 1 and man["assigned_layers"] == [0, 2] and man["layers_hash"] == rp._hash_layers([0, 2])
    assert man["num_layers"] == 3 and man["api_layers_file"] == "api_layers.safetensors"
    assert man["files"] == ["layer_0000.safetensors", "layer_0002.safetensors"] and man["model_id"] == str(src)
    det = get_safetensor_details(out / "layer_0002.safetensors")
    assert set(det) == {f"model.layers.2.{s}" for s in ("self_attn.q_proj.weight", "mlp.up_proj.weight", "input_layernorm.weight")}
    for key, info in det.items():                                                        # bytes identical, dtype tag kept
        assert info.dtype == "BF16" and (out / "layer_0002.safetensors").read_bytes()[info.offset:info.offset + info.size_bytes] == raw[key]
    api = get_safetensor_details(out / "api_layers.safetensors")
    assert set(api) == {"model.embed_tokens.weight", "model.norm.weight", "lm_head.weight"}
    stamp = (out / "layer_0000.safetensors").stat().st_mtime_ns
    out2, did2 = rp.ensure_repacked_for_layers(str(src), [0, 2])
    assert out2 == out and not did2 and (out / "layer_0000.safetensors").stat().st_mtime_ns == stamp
    # the repacked bucket is itself a loadable checkpoint directory for those layers
    meta = get_model_metadata(str(out))
    assert sorted(meta.weight_info) == [0, 2]


def test_layer_record_is_one_sequential_read_and_falls_back_per_tensor(tmp_path, monkeypatch):
    src = tmp_path / "ckpt"
    raw = _checkpoint(src)
    monkeypatch.setenv("DNET_REPACK_DIR", str(tmp_path / "repacked"))
    meta = get_model_metadata(str(src))
    lm = LayerManager(meta, [1, 2], use_mxload_fastpath=True, stage_host=True)
    assert lm.repack_dir is not None and (lm.repack_dir / "layer_0001.safetensors").exists()
    rec = lm._host_record(1)
    assert lm.record_reads == {"sequential": 1, "per-tensor": 0, "checkpoint": 0}
    ref = LayerManager(meta, [1], stage_host=True)._host_record(1)                      # tensor by tensor from the shards
    assert torch.equal(rec, ref)
    views = lm.views(1, rec)
    got = views["layers.1.mlp.up_proj.weight"].contiguous().view(torch.uint8).numpy().tobytes()
    assert got == raw["model.layers.1.mlp.up_proj.weight"]
    # a file with the same tensors in another order (what mx.save_safetensors may produce): per-tensor reads
    f2 = lm.repack_dir / "layer_0002.safetensors"
    det = get_safetensor_details(f2)
    blob = f2.read_bytes()
    _write_safetensors(f2, {k: (det[k].dtype, det[k].shape, blob[det[k].offset:det[k].offset + det[k].size_bytes])
                            for k in sorted(det, reverse=True)})
    rec2 = lm._host_record(2)
    assert lm.record_reads["per-tensor"] == 1
    assert torch.equal(rec2, LayerManager(meta, [2], stage_host=True)._host_record(2))
    # a truncated file is reported and the loader falls back to the checkpoint shards
    lm.drop_host_record(1)
    f1 = lm.repack_dir / "layer_0001.safetensors"
    f1.write_bytes(f1.read_bytes()[:-100])
    assert torch.equal(lm._host_record(1), ref) and lm.record_reads["checkpoint"] == 1


def test_host_record_budget_keeps_least_recently_used_out(tmp_path, monkeypatch):
    src = tmp_path / "ckpt"
    _checkpoint(src, layers=4)
    monkeypatch.setenv("DNET_REPACK_DIR", str(tmp_path / "repacked"))
    lm = LayerManager(get_model_metadata(str(src)), [0, 1, 2, 3], use_mxload_fastpath=True, host_record_budget=2)
    for l in (0, 1, 2):
        lm._host_record(l)
    assert sorted(lm._host) == [1, 2]                     # 0 went back to disk
    lm._host_record(1); lm._host_record(3)
    assert sorted(lm._host) == [1, 3] and lm.record_reads["sequential"] == 4
    lm._host_record(0)
    assert lm.record_reads["sequential"] == 5             # re-read from its per-layer file


def test_delete_repacked_layers_three_ways(tmp_path, monkeypatch):
    base = tmp_path / "repacked"
    monkeypatch.setenv("DNET_REPACK_DIR", str(base))
    src = tmp_path / "ckpt"
    _checkpoint(src)
    out, _ = rp.ensure_repacked_for_layers(str(src), [0])
    bucket = base / rp._sanitize_model_id(str(src))
    assert rp.delete_repacked_layers(current_model_path=str(out)) == [str(bucket)] and not bucket.exists()   # via the manifest
    out, _ = rp.ensure_repacked_for_layers(str(src), [1])
    (out / rp.MANIFEST).unlink()
    assert rp.delete_repacked_layers(current_model_path=str(out)) == [str(bucket.resolve())]                 # by being inside base
    rp.ensure_repacked_for_layers(str(src), [1])
    assert rp.delete_repacked_layers(model_id=str(src)) == [str(bucket)]
    rp.ensure_repacked_for_layers(str(src), [1])
    assert rp.delete_repacked_layers(all_flag=True) == [str(base)] and not base.exists()
    assert rp.delete_repacked_layers(current_model_path=str(src)) == []
