"""GPU-less tests of the host-side mirror.  Each block restates the structural pins the
reference's own tests hold for this path, mlx-free."""
import json
import struct
import threading
import time
from concurrent.futures import Future

import numpy as np
import pytest
import torch

from dnet_b200.core.memory.memory_pool import DynamicMemoryPool, LayerAwareMemoryPool
from dnet_b200.core.memory.weight_cache import WeightCache
from dnet_b200.core.types.messages import ActivationMessage, PoolStatus, TokenResult
from dnet_b200.shard.policies import (FitInMemoryPolicy, NoopPolicy, OffloadPolicy, POLICY_REGISTRY, make_policy,
                                      plan_policy)
from dnet_b200.shard.policies.base import ComputePolicy
from tests.fakes import FakeLayerManagerForCache, FakeModelMetadata, FakeRuntimeForPolicy, FakeWeightSize


# ---- WeightCache (reference tests/test_weight_cache.py:23-231) ----------------------------
def _wc(layers, window_size=1, resident_windows=2, wi=None):
    return WeightCache(layers, FakeModelMetadata(wi or {}), window_size=window_size, resident_windows=resident_windows,
                       layer_manager=FakeLayerManagerForCache(wi or {}))


def test_wc_resident_budget():
    assert _wc([0, 1, 2], window_size=1, resident_windows=2).max_weights == 2
    assert _wc([0, 1, 2], window_size=None).max_weights == 3
    assert _wc([0, 1], window_size=4, resident_windows=9999).max_weights == 2


def test_wc_hit_refcount_and_lru_touch():
    wc = _wc([5])
    data = {"w": torch.tensor([1.0])}
    wc.cache[5] = (data, 0.0)
    wc.reference_counts[5] = 0
    assert wc.get_weight(5, inc_ref=True) is data
    assert wc.reference_counts[5] == 1 and wc.cache[5][1] > 0.0
    assert wc.get_weight(5, inc_ref=False) is data and wc.reference_counts[5] == 1


def test_wc_creator_success_and_future_cleared():
    wc = _wc([7], wi={7: {"w": FakeWeightSize(4)}})
    out = wc.get_weight(7)
    assert out and "w" in out and 7 in wc.cache and wc.reference_counts[7] == 1
    assert 7 not in wc.loading_futures


def test_wc_inflight_wait():
    wc = _wc([8])
    f = Future()
    with wc.lock:
        wc.loading_futures[8] = f

    def fulfill():
        time.sleep(0.01)
        with wc.lock:
            wc.cache[8] = ({"w": torch.tensor([8.0])}, time.time())
        f.set_result(True)

    threading.Thread(target=fulfill).start()
    out = wc.get_weight(8)
    assert out is not None and float(out["w"][0]) == 8.0 and wc.reference_counts[8] == 1


def test_wc_lru_eviction_is_refcount_gated_and_recency_ordered():
    wc = _wc([0, 1, 2], window_size=1, resident_windows=2)
    wc.get_weight(0); time.sleep(0.002)
    wc.get_weight(1); time.sleep(0.002)
    assert wc.get_resident_layers() == [0, 1]
    wc.decrease_reference(0)            # 0 evictable, 1 still referenced
    wc.get_weight(2)
    assert sorted(wc.cache) == [1, 2] and wc.layer_manager._released == [0]
    # nothing evictable -> the cache overfills rather than dropping a referenced layer
    wc.get_weight(0)
    assert sorted(wc.cache) == [0, 1, 2]
    assert wc.evict_layer(1) is False
    wc.decrease_reference(1)
    assert wc.evict_layer(1) is True and wc.evict_layer(1) is True
    assert wc.evict_layers([0, 2]) == 0


def test_wc_load_failure_returns_none_and_unblocks_waiters():
    class Boom(FakeLayerManagerForCache):
        def load_layer_to_gpu(self, layer_id):
            raise ValueError("boom")

    wc = WeightCache([3], FakeModelMetadata({}), window_size=1, layer_manager=Boom())
    assert wc.get_weight(3) is None and 3 not in wc.loading_futures
    wc.shutdown()
    assert wc.get_weight(3) is None


# ---- pools (reference tests/test_memory_pool.py, test_layer_aware_memory_pool.py:19-119) ----
def test_pool_exact_size_reuse_and_release():
    p = DynamicMemoryPool(total_memory_mb=1)
    a = p.allocate(4096, torch.float16)
    assert p.get_buffer(a).numel() == 2048
    p.release(a)
    assert p.buffer_info[a].status == PoolStatus.FREE and p.get_buffer(a) is None
    b = p.allocate(4096, torch.float16)
    assert b == a                                   # exact-size free buffer reused
    c = p.allocate(4096, torch.float16)
    assert c != a and p.get_stats()["total_buffers"] == 2


def test_pool_lru_eviction_of_free_buffers_only():
    p = DynamicMemoryPool(total_memory_mb=1)
    ids = [p.allocate(256 * 1024, torch.uint8) for _ in range(4)]
    assert None not in ids and p.allocate(256 * 1024, torch.uint8) is None   # full, nothing free
    p.release(ids[1]); time.sleep(0.002); p.release(ids[0])
    big = p.allocate(300 * 1024, torch.uint8)       # needs eviction of the LRU free buffers
    assert big is not None and ids[1] not in p.buffers
    assert p.get_buffer_view(big, (10, 10)).shape == (10, 10)
    assert p.get_buffer_view(big, (1 << 20,)) is None


def test_layer_aware_pool_tracks_median_size():
    lp = LayerAwareMemoryPool(total_memory_mb=4)
    for n in (8, 8, 64):
        pid = lp.allocate_for_layer(3, (1, n, 16), torch.bfloat16)
        lp.release(pid)
    assert lp.get_typical_size(3) == 8 * 16 * 2
    assert lp.get_typical_size(9) is None
    assert 3 in lp.get_stats()["layer_stats"]


# ---- policy registry / plan (reference tests/subsystems/test_shard_policies.py:46-69) -------
class _Topo:
    resident_windows = 1


def test_registry_and_plan_truth_table():
    assert {"fit", "offload", "sliding_fit"} <= set(POLICY_REGISTRY)
    rt = FakeRuntimeForPolicy([0, 1])
    assert isinstance(make_policy("fit", rt, 1), FitInMemoryPolicy)
    assert isinstance(make_policy(" OFFLOAD ", rt, 1), OffloadPolicy)
    assert isinstance(make_policy("sliding_fit", rt, 1), OffloadPolicy)
    with pytest.raises(ValueError):
        make_policy("does_not_exist", rt, 1)
    plan = plan_policy(local_count=4, requested_w=4, residency_size=8, topology_config=_Topo())
    assert plan.mode == "fit" and plan.policy_cls is FitInMemoryPolicy and plan.window_size == 4
    assert plan.is_sliding is False and plan.resident_windows == 9999
    plan2 = plan_policy(local_count=6, requested_w=3, residency_size=5, topology_config=_Topo())
    assert plan2.mode == "offload" and plan2.policy_cls is OffloadPolicy and plan2.window_size == 3
    assert plan2.is_sliding is False and plan2.resident_windows == 1
    plan3 = plan_policy(local_count=6, requested_w=4, residency_size=1, topology_config=_Topo())
    assert plan3.mode == "offload" and plan3.is_sliding is True and plan3.window_size == 1


def test_next_local_layers_and_delta_swap():
    s = [1, 2, 3, 5]
    f = ComputePolicy._next_local_layers
    assert f(s, 0, 2) == [1, 2] and f(s, 2, 3) == [3, 5] and f(s, 5, 1) == [] and f(s, 1, 0) == []
    rt = FakeRuntimeForPolicy([0, 1, 2, 3])
    pol = NoopPolicy(rt, 1)
    pol.window_size = 2
    pol.weight_cache = _wc([0, 1, 2, 3], window_size=2, resident_windows=1)
    for l in (0, 1):
        pol.weight_cache.get_weight(l, inc_ref=False)
    pol._bound_versions = {0: 1, 1: 2}
    n = pol._delta_swap_eviction([2, 3], pol.weight_cache.get_resident_layers())
    assert n == 2 and rt.model.unloaded == [[0, 1]] and pol._bound_versions == {}
    assert pol._delta_swap_eviction([2], [2]) == 0


def test_noop_policy_and_messages():
    rt = FakeRuntimeForPolicy([])
    assert NoopPolicy(rt, 1).process(None) is None
    m = ActivationMessage(nonce="n", pool_id=1, batch_size=1, shape=(1,), dtype="tokens", layer_id=-1, timestamp=0,
                          node_origin="api", callback_url="")
    assert m.temperature == 1.0 and m.top_k == -1 and m.is_final is False and m.token_id == -1
    assert TokenResult(3).top_logprobs == {}


def test_chunk_decomposition_covers_T_with_instantiated_sizes():
    from dnet_b200.core.models.base import _decompose
    for tmax in (1, 2, 4):
        for T in range(1, 40):
            c = _decompose(T, tmax)
            assert sum(c) == T and all(x in (1, 2, 4) and x <= tmax for x in c)
    assert _decompose(7, 4) == [4, 2, 1]


# ---- safetensors parser + layer packing (reference tests/test_utils_model_io.py:62-79,
#      tests/test_layer_manager.py:172-180: BF16 words 0x3F80 -> 1.0, 0x3F00 -> 0.5, 0x4000 -> 2.0)
def _write_safetensors(path, tensors):
    header, blob, off = {}, b"", 0
    for name, (dtype, shape, raw) in tensors.items():
        header[name] = {"dtype": dtype, "shape": list(shape), "data_offsets": [off, off + len(raw)]}
        blob += raw
        off += len(raw)
    hj = json.dumps(header).encode()
    path.write_bytes(struct.pack("<Q", len(hj)) + hj + blob)


def test_safetensors_metadata_and_byte_exact_layer_record(tmp_path):
    from dnet_b200.utils.layer_manager import LayerManager
    from dnet_b200.utils.model import get_model_metadata, get_safetensor_details, load_weight

    bf = np.array([0x3F80, 0x3F00, 0x4000, 0xBF80], dtype=np.uint16).tobytes()
    f32 = np.arange(6, dtype=np.float32).tobytes()
    (tmp_path / "config.json").write_text(json.dumps({"model_type": "llama", "num_hidden_layers": 2,
                                                      "hidden_size": 2}))
    _write_safetensors(tmp_path / "model.safetensors", {
        "model.layers.1.self_attn.q_proj.weight": ("BF16", (2, 2), bf),
        "model.layers.1.input_layernorm.weight": ("F32", (6,), f32),
        "model.embed_tokens.weight": ("BF16", (2, 2), bf),
        "model.norm.weight": ("BF16", (4,), bf),
        "lm_head.weight": ("BF16", (2, 2), bf),
    })
    det = get_safetensor_details(tmp_path / "model.safetensors")
    assert det["model.norm.weight"].size_bytes == 8 and det["model.norm.weight"].dtype == "BF16"
    meta = get_model_metadata(str(tmp_path))
    assert meta.num_layers == 2 and meta.model_type == "llama" and list(meta.weight_info) == [1]
    assert set(meta.weight_info[1]) == {"self_attn.q_proj.weight", "input_layernorm.weight"}
    t = load_weight(meta.weight_info[1]["self_attn.q_proj.weight"], {})
    assert t.dtype == torch.bfloat16 and t.flatten().tolist() == [1.0, 0.5, 2.0, -1.0]
    lm = LayerManager(meta, [1], stage_host=True)
    rec = lm._host_record(1)
    views = lm.views(1, rec)
    assert views["layers.1.self_attn.q_proj.weight"].flatten().tolist() == [1.0, 0.5, 2.0, -1.0]
    assert views["layers.1.input_layernorm.weight"].tolist() == [0, 1, 2, 3, 4, 5]
    assert all(e.offset % 256 == 0 for e in lm._layout[1]) and lm.layer_bytes(1) % 256 == 0
    with pytest.raises(RuntimeError):
        lm.load_layer_to_gpu(0)
    (tmp_path / "bad.safetensors").write_bytes(b"")
    _write_safetensors(tmp_path / "bad.safetensors", {"weird.key": ("BF16", (4,), bf)})
    with pytest.raises(RuntimeError, match="Unexpected key"):
        get_model_metadata(str(tmp_path))


def test_host_dict_and_synthetic_sources_share_the_metadata_contract():
    from dnet_b200.utils.model import HostDictSource, SyntheticSource, get_model_metadata
    from oracle.llama_oracle import OracleConfig, make_weights

    cfgd = dict(hidden_size=256, num_attention_heads=2, num_key_value_heads=1, head_dim=128, intermediate_size=256,
                vocab_size=64, num_hidden_layers=2, model_type="llama")
    w = make_weights(OracleConfig.from_dict(cfgd), 1)
    m1 = get_model_metadata(HostDictSource(w, cfgd))
    m2 = get_model_metadata(SyntheticSource(cfgd, seed=0))
    assert set(m1.weight_info) == set(m2.weight_info) == {0, 1}
    for l in (0, 1):
        assert {k: v.shape for k, v in m1.weight_info[l].items()} == {k: v.shape for k, v in m2.weight_info[l].items()}
    assert m1.embed_tokens["weight"].shape == (64, 256) and m1.num_layers == 2


def test_wire_bytes_roundtrip_is_raw_little_endian():
    from dnet_b200.utils.serialization import bytes_to_tensor, tensor_to_bytes
    t = torch.tensor([1.0, 0.5, 2.0, -1.0]).to(torch.bfloat16).view(1, 2, 2)
    b = tensor_to_bytes(t)
    assert b == np.array([0x3F80, 0x3F00, 0x4000, 0xBF80], dtype="<u2").tobytes()
    assert torch.equal(bytes_to_tensor(b, "bfloat16", (1, 2, 2)), t)
    assert torch.equal(bytes_to_tensor(b, "mlx.core.bfloat16", (4,)), t.flatten())


def _mlx_quantise(w: np.ndarray, bits: int, group: int):
    """affine group quantisation with MLX's packing (test-side restatement)"""
    rows, cols = w.shape
    g = w.reshape(rows, cols // group, group).astype(np.float32)
    lo, hi = g.min(-1, keepdims=True), g.max(-1, keepdims=True)
    scale = np.maximum((hi - lo) / (2 ** bits - 1), 1e-7)
    q = np.clip(np.rint((g - lo) / scale), 0, 2 ** bits - 1).astype(np.uint32).reshape(rows, cols)
    per = 32 // bits
    packed = np.zeros((rows, cols // per), dtype=np.uint32)
    for i in range(per):
        packed |= q[:, i::per] << np.uint32(i * bits)
    return packed, scale.reshape(rows, cols // group), lo.reshape(rows, cols // group), q


@pytest.mark.parametrize("bits", [4, 8])
def test_mlx_quantised_checkpoint_dequantises_at_load(bits):
    """`X.scales` next to `X.weight` marks a quantised Linear / Embedding (reference base.py:227-234);
    the loader expands it to bf16 with w = scales * q + biases and the layer record holds bf16."""
    import torch
    from dnet_b200.utils.layer_manager import LayerManager
    from dnet_b200.utils.model import HostDictSource, get_model_metadata, load_weight

    rng = np.random.default_rng(bits)
    cfg = dict(hidden_size=128, num_attention_heads=1, num_key_value_heads=1, head_dim=128, intermediate_size=256, vocab_size=64,
               num_hidden_layers=1, rms_norm_eps=1e-5, rope_theta=10000.0, model_type="llama", tie_word_embeddings=False,
               quantization={"bits": bits, "group_size": 64})
    shapes = {"self_attn.q_proj": (128, 128), "self_attn.k_proj": (128, 128), "self_attn.v_proj": (128, 128),
              "self_attn.o_proj": (128, 128), "mlp.gate_proj": (256, 128), "mlp.up_proj": (256, 128), "mlp.down_proj": (128, 256)}
    tensors, expect = {}, {}

    def add(name, shape):
        w = rng.normal(0, 0.05, size=shape).astype(np.float32)
        packed, sc, bi, q = _mlx_quantise(w, bits, 64)
        sc16, bi16 = torch.from_numpy(sc).to(torch.bfloat16), torch.from_numpy(bi).to(torch.bfloat16)
        tensors[name + ".weight"] = torch.from_numpy(packed.view(np.int32))
        tensors[name + ".scales"], tensors[name + ".biases"] = sc16, bi16
        ref = q.reshape(shape[0], -1, 64).astype(np.float32) * sc16.float().numpy()[:, :, None] + bi16.float().numpy()[:, :, None]
        expect[name + ".weight"] = torch.from_numpy(ref.reshape(shape)).to(torch.bfloat16)

    for k, shp in shapes.items():
        add("model.layers.0." + k, shp)
    add("model.embed_tokens", (64, 128))
    add("lm_head", (64, 128))
    for k in ("model.layers.0.input_layernorm.weight", "model.layers.0.post_attention_layernorm.weight", "model.norm.weight"):
        tensors[k] = torch.ones(128, dtype=torch.bfloat16)
    meta = get_model_metadata(HostDictSource(tensors, cfg))
    assert not any(k.endswith(("scales", "biases")) for k in meta.weight_info[0])
    for suffix, info in meta.weight_info[0].items():
        if suffix.endswith("proj.weight"):
            assert info.dtype == "BF16" and info.shape == shapes[suffix[:-len(".weight")]]
            got = load_weight(info, {}, meta.source)
            assert torch.equal(got, expect["model.layers.0." + suffix])
    assert torch.equal(load_weight(meta.embed_tokens["weight"], {}, meta.source), expect["model.embed_tokens.weight"])
    assert torch.equal(load_weight(meta.lm_head["weight"], {}, meta.source), expect["lm_head.weight"])
    # the packed layer record is plain bf16
    lm = LayerManager(meta, [0])
    assert lm.layer_bytes(0) >= sum(2 * a * b for a, b in shapes.values())
    # a checkpoint with scales but an unsupported packing is rejected loudly
    bad = dict(tensors)
    bad["model.layers.0.mlp.down_proj.weight"] = torch.zeros(128, 24, dtype=torch.int32)      # 3-bit-like width
    with pytest.raises(ValueError):
        get_model_metadata(HostDictSource(bad, cfg))


def test_mixtral_checkpoint_names_flow_through_metadata_layout_and_repack(tmp_path):
    """Sparse-MoE checkpoints (HF mixtral names) are ordinary layer tensors for the loader: metadata, the pinned layer
    record (256-byte aligned offsets), the repacked per-layer file and the synthetic source agree on the tensor set."""
    from dnet_b200.utils.layer_manager import LayerManager
    from dnet_b200.utils.model import HostDictSource, SyntheticSource, get_model_metadata
    from oracle.llama_oracle import OracleConfig, make_weights

    cfgd = dict(hidden_size=256, num_attention_heads=2, num_key_value_heads=1, head_dim=128, intermediate_size=256,
                vocab_size=64, num_hidden_layers=2, model_type="mixtral", num_local_experts=4, num_experts_per_tok=2)
    w = make_weights(OracleConfig.from_dict(cfgd), 3)
    m1 = get_model_metadata(HostDictSource(w, cfgd))
    m2 = get_model_metadata(SyntheticSource(cfgd, seed=0))
    assert m1.model_type == "mixtral" and set(m1.weight_info) == set(m2.weight_info) == {0, 1}
    want = {"block_sparse_moe.gate.weight"} | {f"block_sparse_moe.experts.{e}.{k}.weight" for e in range(4) for k in ("w1", "w2", "w3")}
    for m in (m1, m2):
        suffixes = set(m.weight_info[1])
        assert want <= suffixes and "mlp.gate_proj.weight" not in suffixes
        assert {k: tuple(v.shape) for k, v in m.weight_info[1].items() if k in want} == \
               {k: ((4, 256) if k.endswith("gate.weight") else ((256, 256))) for k in want}
    lm = LayerManager(m1, [1], stage_host=True)
    rec = lm._host_record(1)
    views = lm.views(1, rec)
    for e in range(4):
        for k in ("w1", "w2", "w3"):
            name = f"block_sparse_moe.experts.{e}.{k}.weight"
            assert torch.equal(views[f"layers.1.{name}"], w[f"model.layers.1.{name}"])
    assert torch.equal(views["layers.1.block_sparse_moe.gate.weight"], w["model.layers.1.block_sparse_moe.gate.weight"])
    assert all(e.offset % 256 == 0 for e in lm._layout[1])
