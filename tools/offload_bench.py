"""Layer-swap path (BASELINE config 4 shape): Llama-3-70B dims, one shard holding more layers than
HBM slots; weights live in pinned host memory and are staged per window on the prefetch stream.
Reports decode tok/s, achieved host->HBM GB/s and the raw pinned cudaMemcpyAsync bandwidth of the box."""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("DNET_TRANSPORT_WIRE_DTYPE", "bf16")
import torch
from dnet_b200 import _cabi
from dnet_b200.shard.models import ShardLoadModelRequest
from dnet_b200.shard.runtime import ShardRuntime
from dnet_b200.utils.model import SyntheticSource
from tests.helpers import token_message

LLAMA3_70B = dict(hidden_size=8192, num_attention_heads=64, num_key_value_heads=8, head_dim=128, intermediate_size=28672,
                  vocab_size=128256, num_hidden_layers=80, rms_norm_eps=1e-5, rope_theta=500000.0, model_type="llama",
                  tie_word_embeddings=False, torch_dtype="bfloat16")
torch.cuda.set_device(0); _cabi.init(0); lib = _cabi.load()
NL = int(os.environ.get("LAYERS", "8"))
layers = list(range(NL))          # a middle shard: no embed / head
layers = [l + 1 for l in layers]
layer_bytes = 2 * (2 * 8192 * 8192 + 2 * 1024 * 8192 + 3 * 28672 * 8192 + 2 * 8192)
# raw pinned -> device copy bandwidth of this box
buf_h = torch.empty(1 << 30, dtype=torch.uint8).pin_memory(); buf_d = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    buf_d.copy_(buf_h, non_blocking=True); s.synchronize()
    t0 = time.perf_counter()
    for _ in range(4): buf_d.copy_(buf_h, non_blocking=True)
    s.synchronize()
pcie = 4 * (1 << 30) / (time.perf_counter() - t0) / 1e9
del buf_h, buf_d
out = {"pinned_h2d_gbs": pcie, "layer_bytes": layer_bytes, "local_layers": NL, "runs": []}
for window, resident_windows in ((2, 1), (2, 2)):
    os.environ["DNET_TOPOLOGY_RESIDENT_WINDOWS"] = str(resident_windows)
    from dnet_b200.config import get_settings; get_settings.cache_clear()
    rt = ShardRuntime(shard_id=1); rt.kv_cache_config.max_tokens = 256
    t0 = time.perf_counter()
    rt.load_model_core(ShardLoadModelRequest(model_path=SyntheticSource(LLAMA3_70B, 0, layers=layers), total_layers=80, layers=layers,
                                             window_size=window, residency_size=window, kv_bits="fp16"))
    load_s = time.perf_counter() - t0
    assert rt.policy._mode == "offload", rt.policy._mode
    H = 8192
    from dnet_b200.core.types.messages import ActivationMessage
    x = torch.randn(1, 1, H, device="cuda").to(torch.bfloat16)
    def step():
        msg = ActivationMessage(nonce="o", pool_id=-1, batch_size=1, shape=(1, 1, H), dtype="bfloat16", layer_id=layers[0] - 1,
                                timestamp=0, node_origin="", callback_url="", tensor=x, temperature=0.0)
        rt.policy.process(msg)
        return rt.activation_send_queue.get_nowait()
    for _ in range(2): step()
    rt.compute_stream.synchronize(); torch.cuda.synchronize()
    K = 6
    t0 = time.perf_counter()
    for _ in range(K): r = step()
    rt.compute_stream.synchronize(); torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / K
    swapped = NL if rt.policy.weight_cache.max_weights < NL else 0
    run = {"window": window, "resident_windows": resident_windows, "hbm_slots": rt.policy.weight_cache.max_weights, "load_s": load_s,
           "ms_per_token": dt * 1e3, "tok_s": 1 / dt, "layers_swapped_per_token": swapped,
           "h2d_gbs": swapped * layer_bytes / dt / 1e9, "frac_of_pinned_copy_bw": swapped * layer_bytes / dt / 1e9 / pcie,
           "resident_after": sorted(rt.policy.weight_cache.cache.keys())}
    print(json.dumps(run), flush=True)
    out["runs"].append(run)
    rt.unload_model_core()
print(json.dumps(out))
