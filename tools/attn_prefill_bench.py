"""Prefill throughput with the tcgen05 attention kernel vs the CUDA-core one (Llama-3-8B dims)."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("DNET_TRANSPORT_WIRE_DTYPE", "bf16")
import torch
import bench as B
from dnet_b200 import _cabi
from dnet_b200.shard.models import ShardLoadModelRequest
from dnet_b200.shard.runtime import ShardRuntime
from dnet_b200.utils.model import SyntheticSource
from tests.helpers import token_message

torch.cuda.set_device(0); _cabi.init(0); lib = _cabi.load()
lib.dn_set_option(b"gemm_bn256", int(os.environ.get("BN256", "1")))
cfg = dict(B.LLAMA3_8B); L = int(os.environ.get("LAYERS", cfg["num_hidden_layers"])); cfg["num_hidden_layers"] = L
NMAX = int(os.environ.get("NMAX", "8192"))
rt = ShardRuntime(0); rt.kv_cache_config.max_tokens = NMAX + 64
os.environ["DNET_KV_POOL_PAGES"] = str((NMAX // 64 + 2) * 2)
from dnet_b200.config import get_settings; get_settings.cache_clear()
rt.load_model_core(ShardLoadModelRequest(model_path=SyntheticSource(cfg, 0), total_layers=L, layers=list(range(L)), window_size=L, residency_size=L, kv_bits="fp16"))
g = torch.Generator().manual_seed(1234)
res = {}
for n in (128, 2048, NMAX):
    prompt = torch.randint(0, cfg["vocab_size"], (n,), generator=g).tolist()
    for tc in (1, 0):
        if tc == 0 and n > 2048 and not int(os.environ.get("SLOW", "0")):
            continue
        lib.dn_set_option(b"tc_attn", tc)
        for rep in range(2):
            nonce = f"p{tc}_{n}_{rep}"
            rt.get_or_make_kv(nonce).x_view(n)
            msg = token_message(rt, nonce, prompt)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            rt.policy.process(msg); out = rt.activation_send_queue.get_nowait()
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            rt.release_nonce(nonce)
        res[f"tc_attn{tc}_T{n}"] = {"ms": dt * 1e3, "tok_s": n / dt, "token": out.token_id}
        print(f"tc_attn={tc} T={n}: {dt*1e3:.2f} ms  {n/dt:.0f} tok/s  first token {out.token_id}", flush=True)
lib.dn_set_option(b"tc_attn", 1)
print(json.dumps(res))
