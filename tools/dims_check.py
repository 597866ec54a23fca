"""Step kernel at other model widths (fit mode, a few layers): agreement with the per-op path and
achieved HBM bandwidth.  Llama-3-70B dims (H=8192, 64/8 heads, FFN 28672) and Qwen2.5-32B dims
(H=5120, 40/8 heads -> GQA group 5, FFN 27648, QKV bias)."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("DNET_TRANSPORT_WIRE_DTYPE", "bf16")
import torch
from dnet_b200 import _cabi
from dnet_b200.shard.models import ShardLoadModelRequest
from dnet_b200.shard.runtime import ShardRuntime
from dnet_b200.utils.model import SyntheticSource
from tests.helpers import token_message

torch.cuda.set_device(0); _cabi.init(0); lib = _cabi.load()
CFGS = {
    "llama3_70b_dims": dict(hidden_size=8192, num_attention_heads=64, num_key_value_heads=8, head_dim=128, intermediate_size=28672,
                            vocab_size=128256, num_hidden_layers=4, rms_norm_eps=1e-5, rope_theta=500000.0, model_type="llama",
                            tie_word_embeddings=False, torch_dtype="bfloat16"),
    "qwen25_32b_dims": dict(hidden_size=5120, num_attention_heads=40, num_key_value_heads=8, head_dim=128, intermediate_size=27648,
                            vocab_size=152064, num_hidden_layers=4, rms_norm_eps=1e-6, rope_theta=1000000.0, model_type="qwen2",
                            tie_word_embeddings=False, torch_dtype="bfloat16"),
}
out = {}
for name, cfg in CFGS.items():
    L = cfg["num_hidden_layers"]
    H, F, V = cfg["hidden_size"], cfg["intermediate_size"], cfg["vocab_size"]
    qd, kd = cfg["num_attention_heads"] * 128, cfg["num_key_value_heads"] * 128
    step_bytes = L * 2 * (qd * H + 2 * kd * H + H * qd + 3 * F * H) + 2 * V * H
    g = torch.Generator().manual_seed(7)
    prompt = torch.randint(0, V, (40,), generator=g).tolist()
    res = {}
    for mk in (1, 0):
        rt = ShardRuntime(f"{name}{mk}"); rt.kv_cache_config.max_tokens = 256
        rt.use_megakernel = bool(mk)
        rt.load_model_core(ShardLoadModelRequest(model_path=SyntheticSource(cfg, 0), total_layers=L, layers=list(range(L)),
                                                 window_size=L, residency_size=L, kv_bits="fp16"))
        try:
            rt.policy.process(token_message(rt, "n", prompt)); r = rt.activation_send_queue.get_nowait()
            toks = [r.token_id]
            for _ in range(8):
                rt.policy.process(token_message(rt, "n", [toks[-1]], req_logprobs=True)); r = rt.activation_send_queue.get_nowait()
                toks.append(r.token_id)
            f32, _ = rt.model.head_logits(rt._kv_by_nonce["n"].x1)
            torch.cuda.synchronize()
            err = lib.dn_step_error(rt.model._h, rt.compute_stream_ptr)
            # timing: device loop
            ns = rt._kv_by_nonce["n"]; run = list(range(L))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(5): rt.policy._graph_step(ns, ns.x1, True, run, True)
            with torch.cuda.stream(rt.compute_stream):
                e0.record()
                for _ in range(50): rt.policy._graph_step(ns, ns.x1, True, run, True)
                e1.record()
            e1.synchronize()
            ms = e0.elapsed_time(e1) / 50
            res[mk] = dict(tokens=toks, logits=f32.cpu(), ms=ms, gbs=step_bytes / ms / 1e6, err=int(err))
        finally:
            rt.unload_model_core()
    a, b = res[1]["logits"], res[0]["logits"]
    rel = float((a - b).abs().max() / b.abs().max())
    same = sum(int(x == y) for x, y in zip(res[1]["tokens"], res[0]["tokens"]))
    out[name] = dict(step_bytes=step_bytes, step_kernel_ms=res[1]["ms"], step_kernel_gbs=res[1]["gbs"], per_op_graph_ms=res[0]["ms"],
                     per_op_gbs=res[0]["gbs"], logits_rel_diff=rel, same_tokens=f"{same}/9", step_error=res[1]["err"],
                     tokens_step_kernel=res[1]["tokens"], tokens_per_op=res[0]["tokens"])
    print(name, json.dumps({k: v for k, v in out[name].items() if not k.startswith("tokens_")}), flush=True)
print(json.dumps(out))
