"""Per-phase timing of k_shard_step from in-kernel globaltimer stamps (option mk_debug)."""
import ctypes as C, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("DNET_TRANSPORT_WIRE_DTYPE", "bf16")
import numpy as np, torch
import bench as B
from dnet_b200 import _cabi
from dnet_b200.shard.models import ShardLoadModelRequest
from dnet_b200.shard.runtime import ShardRuntime
from dnet_b200.utils.model import SyntheticSource
from tests.helpers import token_message

torch.cuda.set_device(0); _cabi.init(0); lib = _cabi.load()
cfg = dict(B.LLAMA3_8B); L = cfg["num_hidden_layers"]
PROMPT = int(os.environ.get("PROMPT", "128"))
rt = ShardRuntime(0); rt.kv_cache_config.max_tokens = PROMPT + 64
rt.load_model_core(ShardLoadModelRequest(model_path=SyntheticSource(cfg, 0), total_layers=L, layers=list(range(L)), window_size=L, residency_size=L, kv_bits="fp16"))
lib.dn_set_option(b"pf_depth", int(os.environ.get("PF", "0")))
lib.dn_set_option(b"inflight", int(os.environ.get("INFLIGHT", "0")))
if "ATTN_TC" in os.environ:
    lib.dn_set_option(b"attn_tc", int(os.environ["ATTN_TC"]))
pol = rt.policy
g = torch.Generator().manual_seed(1234)
prompt = torch.randint(0, cfg["vocab_size"], (PROMPT,), generator=g).tolist()
pol.process(token_message(rt, "d", prompt)); first = rt.activation_send_queue.get_nowait()
ns = rt.get_or_make_kv("d"); ns.kv.set_token(first.token_id, rt.compute_stream_ptr)
run = list(range(L))
if int(os.environ.get("CALIB", "0")):
    from dnet_b200.shard.calibrate import calibrate
    b = calibrate(rt)
    print("calibrated; gu rows per SM min/max", int(np.diff(b[2]).min()), int(np.diff(b[2]).max()))
for _ in range(5): pol._graph_step(ns, ns.x1, True, run, True)
lib.dn_set_option(b"mk_debug", 1)
pol._graph_step(ns, ns.x1, True, run, True)
rt.compute_stream.synchronize()
sms = lib.dn_device_sm_count()
W = 32
buf = (C.c_uint64 * (sms * L * W))()
n = lib.dn_step_debug(rt.model._h, buf, sms * L * W, rt.compute_stream_ptr)
a = np.frombuffer(buf, dtype=np.uint64).reshape(sms, L, W).astype(np.int64)
names = ["stage_norm1", "consume_qkv", "bar1", "attention", "bar2", "merge_stage", "consume_o", "bar3", "stage_norm2", "consume_gu", "bar4", "stage_act", "consume_down", "bar5"]
d = np.diff(a[:, :, :15], axis=2)          # [sm][layer][14]
mid = d[:, 4:28, :]
print("per-phase ns: mean over SMs&layers | max over SMs (mean over layers) | min over SMs")
tot = 0
for i, nm in enumerate(names):
    print(f"{nm:14s} {mid[:, :, i].mean():8.0f} {mid[:, :, i].max(axis=0).mean():8.0f} {mid[:, :, i].min(axis=0).mean():8.0f}")
    tot += mid[:, :, i].mean()
print("sum per layer", tot, "layer wall (CTA0)", np.diff(a[0, :, 0]).mean())

gu = d[:, 4:28, 9]                       # consume_gu [sm][layer]
per = gu.mean(axis=1)
order = np.argsort(per)
print("consume_gu per CTA (mean over layers): fastest", [(int(i), int(per[i])) for i in order[:6]], "slowest", [(int(i), int(per[i])) for i in order[-6:]])
print("std over layers within a CTA (mean)", gu.std(axis=1).mean(), " std across CTAs of the per-CTA mean", per.std())
dn = d[:, 4:28, 12].mean(axis=1)
print("corr(consume_gu, consume_down) across CTAs", float(np.corrcoef(per, dn)[0, 1]))

pb = a[:, 4:28, 16:20]; cwt = a[:, 4:28, 24:28]
print("producer-0 blocked on a full ring, ns per layer while producing [qkv, o, gu, down] (mean over SMs):", pb.mean(axis=(0, 1)).round(0).tolist(), "sum", float(pb.sum(axis=2).mean()))
print("consumer warp 0 waiting for weights (ring empty), ns per layer in [qkv, o, gu, down]:", cwt.mean(axis=(0, 1)).round(0).tolist(), "sum", float(cwt.sum(axis=2).mean()))
