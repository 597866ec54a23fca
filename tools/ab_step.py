"""A/B of library builds on ONE box: step-kernel time at a short context for each libdnet_b200 variant given.
usage: PROMPT=128 python tools/ab_step.py lib.so[:opt=val[,opt=val]] other.so ...   (each variant twice, interleaved)"""
import os, subprocess, sys, json

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json, pathlib
sys.path.insert(0, %(root)r)
os.environ.setdefault("DNET_TRANSPORT_WIRE_DTYPE", "bf16")
import torch
from dnet_b200 import _cabi
_cabi.LIB_PATH = pathlib.Path(%(lib)r)
import bench as B
from dnet_b200.shard.models import ShardLoadModelRequest
from dnet_b200.shard.runtime import ShardRuntime
from dnet_b200.utils.model import SyntheticSource
from tests.helpers import token_message
torch.cuda.set_device(0); _cabi.init(0); lib = _cabi.load()
for kv in %(opts)r.split(","):
    if kv:
        k, v = kv.split("="); lib.dn_set_option(k.encode(), int(v))
cfg = dict(B.LLAMA3_8B); L = cfg["num_hidden_layers"]
PROMPT = int(os.environ.get("PROMPT", "128"))
rt = ShardRuntime(0); rt.kv_cache_config.max_tokens = PROMPT + 1024
rt.load_model_core(ShardLoadModelRequest(model_path=SyntheticSource(cfg, 0), total_layers=L, layers=list(range(L)), window_size=L, residency_size=L, kv_bits="fp16"))
pol = rt.policy
g = torch.Generator().manual_seed(1234)
prompt = torch.randint(0, cfg["vocab_size"], (PROMPT,), generator=g).tolist()
pol.process(token_message(rt, "d", prompt)); first = rt.activation_send_queue.get_nowait()
ns = rt.get_or_make_kv("d"); ns.kv.set_token(first.token_id, rt.compute_stream_ptr)
run = list(range(L))
res = []
for rep in range(3):
    for _ in range(8): pol._graph_step(ns, ns.x1, True, run, True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rt.compute_stream.synchronize()
    e0.record(rt.compute_stream)
    for _ in range(64): pol._graph_step(ns, ns.x1, True, run, True)
    e1.record(rt.compute_stream); rt.compute_stream.synchronize()
    res.append(e0.elapsed_time(e1) / 64)
print("AB", json.dumps({"lib": os.path.basename(%(lib)r), "opts": %(opts)r, "prompt": PROMPT, "ms_per_step": res, "ctx_end": int(ns.kv.offset)}))
'''
libs = sys.argv[1:]
for rnd in range(2):
    for spec in libs:
        lib, _, opts = spec.partition(":")
        p = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT, "lib": os.path.abspath(lib), "opts": opts}], capture_output=True, text=True, timeout=600)
        line = [l for l in p.stdout.splitlines() if l.startswith("AB ")]
        print(line[0] if line else ("FAILED " + lib + " " + p.stderr[-800:]), flush=True)
