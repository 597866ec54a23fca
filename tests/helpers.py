"""Shared test plumbing: golden fixtures, oracle weights -> ShardRuntime, ring driving."""
from __future__ import annotations

import json
import queue
from pathlib import Path
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

GOLDEN = Path(__file__).resolve().parent / "golden"


def load_golden(name: str) -> dict:
    z = np.load(GOLDEN / f"{name}.npz", allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d["config"] = json.loads(str(d["config"]))
    d["wseed"], d["pseed"], d["steps"] = int(d["wseed"]), int(d["pseed"]), int(d["steps"])
    return d


def oracle_weights(cfgd: dict, wseed: int, layers=None):
    from oracle.llama_oracle import OracleConfig, make_weights

    return make_weights(OracleConfig.from_dict(cfgd), wseed, layers=layers)


def make_runtime(cfgd: dict, weights: Dict[str, torch.Tensor], layers: Sequence[int], *, shard_id="s0",
                 window_size: Optional[int] = None, residency_size: Optional[int] = None, cuda_graphs: bool = True,
                 max_tokens: int = 512, megakernel: bool = True, kv_bits: str = "fp16"):
    """A ShardRuntime loaded through the reference-facing path (load_model_core)."""
    from dnet_b200.shard.models import ShardLoadModelRequest
    from dnet_b200.shard.runtime import ShardRuntime
    from dnet_b200.utils.model import HostDictSource

    rt = ShardRuntime(shard_id=shard_id, queue_size=64)
    rt.kv_cache_config.max_tokens = max_tokens
    rt.use_cuda_graphs = cuda_graphs
    rt.use_megakernel = megakernel
    n = len(layers)
    req = ShardLoadModelRequest(model_path=HostDictSource(weights, cfgd), total_layers=cfgd["num_hidden_layers"],
                                layers=list(layers), window_size=window_size or n,
                                residency_size=residency_size or (window_size or n), kv_bits=kv_bits)
    rt.load_model_core(req)
    return rt


def token_message(rt, nonce: str, ids: Sequence[int], **kw):
    """What ActivationCodec.deserialize produces for a "tokens" frame."""
    from dnet_b200.shard.codec import ActivationCodec

    return ActivationCodec(rt).tokens_message(nonce, ids, **kw)


def forward_message(msg, use_bytes_rt=None):
    """Turn a shard's emitted activation into the next shard's input.  With ``use_bytes_rt``
    the tensor goes through wire bytes + the receiver's pinned pool (the gRPC path),
    otherwise the device tensor is handed over (the NVLink hop path)."""
    from dnet_b200.core.types.messages import ActivationMessage

    if use_bytes_rt is None:
        return msg
    from dnet_b200.utils.serialization import tensor_to_bytes

    rt = use_bytes_rt
    if msg.ready_event is not None:
        msg.ready_event.synchronize()
    data = tensor_to_bytes(msg.tensor)
    shape = tuple(msg.shape)
    n = int(np.prod(shape))
    pid = rt.input_pool.allocate_for_layer(layer_id=msg.layer_id, dtype=torch.bfloat16, shape=shape)
    buf = rt.input_pool.get_buffer(pid)
    buf[:n] = torch.frombuffer(bytearray(data), dtype=torch.uint8).view(torch.bfloat16)
    return ActivationMessage(nonce=msg.nonce, pool_id=pid, batch_size=msg.batch_size, shape=shape, dtype="bfloat16",
                             layer_id=msg.layer_id, timestamp=0, node_origin=msg.node_origin,
                             callback_url=msg.callback_url, req_logprobs=msg.req_logprobs,
                             req_top_logprobs=msg.req_top_logprobs, temperature=msg.temperature, top_p=msg.top_p,
                             top_k=msg.top_k, min_p=msg.min_p)


def ring_generate(rts: List, nonce: str, prompt: Sequence[int], steps: int, *, via_bytes: bool = False,
                  req_logprobs: bool = True, collect_logits=None, **kw):
    """Drive shards like InferenceManager.generate_stream does (reference api/inference.py:135-212):
    prompt as one tokens message, then one token per step; returns [(token, logprob)]."""
    out = []
    ids = list(prompt)
    for _ in range(steps):
        msg = token_message(rts[0], nonce, ids, req_logprobs=req_logprobs, **dict(kw))
        for i, rt in enumerate(rts):
            rt.policy.process(msg)
            try:
                res = rt.activation_send_queue.get_nowait()
            except queue.Empty:
                raise AssertionError(f"shard {i} emitted nothing (see log)")
            if res.is_final:
                break
            if collect_logits is not None and i == len(rts) - 1:
                pass
            msg = forward_message(res, rts[i + 1] if via_bytes else None)
        assert res.is_final, "ring ended without a final token"
        out.append((res.token_id, res.logprob, res.top_logprobs))
        ids = [res.token_id]
    return out


def rel_inf(a: torch.Tensor, b: torch.Tensor) -> float:
    """max|a-b| / max|b|  -- the relative error the 1e-3 logit tolerance is stated in."""
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())
