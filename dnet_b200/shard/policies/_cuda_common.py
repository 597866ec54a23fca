"""Pieces shared by the CUDA fit / offload policies: staging the message payload into the
nonce's HBM activation buffer, end-shard sampling and building the output message
(reference fit_in_memory.py:54-75,134-200 and offload.py:143-168,342-392)."""
from __future__ import annotations

from typing import Any, List, Optional, Tuple

import numpy as np
import torch

from dnet_b200 import _cabi
from dnet_b200.core.decoding.config import DecodingConfig
from dnet_b200.core.decoding.sampler import Sampler
from dnet_b200.core.types.messages import ActivationMessage, TokenResult
from dnet_b200.utils.serialization import canonical_dtype
from dnet_b200.utils.time import utc_epoch_now


def model_ready(rt) -> bool:
    return bool(rt.model and rt.model_metadata and rt.policy.weight_cache and rt.input_pool and rt.output_pool)


def msg_tokens(rt, msg: ActivationMessage) -> int:
    """Number of positions carried by a message (tokens: payload length, codec.py:57-58)."""
    H = rt.model.hidden_size
    if msg.dtype != "tokens" and msg.tensor is not None:
        return int(msg.tensor.numel() // H)
    n = int(np.prod(msg.shape))
    return n if msg.dtype == "tokens" else n // H


def local_run(rt, current_layer: int) -> List[int]:
    """The contiguous run of local layers starting at current_layer."""
    run: List[int] = []
    lyr = current_layer
    while lyr in rt._assigned_set:
        run.append(lyr)
        lyr += 1
    return run


def more_chunks_follow(msg: ActivationMessage) -> bool:
    """Chunked prefill: a prompt sent as several ``tokens`` frames.  Every frame but
    the last carries ``Activation.batch_size == 0`` (the reference always sends 1 and never reads the field on
    a shard, api/strategies/ring.py:74-140); the flag travels with the activation round the ring
    (build_output copies it), and the finalising shard samples only after the chunk that has it clear."""
    return int(getattr(msg, "batch_size", 1)) == 0


def note_lane(rt, msg: ActivationMessage, ns) -> None:
    """Remember which hop lane the nonce rides (device-hop transport; set by the adapter's ingress)."""
    if msg.lane >= 0 and ns.lane != msg.lane:
        ns.lane = msg.lane
        rt.lane_nonce[msg.lane] = msg.nonce
    if msg.lane >= 0:
        ns.params = {"callback_url": msg.callback_url, "logprobs": bool(msg.req_logprobs), "seq0": msg.seq0}


def _input_copy_event(rt):
    """Event behind the cudaMemcpyAsync that reads a pinned input-pool buffer: the buffer is released
    (and may be refilled by codec.deserialize for another request) only after the copy has run."""
    import ctypes as C

    lib = _cabi.load()
    ev = C.c_void_p()
    _cabi.check(lib.dn_event_create(C.byref(ev), 0))
    _cabi.check(lib.dn_event_record(ev.value, rt.compute_stream_ptr))
    return ev.value


def finish_input(rt, msg: ActivationMessage, ns) -> None:
    """Release the message's input buffer -- deferred behind its H2D copy when one is queued."""
    ev = getattr(ns, "input_copy_event", None) if ns is not None else None
    if ns is not None:
        ns.input_copy_event = None
    rt.release_input(msg.pool_id, ev)


def stage_input(rt, msg: ActivationMessage, ns) -> Optional[Tuple[torch.Tensor, int, bool]]:
    """Returns (x [T,H] bf16 device view of the nonce's activation buffer, T, is_tokens).

    tokens : pinned int32 ids -> HBM (cudaMemcpyAsync) -> embed kernel -> cast to wire dtype
    tensor : device tensor handed over by the NVLink hop (msg.tensor; with ``hop_wait`` it is this
             shard's own bulk slot and the stream first waits for the sender's sequence flag) or
             pinned wire bytes from the pool -> HBM
    """
    lib = _cabi.load()
    s = rt.compute_stream_ptr
    model = rt.model
    H = model.hidden_size
    ns.input_copy_event = None
    if ns.hop_sent_event is not None:      # the previous result of this nonce may still be leaving over NVLink
        rt.compute_stream.wait_event(ns.hop_sent_event)
        ns.hop_sent_event = None
    if msg.tensor is not None and msg.dtype != "tokens":
        src = msg.tensor.reshape(-1, H)
        T = src.shape[0]
        if msg.hop_wait is not None:
            flag, seq, consumed = msg.hop_wait
            err = rt.hop.rx_bulk.err_flag if rt.hop is not None else None
            _cabi.check(lib.dn_hop_wait(flag, seq, 20000, err, s))
            x = ns.x_view(T)
            _cabi.check(lib.dn_hop_send(x.data_ptr(), src.data_ptr(), T * H * 2, consumed, seq, s))   # copy out + credit
            return x, T, False
        if msg.ready_event is not None:
            rt.compute_stream.wait_event(msg.ready_event)
        else:   # unknown producer stream: order after everything enqueued on the current stream
            rt.compute_stream.wait_stream(torch.cuda.current_stream())
        x = ns.x_view(T)
        if src.data_ptr() != x.data_ptr():
            if src.dtype != torch.bfloat16:
                raise ValueError(f"activation dtype {src.dtype} != wire dtype bfloat16")
            with torch.cuda.stream(rt.compute_stream):
                x.copy_(src, non_blocking=True)
        return x, T, False
    input_buffer = rt.input_pool.get_buffer(msg.pool_id)
    if input_buffer is None:
        return None
    input_size = int(np.prod(msg.shape))
    if msg.dtype == "tokens":
        T = input_size
        ids_dev = ns.ids_view(T)
        if T == 1:
            ns.kv.set_token(int(input_buffer[0]), s)  # the id rides in the kernel argument
            ids_ptr = ns.kv.token_ptr
        else:
            src = input_buffer[:T]
            if src.dtype != torch.int32:
                src = src.to(torch.int32)
            _cabi.check(lib.dn_memcpy_h2d(ids_dev.data_ptr(), src.data_ptr(), T * 4, s))
            ns.keepalive = src
            ns.input_copy_event = _input_copy_event(rt)
            ids_ptr = ids_dev.data_ptr()
        x = ns.x_view(T)
        _cabi.check(lib.dn_embed(model._h, ids_ptr, T, x.data_ptr(), s))
        return x, T, True
    # raw activation bytes staged in the (pinned) pool
    if canonical_dtype(msg.dtype) != "bfloat16":
        raise ValueError(f"activation dtype {msg.dtype} != wire dtype bfloat16 (set DNET_TRANSPORT_WIRE_DTYPE=bf16)")
    T = input_size // H
    x = ns.x_view(T)
    src = input_buffer[:input_size]
    _cabi.check(lib.dn_memcpy_h2d(x.data_ptr(), src.data_ptr(), input_size * 2, s))
    ns.keepalive = src
    ns.input_copy_event = _input_copy_event(rt)
    return x, T, False


def sample_end_shard(rt, msg: ActivationMessage, ns, x: torch.Tensor) -> TokenResult:
    """normalize + lm_project + Sampler.sample (reference fit_in_memory.py:134-157).
    Greedy without top-logprobs is one fused kernel writing (token, logprob) to pinned
    host memory; everything else samples from the bf16 logits the kernel leaves in HBM."""
    model = rt.model
    if msg.temperature == 0 and msg.req_top_logprobs <= 0:
        model.head_sample_greedy(x, ns.kv, ns.result_token_ptr, ns.result_logprob_ptr, rt.compute_stream_ptr)
        rt.compute_stream.synchronize()
        tok = int(ns.result_i32[0].item())
        lp = float(ns.result_f32[1].item()) if msg.req_logprobs else 0.0
        return TokenResult(token_id=tok, logprob=lp, top_logprobs={})
    _, b16 = model.head_logits(x, want_f32=False, want_bf16=True, stream=rt.compute_stream_ptr)
    with torch.cuda.stream(rt.compute_stream):
        cfg = DecodingConfig(temperature=msg.temperature, top_p=msg.top_p, top_k=msg.top_k,
                             repetition_penalty=msg.repetition_penalty, min_p=msg.min_p,
                             min_tokens_to_keep=msg.min_tokens_to_keep)
        res = Sampler.sample(b16, cfg, req_logprobs=msg.req_logprobs, req_top_logprobs=msg.req_top_logprobs)
    ns.kv.set_token(res.token_id, rt.compute_stream_ptr)
    return res


def build_output(rt, msg: ActivationMessage, x: torch.Tensor, last_layer: int, final: Optional[TokenResult]) -> ActivationMessage:
    shape = (1, int(x.shape[0]), int(x.shape[1]))
    common = dict(nonce=msg.nonce, layer_id=last_layer, pool_id=-1, shape=shape, batch_size=msg.batch_size,
                  timestamp=utc_epoch_now(), node_origin=f"shard_{rt.shard_id}", dtype=rt._wire_dtype_str,
                  callback_url=msg.callback_url, lane=msg.lane, seq0=msg.seq0)
    if final is not None:
        return ActivationMessage(**common, is_final=True, token_id=final.token_id, logprob=final.logprob,
                                 top_logprobs=final.top_logprobs)
    ev = torch.cuda.Event()
    ev.record(rt.compute_stream)
    return ActivationMessage(**common, tensor=x.view(shape), ready_event=ev, req_logprobs=msg.req_logprobs,
                             req_top_logprobs=msg.req_top_logprobs, temperature=msg.temperature, top_p=msg.top_p,
                             top_k=msg.top_k, repetition_penalty=msg.repetition_penalty, min_p=msg.min_p,
                             min_tokens_to_keep=msg.min_tokens_to_keep)


def wait_ready(rt, weights: dict) -> None:
    """Order the compute stream after the layer's pinned->HBM copy (no host wait)."""
    ev = weights.get("_ready_event") if isinstance(weights, dict) else None
    if ev is not None:
        _cabi.check(_cabi.load().dn_stream_wait_event(rt.compute_stream_ptr, ev))


def wait_layers_ready(rt, weight_cache, layers) -> None:
    """Every resident layer record carries the event of its latest pinned->HBM copy."""
    for lid in layers:
        ent = weight_cache.cache.get(lid)
        if ent is not None:
            wait_ready(rt, ent[0])


def release_event(rt):
    """Event recorded on the compute stream after a layer's last kernel was enqueued."""
    import ctypes as C

    lib = _cabi.load()
    ev = C.c_void_p()
    _cabi.check(lib.dn_event_create(C.byref(ev), 0))
    _cabi.check(lib.dn_event_record(ev.value, rt.compute_stream_ptr))
    return ev.value
