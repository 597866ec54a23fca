"""bench_ring.py -- the N-shard leg of bench.py: one process per GPU (torchrun), each rank one
dnet shard holding a contiguous slice of Llama-3-8B layers; N sequences (nonces) in flight so
every shard computes while its neighbours' hops are in flight (micro-batch ring overlap).

Data path per token, no collective and no host in the loop:
  rank 0   : wait token flag -> [graph: embed(token slot) + layers] -> hop (8 KiB) -> rank 1
  rank r   : wait activation flag -> [graph: layers, in place in the slot]  -> hop -> rank r+1
  rank N-1 : wait -> [graph: layers + final norm + lm_head + argmax] -> hop (4 B token) -> rank 0
Hops are dn_hop_send / dn_hop_wait (peer cudaMemcpyAsync over NVLink + release/acquire flag).
torch.distributed (NCCL) is plumbing only: IPC-handle exchange, the untimed prefill relay,
barriers and the max-over-ranks reduction of the device time.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import sys
import time

import torch
import torch.distributed as dist


def run_ring(args, rank: int, local_rank: int, world: int) -> None:
    from bench import LLAMA3_8B, METRIC, ClockSampler, log, peaks, token_bytes, layer_bytes
    import bench as B
    from dnet_b200 import _cabi
    from dnet_b200.core.types.messages import ActivationMessage
    from dnet_b200.shard.models import ShardLoadModelRequest
    from dnet_b200.shard.ring import HopReceiver, HopSender, balanced_split, device_view, even_split
    from dnet_b200.shard.runtime import ShardRuntime
    from dnet_b200.utils.model import SyntheticSource
    from tests.helpers import token_message

    PROMPT_LEN = B.PROMPT_LEN
    os.environ["NCCL_DEBUG"] = os.environ.get("DNET_NCCL_DEBUG", "WARN")   # stdout must stay ONE JSON line
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    gloo = dist.new_group(backend="gloo")
    lib = _cabi.load()
    cfg = dict(LLAMA3_8B)
    if args.layers:
        cfg["num_hidden_layers"] = args.layers
    L, H = cfg["num_hidden_layers"], cfg["hidden_size"]
    K, W = args.steps, args.warmup
    NS = world if args.in_flight <= 0 else args.in_flight
    if args.split == "equal":
        split = even_split(L, world)
    else:
        # contiguous slices balanced by the bytes a shard streams per token (the last shard also owns the
        # lm_head = 2.4 layers' worth): the assignment an operator posts to /v1/prepare_topology_manual
        split = balanced_split(L, world, layer_bytes(cfg), first_extra=2 * cfg["hidden_size"],
                               last_extra=2 * cfg["vocab_size"] * cfg["hidden_size"] + 2 * cfg["hidden_size"])
    mine = split[rank]
    first, last = rank == 0, rank == world - 1
    need = PROMPT_LEN + 2 * (W + K) + 96
    rt = ShardRuntime(shard_id=rank)
    rt.kv_cache_config.max_tokens = need
    os.environ["DNET_KV_POOL_PAGES"] = str(((need + 63) // 64) * (NS + 1))
    from dnet_b200.config import get_settings
    get_settings.cache_clear()
    rt.load_model_core(ShardLoadModelRequest(model_path=SyntheticSource(cfg, seed=0), total_layers=L, layers=mine,
                                             window_size=len(mine), residency_size=len(mine), kv_bits="fp16"))
    lib.dn_set_option(b"pdl", 1 if args.pdl else 0)
    lib.dn_set_option(b"l2_prefetch_kb", args.l2_prefetch_kb)
    if args.pf_depth >= 0:
        lib.dn_set_option(b"pf_depth", args.pf_depth)
    pol, model = rt.policy, rt.model
    s = rt.compute_stream_ptr
    stream = rt.compute_stream

    # ---- hop endpoints: receiver-owned slots, exported once (configure_topology analogue)
    slot_bytes = H * 2
    rx = HopReceiver(NS, slot_bytes)
    eps = [None] * world
    dist.all_gather_object(eps, rx.endpoint())
    nxt = (rank + 1) % world
    transport = "cuda-ipc peer memcpy + flag"
    ok = torch.ones(1, device="cuda")
    tx = None
    try:
        tx = HopSender(eps[nxt]) if world > 1 else HopSender(eps[rank], (rx.data_ptr, rx.flag_ptr))
    except Exception as e:       # CUDA IPC unavailable in this container: fall back to NCCL p2p for the hop
        log(f"rank {rank}: CUDA IPC import failed ({e}); falling back to NCCL p2p hops")
        ok.zero_()
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    use_ipc = bool(int(ok.item()))
    if not use_ipc:
        transport = "nccl p2p (CUDA IPC unavailable)"

    # ---- untimed prefill of every nonce through the public policy API, relayed with NCCL p2p
    g = torch.Generator().manual_seed(1234)
    prompts = [torch.randint(0, cfg["vocab_size"], (PROMPT_LEN,), generator=g).tolist() for _ in range(NS)]
    first_tok = torch.zeros(NS, dtype=torch.int32, device="cuda")
    for n in range(NS):
        nonce = f"n{n}"
        if first:
            pol.process(token_message(rt, nonce, prompts[n], req_logprobs=True))
        else:
            buf = torch.empty(PROMPT_LEN, H, dtype=torch.bfloat16, device="cuda")
            dist.recv(buf, src=rank - 1)
            msg = ActivationMessage(nonce=nonce, pool_id=-1, batch_size=1, shape=(1, PROMPT_LEN, H), dtype="bfloat16",
                                    layer_id=mine[0] - 1, timestamp=0, node_origin=f"shard_{rank - 1}",
                                    callback_url="", tensor=buf, temperature=0.0, req_logprobs=True)
            pol.process(msg)
        res = rt.activation_send_queue.get_nowait()
        if not last:
            stream.synchronize()
            dist.send(res.tensor.reshape(PROMPT_LEN, H).contiguous(), dst=rank + 1)
        else:
            first_tok[n] = res.token_id
    dist.broadcast(first_tok, src=world - 1)
    torch.cuda.synchronize()
    log(f"rank {rank}: layers {mine[0]}..{mine[-1]} prefilled {NS} nonces; first tokens {first_tok.tolist()}")

    # ---- per-nonce step graphs over the hop slots
    states = [rt.get_or_make_kv(f"n{n}") for n in range(NS)]
    tok_local = torch.zeros(NS * 16, dtype=torch.int32, device="cuda")        # last rank: sampled token per nonce
    lp_local = torch.zeros(NS * 16, dtype=torch.float32, device="cuda")
    arr = (C.c_int32 * len(mine))(*mine)
    graphs = []
    use_mk = bool(args.megakernel)
    for n in range(NS):
        ns = states[n]
        xptr = ns.x1.data_ptr() if first else rx.slot(n)   # ranks > 0 compute in place in the hop slot
        if use_mk:
            graphs.append((None, xptr))
            continue
        _cabi.check(lib.dn_graph_begin(s))
        gp = C.c_void_p()
        try:
            if first:
                # the token slot of rank 0 is the first 4 bytes of its slot; embed into a private buffer
                _cabi.check(lib.dn_embed(model._h, rx.slot(n), 1, xptr, s))
            _cabi.check(lib.dn_window_forward(model._h, arr, len(mine), xptr, 1, ns.kv._h, s))
            if last:
                _cabi.check(lib.dn_head_sample_greedy(model._h, xptr, 1, ns.kv._h, tok_local.data_ptr() + n * 64,
                                                      lp_local.data_ptr() + n * 64, s))
            _cabi.check(lib.dn_kv_advance(ns.kv._h, 1, s))
        finally:
            rc = lib.dn_graph_end(s, C.byref(gp))
        _cabi.check(rc)
        graphs.append((gp.value, xptr))

    fused = use_mk and bool(args.fused_hop)

    def run_step_fused(n: int, st: int) -> None:
        """wait + step + hop in ONE kernel launch (dn_shard_step_hop)"""
        ns = states[n]
        if last:
            dst, dflag, dseq = tx.data_ptr + n * tx.ep.slot_bytes, tx.flag_ptr + n * 64, st + 2
        else:
            dst, dflag, dseq = tx.data_ptr + n * tx.ep.slot_bytes, tx.flag_ptr + n * 64, st + 1
        _cabi.check(lib.dn_shard_step_hop(model._h, arr, len(mine), graphs[n][1], ns.kv._h, 1 if first else 0,
                                          1 if last else 0, tok_local.data_ptr() + n * 64 if last else None,
                                          lp_local.data_ptr() + n * 64 if last else None, 1,
                                          rx.flag(n), st + 1, rx.slot(n) if first else None, dst, dflag, dseq, s))

    def run_step(n: int) -> None:
        """one shard step of nonce n on the compute stream"""
        ns = states[n]
        if use_mk:
            if first:   # token arrives in the hop slot: move it into the nonce's step state, then embed from it
                _cabi.check(lib.dn_memcpy_d2h(ns.kv.token_ptr, rx.slot(n), 4, s))   # 4-byte D2D (UVA)
            _cabi.check(lib.dn_shard_step(model._h, arr, len(mine), graphs[n][1], ns.kv._h, 1 if first else 0,
                                          1 if last else 0, tok_local.data_ptr() + n * 64 if last else None,
                                          lp_local.data_ptr() + n * 64 if last else None, None, 1, s))
        else:
            _cabi.check(lib.dn_graph_launch(graphs[n][0], s))
            ns.kv.note_advance(1)

    if first:   # seed step 0: tokens into our own token slots, flags -> seq 1
        for n in range(NS):
            device_view(rx.slot(n), (1,), torch.int32).copy_(first_tok[n:n + 1])
        torch.cuda.synchronize()
        for n in range(NS):
            rx.set_local(n, 1, s)
    stream.synchronize()
    dist.barrier()

    step_no = [0]
    tok_seeded = [True] * NS          # NCCL fallback: rank 0 owns the step-0 tokens already
    prev_rank = (rank - 1) % world

    def enqueue(nsteps: int, nonces):
        for _ in range(nsteps):
            st = step_no[0]
            for n in nonces:
                if fused and use_ipc:
                    run_step_fused(n, st)
                    continue
                if use_ipc:
                    rx.wait(n, st + 1, s)
                else:
                    with torch.cuda.stream(stream):
                        if first:
                            if tok_seeded[n]:
                                tok_seeded[n] = False
                            else:
                                dist.recv(device_view(rx.slot(n), (1,), torch.int32), src=prev_rank)
                        else:
                            dist.recv(device_view(rx.slot(n), (H,), torch.bfloat16), src=prev_rank)
                run_step(n)
                if use_ipc:
                    if last:
                        tx.send(n, tok_local.data_ptr() + n * 64, 16, st + 2, s)     # token (16-byte line) for the NEXT step
                    else:
                        tx.send(n, graphs[n][1], slot_bytes, st + 1, s)
                else:
                    with torch.cuda.stream(stream):
                        if last:
                            dist.send(tok_local[n * 16:n * 16 + 1], dst=nxt)
                        else:
                            dist.send(device_view(graphs[n][1], (H,), torch.bfloat16), dst=nxt)
            step_no[0] += 1

    def drain():
        """NCCL fallback: rank 0 posts the receives matching the last shard's final token sends."""
        if use_ipc or not first:
            return
        with torch.cuda.stream(stream):
            for n in range(NS):
                if not tok_seeded[n]:
                    dist.recv(device_view(rx.slot(n), (1,), torch.int32), src=prev_rank)
                    tok_seeded[n] = True

    allns = list(range(NS))
    enqueue(W, allns)
    stream.synchronize()
    dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    l0 = lib.dn_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    dist.barrier()
    tw0 = time.perf_counter()
    e0.record(stream)
    enqueue(K, allns)
    e1.record(stream)
    drain()
    stream.synchronize()
    torch.cuda.synchronize()
    dist.barrier()
    tw1 = time.perf_counter()
    ms_local = e0.elapsed_time(e1)
    # nonce 0's token after W+K decode steps: must equal the 1-shard run's (bench.py N=1 "check")
    chk = torch.zeros(1, dtype=torch.int32, device="cuda")
    if last:
        chk.copy_(tok_local[0:1])
    dist.broadcast(chk, src=world - 1)
    check_token = int(chk.item())
    launches = int(lib.dn_launch_count() - l0)
    t = torch.tensor([ms_local], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
    dist.all_reduce(lt, op=dist.ReduceOp.SUM)
    timed_out = torch.tensor([1 if rx.timed_out() else 0], device="cuda")
    dist.all_reduce(timed_out, op=dist.ReduceOp.MAX)
    value = K * NS / ms * 1e3
    clocks = sampler.summary(tw0, tw1) if rank == 0 else None

    # ---- single sequence around the ring: latency view (tok/s of ONE sequence, hop cost)
    dist.barrier()
    K1 = min(K, 64)
    for n in range(1, NS):      # nonces 1.. sit out: nothing to do, their flags simply do not advance
        pass
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    dist.barrier()
    e2.record(stream)
    enqueue(K1, [0])
    e3.record(stream)
    drain()
    stream.synchronize()
    dist.barrier()
    t1 = torch.tensor([e2.elapsed_time(e3)], device="cuda")
    dist.all_reduce(t1, op=dist.ReduceOp.MAX)
    single_ms = float(t1.item()) / K1
    # this rank's pure compute per step (graph alone, flags already satisfied -> no waiting)
    torch.cuda.synchronize()
    e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    # replay nonce NS-1's graph on its own (KV grows, results discarded; untimed state only)
    e4.record(stream)
    reps = 16
    for _ in range(reps):
        run_step(NS - 1)
    e5.record(stream)
    stream.synchronize()
    comp = torch.tensor([e4.elapsed_time(e5) / reps], device="cuda")
    comp_all = [torch.zeros_like(comp) for _ in range(world)]
    dist.all_gather(comp_all, comp)
    comp_ms = [float(c.item()) for c in comp_all]
    hop_us = (single_ms - sum(comp_ms)) / world * 1e3

    # ---- e2e: host-driven ring through policy.process; token id crosses the HOST between the
    #      last shard and shard 0 (the reference closes the ring through the API node); one sequence
    e2e = None
    if not args.no_e2e:
        nonce = "e2e"
        ns = None
        E = min(K, 128)
        tokbuf = torch.zeros(1, dtype=torch.int32)
        # prefill
        if first:
            pol.process(token_message(rt, nonce, prompts[0], req_logprobs=True))
        else:
            buf = torch.empty(PROMPT_LEN, H, dtype=torch.bfloat16, device="cuda")
            dist.recv(buf, src=rank - 1)
            pol.process(ActivationMessage(nonce=nonce, pool_id=-1, batch_size=1, shape=(1, PROMPT_LEN, H),
                                          dtype="bfloat16", layer_id=mine[0] - 1, timestamp=0, node_origin="",
                                          callback_url="", tensor=buf, temperature=0.0, req_logprobs=True))
        res = rt.activation_send_queue.get_nowait()
        if not last:
            stream.synchronize()
            dist.send(res.tensor.reshape(PROMPT_LEN, H).contiguous(), dst=rank + 1)
        else:
            tokbuf[0] = res.token_id
        dist.broadcast(tokbuf, src=world - 1, group=gloo)
        xin = torch.empty(1, H, dtype=torch.bfloat16, device="cuda")
        torch.cuda.synchronize()
        dist.barrier()
        for i in range(W + E):
            if i == W:
                torch.cuda.synchronize()
                dist.barrier()
                te0 = time.perf_counter()
            if first:
                pol.process(token_message(rt, nonce, [int(tokbuf[0])], req_logprobs=True))   # host id -> H2D
            else:
                with torch.cuda.stream(stream):
                    dist.recv(xin, src=rank - 1)  # NVLink p2p (NCCL) -- host-driven variant of the hop
                pol.process(ActivationMessage(nonce=nonce, pool_id=-1, batch_size=1, shape=(1, 1, H), dtype="bfloat16",
                                              layer_id=mine[0] - 1, timestamp=0, node_origin="", callback_url="",
                                              tensor=xin, temperature=0.0, req_logprobs=True))
            res = rt.activation_send_queue.get_nowait()
            if not last:
                with torch.cuda.stream(stream):
                    dist.send(res.tensor.reshape(1, H), dst=rank + 1)
            else:
                tokbuf[0] = res.token_id          # D2H happened inside process (pinned result)
            dist.broadcast(tokbuf, src=world - 1, group=gloo)   # last shard -> API -> shard 0, on the host
        torch.cuda.synchronize()
        dist.barrier()
        te1 = time.perf_counter()
        tt = torch.tensor([te1 - te0], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": E / float(tt.item()), "unit": "tok/s", "h2d_bytes_per_step": 4, "d2h_bytes_per_step": 8,
               "sequences_in_flight": 1, "steps": E,
               "api": "policy.process per shard; activations rank->rank over NVLink (NCCL p2p), token last shard -> "
                      "host -> shard 0 (gloo), like the reference's SendToken / API loop"}

    if rank == 0:
        peak, peak_src = peaks()
        tb = token_bytes(cfg)
        per_shard = [len(x) * layer_bytes(cfg) for x in split]
        per_shard[-1] += 2 * cfg["vocab_size"] * H + 2 * H
        per_shard[0] += 2 * H
        slow = max(per_shard)
        roofline = {"bound": "hbm", "kernel": "whole shard step (see N=1 line for the per-kernel roofline)",
                    "achieved": tb * value / 1e9 / world, "peak": peak, "unit": "GB/s per GPU",
                    "frac": (value / (peak * 1e9 / slow)), "traffic": None, "peak_source": peak_src,
                    "note": "frac = aggregate tok/s / (1 / time for the busiest shard to stream its bytes at peak)",
                    "busiest_shard_bytes": slow, "pipelined_roofline_tok_s": peak * 1e9 / slow,
                    "single_sequence": {"tok_s": 1e3 / single_ms, "ms_per_token": single_ms,
                                        "roofline_tok_s": peak * 1e9 / tb, "frac": (1e3 / single_ms) / (peak * 1e9 / tb),
                                        "per_rank_compute_ms": comp_ms, "ring_hop_us": hop_us,
                                        "ring_hop_note": "(one sequence's round trip - sum of the shards' stand-alone step times) / shards; "
                                                         "negative = the hop is hidden: a waiting shard's step kernel already streams its first "
                                                         "weight stages while it spins on the predecessor's flag"}}
        out = {
            "metric": METRIC, "value": value, "unit": "tok/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": f"Llama-3-8B bf16 bs=1 decode, {world} shards x {'/'.join(str(len(x)) for x in split)} layers ({args.split} contiguous split) pipelined ring "
                                   f"(BASELINE configs[1]), {NS} sequences in flight (one per shard), each bs=1",
                       "prompt_len": PROMPT_LEN, "kv": "fp16 paged (64-token pages)", "wire_dtype": "bf16",
                       "l2": "inputs larger than L2 (>=1.7 GB of weights per shard step vs 126 MB L2); no flush",
                       "pdl": bool(args.pdl), "step_kernel": "k_shard_step" if use_mk else "per-op kernels in a CUDA graph", "sequences_in_flight": NS, "hop": transport + (" fused into k_shard_step (dn_shard_step_hop)" if (fused and use_ipc) else " (dn_hop_wait / dn_hop_send kernels)"),
                       "split": [f"{x[0]}-{x[-1]}" for x in split]},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(lt.item()), "roofline": roofline, "cpu_baseline": None,
            "hop_timeout": bool(int(timed_out.item())),
            "check": {"nonce0_token_after_steps": W + K, "token": check_token,
                      "note": "equals the N=1 line's check.token for the same --steps/--warmup (split is bit-exact)"},
        }
        B.emit(out)
    sampler.stop()
    dist.barrier()
    rt.unload_model_core()
    dist.destroy_process_group()
