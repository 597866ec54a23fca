"""bench_ring.py -- the measured leg of bench.py at any N: one process per GPU, each rank one dnet
shard (``ShardNode`` = ShardRuntime + RingAdapter + Shard + gRPC server) holding a contiguous slice
of Llama-3-8B layers, driven ONLY through the product's public transport API:

  ShardNode.load_model(req)            load + configure_topology: hop lanes exported by CUDA IPC, the
                                       successor's lanes mapped (b200.hop.open over the control channel)
  ApiNode / InferenceManager           the API-side token loop: send_tokens(prompt) -> first token,
                                       lease(nonce, steps) -> the ring decodes with the token loop closed
                                       on the device, tokens observed through SendToken / TokenTap

There is no bench-private scheduling: the head shard's RingAdapter merges the leases of the N
nonces in flight into the schedule every shard launches (dn_shard_step_hop, one fused
wait+step+hop kernel per (nonce, step)).  torch.distributed is plumbing only (barriers, the
max-over-ranks reduction of the device time); no collective is on the data path.

Timed region (both numbers come from the SAME K steps):
  value  CUDA events on every rank's compute stream from right before its first kernel of the K steps x NS
         nonces to behind its last, max over ranks  (device time: pipeline fill and every gap the host-side
         scheduling leaves between launches included; the lease / schedule-frame latency before the first launch
         is not -- that is in e2e)
  e2e    wall clock on the API rank from submitting the leases to holding the last token on the host
         (per step and nonce: 8 bytes device->host through the pinned token ring; host->device: the
         schedule frame entries, 8 bytes per (nonce, step))
"""
from __future__ import annotations

import json
import os
import sys
import time
import types


def _wait(cond, timeout_s: float, what: str, poll: float = 2e-4):
    t0 = time.perf_counter()
    while not cond():
        if time.perf_counter() - t0 > timeout_s:
            raise TimeoutError(f"timed out waiting for {what}")
        time.sleep(poll)


def run_ring(args, rank: int, local_rank: int, world: int) -> None:
    import torch

    import bench as B
    from bench import LLAMA3_8B, METRIC, ClockSampler, log, peaks, token_bytes, layer_bytes
    from dnet_b200 import _cabi
    from dnet_b200.config import TransportSettings
    from dnet_b200.shard.models import ShardLoadModelRequest
    from dnet_b200.shard.node import ApiNode, ShardNode
    from dnet_b200.shard.ring import balanced_split, even_split
    from dnet_b200.utils.model import SyntheticSource

    dist = None
    gloo = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG", "INFO")                  # NCCL's own log goes to stderr (stdout is guarded)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # Barriers and the few scalar reductions of the harness run over gloo, on the HOST: an NCCL barrier is a
        # kernel that spins on the GPU (on the legacy default stream) until every rank arrives, and here ranks
        # arrive at very different times while their shards must keep serving the ring.
        gloo = dist.new_group(backend="gloo")
        ok = torch.ones(1, device="cuda")
        dist.all_reduce(ok)                                            # one NCCL collective: rank count check
        assert int(ok.item()) == world

    def barrier():
        if dist is not None:
            dist.barrier(group=gloo)

    PROMPT_LEN = B.PROMPT_LEN
    lib = _cabi.load()
    cfg = dict(LLAMA3_8B)
    if args.layers:
        cfg["num_hidden_layers"] = args.layers
    L, H = cfg["num_hidden_layers"], cfg["hidden_size"]
    K, W = args.steps, args.warmup
    ts_lag = max(0, int(os.environ.get("DNET_TRANSPORT_HEAD_TP_LAG", "1")))
    tp_wanted = world >= 2 and (args.head_tp == "on" or (args.head_tp == "auto" and world >= 4))
    # sequences in flight: one per shard keeps a plain ring busy; with the lm_head tensor-parallel over the ring a
    # token is finalised one slot after its last layer, so S + 1 sequences keep all S shards busy
    NS = (world + (1 + ts_lag if tp_wanted else 0)) if args.in_flight <= 0 else args.in_flight
    if args.split == "equal" or world == 1 or (tp_wanted and L % world == 0):
        split = even_split(L, world)            # tensor-parallel head: every shard streams the same bytes with equal counts
    else:
        # contiguous slices balanced by the bytes a shard streams per token (the last shard also owns the
        # lm_head = 2.4 layers' worth): the assignment an operator posts to /v1/prepare_topology_manual
        split = balanced_split(L, world, layer_bytes(cfg), first_extra=2 * H,
                               last_extra=2 * cfg["vocab_size"] * H + 2 * H)
    mine = split[rank]
    first, last = rank == 0, rank == world - 1

    # ---- the shard process: runtime + adapter + gRPC server (what dnet-shard assembles)
    base_port = (int(os.environ.get("MASTER_PORT", "29500")) % 20000) + 30000
    ports = [base_port + 7 * r for r in range(world)]
    ts = TransportSettings()
    ts.hop_lanes = max(NS + 2, 4)
    ts.head_tp = "on" if tp_wanted else "off"
    ts.sched_rounds_per_frame = args.sched_rounds
    ts.sched_frames_in_flight = args.sched_depth
    node = ShardNode(rank, ports[rank], transport_settings=ts, queue_size=256).start()
    rt = node.runtime
    need = PROMPT_LEN + 2 * (W + K) + 160
    rt.kv_cache_config.max_tokens = need
    os.environ["DNET_KV_POOL_PAGES"] = str(((need + 63) // 64) * (NS + 3))
    from dnet_b200.config import get_settings
    get_settings.cache_clear()
    nxt = None if world == 1 else types.SimpleNamespace(local_ip="127.0.0.1", shard_port=ports[(rank + 1) % world],
                                                        instance=f"shard{(rank + 1) % world}")
    t_load = time.perf_counter()
    res = node.load_model(ShardLoadModelRequest(model_path=SyntheticSource(cfg, seed=0), total_layers=L, layers=mine,
                                                window_size=len(mine), residency_size=len(mine), kv_bits="fp16",
                                                next_node=nxt))
    assert res.success, res.message
    ad = node.adapter
    if ad.hop is None:
        raise RuntimeError("device hop link did not come up (CUDA IPC unavailable?): this bench measures the hop transport")
    for k, v in (("pdl", 1 if args.pdl else 0), ("l2_prefetch_kb", args.l2_prefetch_kb), ("mk_flags", args.mk_flags)):
        lib.dn_set_option(k.encode(), v)
    for k, v in (("pf_depth", args.pf_depth), ("inflight", args.inflight), ("inflight_hi", args.inflight_hi), ("park", args.park)):
        if v >= 0:
            lib.dn_set_option(k.encode(), v)
    if args.attn_chunk:
        lib.dn_set_option(b"attn_chunk", args.attn_chunk)
    if args.attn_tc >= 0:
        lib.dn_set_option(b"attn_tc", args.attn_tc)
    barrier()
    log(f"rank {rank}: layers {mine[0]}..{mine[-1]} loaded + topology configured in {time.perf_counter() - t_load:.1f}s "
        f"(hop lanes {ad.n_lanes}, head={ad.is_head} tail={ad.is_tail})")

    # ---- the API node sits with the finalising shard's process: tokens reach it through the TokenTap (pinned
    #      host ring) and an in-process sink; prompts and leases go to the head shard over gRPC
    api = None
    g = torch.Generator().manual_seed(1234)
    prompts = [torch.randint(0, cfg["vocab_size"], (PROMPT_LEN,), generator=g).tolist() for _ in range(NS)]
    nonces = [f"n{n}" for n in range(NS)]
    got = {n: [] for n in nonces}            # tokens as the API receives them
    stamps = []                              # host arrival time of every token (steady-state rate, start-up share)
    stream = rt.compute_stream
    pol = rt.policy

    def entries_done() -> int:
        return int(getattr(pol, "sched_entries_done", 0))

    tp = bool(ad.head_tp)
    api_rank = 0 if tp else world - 1          # tokens surface on the head shard with a tensor-parallel head, else on the tail
    api_port = base_port + 7 * world + 3
    cb = f"grpc://127.0.0.1:{api_port}" if tp else "local://"
    on_api = rank == api_rank
    if on_api:
        api = ApiNode(f"127.0.0.1:{ports[0]}", callback="local://" if not tp else "grpc", grpc_port=api_port)

        def sink(msg):
            got.setdefault(msg.nonce, []).append((int(msg.token_id), float(msg.logprob)))
            stamps.append(time.perf_counter())
        ad.token_sink = sink                    # decode tokens: in-process, straight from the TokenTap
        if tp:                                  # the tail's first tokens (prefill) arrive over gRPC SendToken
            api.manager.resolve_request = lambda nonce, res: got.setdefault(nonce, []).append((int(res.token_id), float(res.logprob)))

        async def send_prompts():
            import numpy as np
            for n, nonce in enumerate(nonces):
                await api.adapter.send_tokens(nonce, np.asarray(prompts[n], np.int32).tobytes(), cb, logprobs=True,
                                              decoding_config=types.SimpleNamespace(temperature=0.0, top_p=1.0, top_k=-1,
                                                                                    repetition_penalty=1.0, min_p=0.0,
                                                                                    min_tokens_to_keep=1))
        api.call(send_prompts())
        _wait(lambda: all(len(got[n]) >= 1 for n in nonces), 120, "first tokens (prefill)")
        log(f"prefilled {NS} nonces through the ring; first tokens {[got[n][0][0] for n in nonces]}")
    barrier()

    def lease_all(steps: int):
        async def go():
            for nonce in nonces:
                await api.adapter.lease(nonce, steps, cb)
        api.call(go())

    def run_steps(steps: int, base_tokens: int, base_entries: int, timed: bool):
        """`steps` decode steps of every nonce; returns (device ms on this rank, wall seconds on the API rank)"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stream.synchronize()
        torch.cuda.synchronize()
        barrier()
        # e0 is recorded by the policy on the compute stream right before this rank launches its first kernel of the
        # window (device time of the K steps; the control plane's latency -- lease, schedule, frame -- is in `e2e`)
        pol.sched_marks[base_entries] = e0
        t0 = time.perf_counter()
        if on_api:
            lease_all(steps)
        _wait(lambda: entries_done() >= base_entries + steps * NS, 180, "schedule frames", poll=2e-4)
        wall = None
        if on_api and tp:      # head shard: its last kernels are the merges of the final tokens -> time through them
            _wait(lambda: all(len(got[n]) >= base_tokens + steps for n in nonces), 180, "tokens", poll=1e-4)
            wall = time.perf_counter() - t0
        e1.record(stream)
        if on_api and not tp:
            _wait(lambda: all(len(got[n]) >= base_tokens + steps for n in nonces), 180, "tokens", poll=1e-4)
            wall = time.perf_counter() - t0
        stream.synchronize()
        torch.cuda.synchronize()
        barrier()
        run_steps.t_lease = t0
        return e0.elapsed_time(e1), wall

    # ---- warm-up, then the timed K steps
    run_steps(W, 1, 0, False)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    import gc
    gc.collect()
    gc.freeze()                 # a serving process does the same: no stop-the-world collection of the (large, static) heap
    gc.disable()                # inside a request's latency path
    l0 = lib.dn_launch_count()
    h0, n0 = pol.sched_host_s, pol.sched_host_entries
    stamp0 = len(stamps)
    tw0 = time.perf_counter()
    ms_local, wall = run_steps(K, 1 + W, W * NS, True)
    host_us = (pol.sched_host_s - h0) / max(1, pol.sched_host_entries - n0) * 1e6
    tw1 = time.perf_counter()
    gc.enable()
    launches = int(lib.dn_launch_count() - l0)
    step_err = int(lib.dn_step_error(rt.model._h, rt.compute_stream_ptr))

    def allmax(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=gloo)
        return float(t.item())

    def allsum(x: int) -> int:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=gloo)
        return int(t.item())

    ms = allmax(ms_local)
    host_us_max = allmax(host_us)
    steady = None
    if on_api:
        ts_ = stamps[stamp0:stamp0 + K * NS]
        if len(ts_) >= 20:
            a, b = len(ts_) // 10, len(ts_) - len(ts_) // 10 - 1
            steady = {"tok_s_middle_80pct": (b - a) / (ts_[b] - ts_[a]), "first_token_ms_after_lease": (ts_[0] - run_steps.t_lease) * 1e3,
                      "note": "host arrival times of the timed tokens: rate between the 10 % and 90 % marks, and how long after "
                              "submitting the leases the first token arrived (lease -> schedule frames round the ring -> pipeline fill)"}
    steady_v = allmax(steady["tok_s_middle_80pct"] if steady else 0.0)
    first_ms = allmax(steady["first_token_ms_after_lease"] if steady else 0.0)
    e2e_s = allmax(wall if wall is not None else 0.0)
    launches_all = allsum(launches)
    step_err = int(allmax(float(step_err)))
    value = K * NS / ms * 1e3
    check_token = int(allmax(float(got[nonces[0]][W + K][0]) if on_api else -1.0))
    clocks = sampler.summary(tw0, tw1) if rank == 0 else None
    tokens_ok = bool(allmax(0.0 if (not on_api or all(t >= 0 for n in nonces for t, _ in got[n])) else 1.0) == 0.0)

    # ---- one sequence alone around the ring (latency view) + this rank's stand-alone step time
    K1 = min(K, 64)
    single_ms = None
    if not args.no_single:
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stream.synchronize()
        barrier()
        base_e = entries_done()
        base_t = len(got[nonces[0]])
        e2.record(stream)
        if on_api:
            api.call(api.adapter.lease(nonces[0], K1, cb))
        _wait(lambda: entries_done() >= base_e + K1, 120, "single-sequence schedule", poll=2e-4)
        if on_api:
            _wait(lambda: len(got[nonces[0]]) >= base_t + K1, 120, "single-sequence tokens")
        e3.record(stream)
        stream.synchronize()
        barrier()
        single_ms = allmax(e2.elapsed_time(e3)) / K1
    hop = B.measure_hop(rt, ad, barrier, rank, world) if world > 1 else None

    # ---- N=1 extras: per-kernel times, DRAM traffic of the dominant kernel, CPU baseline
    extras = B.single_gpu_extras(args, rt, cfg, K, ms) if world == 1 else {}
    cpu, parity = None, None
    if rank == 0 and not args.no_cpu:
        if world == 1:
            cpu, parity = B.cpu_baseline_leg(args, cfg, rt=rt, prompt=prompts[0], gpu_tokens=[t for t, _ in got[nonces[0]]])
        else:
            cpu, parity = B.cpu_baseline_leg(args, cfg)

    if rank == 0:
        peak, peak_src = peaks()
        tb = token_bytes(cfg)
        per_shard = [len(x) * layer_bytes(cfg) for x in split]
        if tp:
            per_shard = [b + (2 * cfg["vocab_size"] * H) // world + 2 * H for b in per_shard]
        else:
            per_shard[-1] += 2 * cfg["vocab_size"] * H + 2 * H
        per_shard[0] += 2 * H
        slow = max(per_shard)
        if world == 1:
            ach = tb / (ms / K) / 1e6
            roofline = {"bound": "hbm", "kernel": extras.get("kernel_name"), "achieved": ach, "peak": peak, "unit": "GB/s",
                        "frac": ach / peak, "traffic": extras.get("traffic"), "traffic_source": extras.get("traffic_source"),
                        "peak_source": peak_src, "algorithmic_bytes_per_launch": tb, "launch_ms": ms / K,
                        "note": "launch_ms = CUDA-event time of the K timed launches / K on the compute stream (includes host "
                                "scheduling gaps between launches); one k_shard_step launch per token",
                        "step": {"algorithmic_bytes_per_token": tb, "achieved_gbs": tb * value / 1e9,
                                 "frac": tb * value / 1e9 / peak, "roofline_tok_s": peak * 1e9 / tb},
                        "per_op_kernels_timed_alone": extras.get("kernels")}
        else:
            roofline = {"bound": "hbm", "kernel": "k_shard_step (whole shard step incl. fused hop; see the N=1 line for the per-kernel roofline)",
                        "achieved": tb * value / 1e9 / world, "peak": peak, "unit": "GB/s per GPU",
                        "frac": (value / (peak * 1e9 / slow)), "traffic": None, "peak_source": peak_src,
                        "note": "frac = aggregate tok/s / (1 / time for the busiest shard to stream its bytes at peak)",
                        "busiest_shard_bytes": slow, "pipelined_roofline_tok_s": peak * 1e9 / slow}
        if single_ms is not None:
            roofline["single_sequence"] = {"tok_s": 1e3 / single_ms, "ms_per_token": single_ms, "roofline_tok_s": peak * 1e9 / tb,
                                           "frac": (1e3 / single_ms) / (peak * 1e9 / tb)}
        e2e = {"value": K * NS / e2e_s, "unit": "tok/s", "h2d_bytes_per_step": 8 * NS, "d2h_bytes_per_step": 8 * NS,
               "sequences_in_flight": NS, "steps": K,
               "api": "ApiNode: RingApiAdapter.send_tokens(prompt) + lease(nonce, steps) over gRPC to the head shard's RingAdapter; "
                      "every token read on the host from the finalising shard's pinned TokenTap ring (token + logprob) and "
                      "delivered to the API's resolve_token; wall clock from lease submission to the last token on the host",
               "frames": dict(ad.stats)}
        out = {
            "metric": METRIC, "value": value, "unit": "tok/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": f"Llama-3-8B bf16 bs=1 decode, {world} shard{'s' if world > 1 else ''} x "
                                   f"{'/'.join(str(len(x)) for x in split)} layers ({'1 shard' if world == 1 else args.split + ' contiguous split'}) "
                                   f"pipelined ring (BASELINE configs[1]), {NS} sequence{'s' if NS > 1 else ''} in flight, each bs=1",
                       "prompt_len": PROMPT_LEN, "kv": "fp16 paged (64-token pages)", "wire_dtype": "bf16",
                       "l2": "inputs larger than L2 (>=1.7 GB of weights per shard step vs 126 MB L2); no flush",
                       "step_kernel": "k_shard_step via RingAdapter schedule (dn_shard_step_hop: wait + step + hop in one launch)",
                       "sequences_in_flight": NS, "hop": "CUDA-IPC peer stores + system-scope flag, fused into k_shard_step",
                       "lm_head": (f"tensor-parallel over the {world} shards (vocab/{world} rows each; final hidden state broadcast + "
                                   "partial (max, sum-exp, argmax) gather over NVLink inside k_shard_step)") if tp else "on the last shard",
                       "split": [f"{x[0]}-{x[-1]}" for x in split], "step_error": step_err,
                       "sched": {"rounds_per_frame": ts.sched_rounds_per_frame, "frames_in_flight": ts.sched_frames_in_flight,
                                 "host_us_per_entry_max_rank": host_us_max}},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches_all, "roofline": roofline, "cpu_baseline": cpu,
            "ring_hop_us": hop,
            "steady_state": {"tok_s_middle_80pct": steady_v, "first_token_ms_after_lease": first_ms},
            "check": {"nonce0_token_after_steps": W + K, "token": check_token, "all_tokens_valid": tokens_ok,
                      "oracle_full_depth": parity,
                      "note": "nonce 0's token after W+K decode steps: identical at every N and for both splits"},
        }
        B.emit(out)
    sampler.stop()
    barrier()
    if api is not None:
        async def end_all():
            for nonce in nonces:
                await api.adapter.end_request(nonce)
        try:
            api.call(end_all(), 30)
        except Exception:
            pass
        api.shutdown()
    barrier()
    node.unload_model()
    node.shutdown()
    if dist is not None:
        dist.destroy_process_group()
