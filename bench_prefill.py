"""bench_prefill.py -- the long-context path: a long prompt prefilled through the ring as a
stream of chunks (`bench.py --config prefill`, one process per GPU under torchrun).

The API (`InferenceManager.generate_stream(prefill_chunk=C)`) cuts the prompt into C-token ``tokens`` frames, all but
the last flagged "more follows"; every shard runs its layers over a chunk (tcgen05 GEMMs + causal flash attention over
the paged KV written so far), hands the chunk's activation to its successor's bulk slot over NVLink (credit flags, a
metadata-only frame per chunk) and starts on the next chunk, so S shards work on S chunks at once; the last shard
samples after the final chunk only.  Reported: prompt tokens / second from the first frame leaving the API to the
first generated token arriving back (wall clock at the API -- the interval spans all GPUs, no single device clock
covers it), K prompts after W warm-up prompts, each a fresh request (fresh KV).  Llama-3-8B dims, bf16, 16-bit KV."""
from __future__ import annotations

import os
import time
import types


def run_prefill(args, rank: int, local_rank: int, world: int) -> None:
    import torch
    import torch.distributed as dist

    import bench as B
    from bench import ClockSampler, log
    from dnet_b200 import _cabi
    from dnet_b200.config import TransportSettings, get_settings
    from dnet_b200.shard.models import ShardLoadModelRequest
    from dnet_b200.shard.node import ApiNode, ShardNode
    from dnet_b200.shard.ring import even_split
    from dnet_b200.utils.model import SyntheticSource

    gloo = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        gloo = dist.new_group(backend="gloo")

    def barrier():
        if gloo is not None:
            dist.barrier(group=gloo)

    cfg = dict(B.LLAMA3_8B)
    L = args.layers or cfg["num_hidden_layers"]
    cfg["num_hidden_layers"] = L
    split = even_split(L, world)
    mine = split[rank]
    P, CH = int(args.prefill_len), int(args.prefill_chunk)
    K, Wm = args.steps, args.warmup
    lib = _cabi.load()

    os.environ["DNET_KV_CACHE_POOL_PAGES"] = str(2 * ((P + 64 + 63) // 64) + 8)     # two prompts' worth of KV pages
    get_settings.cache_clear()
    base_port = (int(os.environ.get("MASTER_PORT", "29500")) % 20000) + 33000
    ports = [base_port + 7 * r for r in range(world)]
    ts = TransportSettings()
    node = ShardNode(rank, ports[rank], transport_settings=ts, queue_size=max(128, P // CH + 8)).start()
    rt = node.runtime
    rt.kv_cache_config.max_tokens = P + 64
    nxt = None if world == 1 else types.SimpleNamespace(local_ip="127.0.0.1", shard_port=ports[(rank + 1) % world], instance="n")
    res = node.load_model(ShardLoadModelRequest(model_path=SyntheticSource(cfg, seed=0, layers=mine), total_layers=L, layers=mine,
                                                window_size=len(mine), residency_size=len(mine), kv_bits="fp16", next_node=nxt),
                          timeout=1800)
    assert res.success, res.message
    barrier()

    on_api = rank == 0
    api = None
    times, firsts = [], []
    sampler = ClockSampler(local_rank)
    tw0 = tw1 = time.perf_counter()
    if on_api:
        if world == 1:
            api = ApiNode(f"127.0.0.1:{ports[0]}", callback="local://")
            node.adapter.token_sink = api.token_sink
        else:
            api = ApiNode(f"127.0.0.1:{ports[0]}", callback="grpc", grpc_port=base_port + 7 * world + 3)
        api.manager.request_timeout_s = 600.0
        g = torch.Generator().manual_seed(1234)
        prompt = torch.randint(0, cfg["vocab_size"], (P,), generator=g).tolist()
        sampler.start()
        for i in range(Wm + K):
            if i == Wm:
                tw0 = time.perf_counter()
            t0 = time.perf_counter()
            out = api.generate(f"prefill{i}", prompt, 1, device_loop=False, logprobs=True, prefill_chunk=CH)
            dt = time.perf_counter() - t0
            firsts.append(out[0].token_id)
            if i >= Wm:
                times.append(dt)
            log(f"prompt {i}: {P} tokens in {dt * 1e3:.1f} ms -> {P / dt:,.0f} tok/s, first token {out[0].token_id}")
            time.sleep(0.05)        # let end_request release the KV before the next prompt claims pages
        tw1 = time.perf_counter()
    barrier()
    rt.compute_stream.synchronize()
    if on_api:
        assert len(set(firsts)) == 1, f"the same prompt gave different first tokens: {firsts}"
        mean = sum(times) / len(times)
        ad = node.adapter
        flops = 2.0 * P * (B.layer_bytes(cfg) / 2) * L + 4.0 * L * cfg["num_attention_heads"] * cfg["head_dim"] * P * P / 2
        out = {"metric": "prefill tok/s Llama-3-8B bf16, one long prompt streamed through the ring in chunks", "value": P / mean,
               "unit": "tok/s", "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": mean * 1e3, "higher_is_better": True,
               "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
               "config": {"workload": f"Llama-3-8B dims, {world} shards x {'/'.join(str(len(x)) for x in split)} layers, one {P}-token prompt "
                                      f"as {-(-P // CH)} chunks of {CH} tokens pipelined across the shards", "prompt_len": P, "chunk": CH,
                          "wire_dtype": "bf16", "kv": "fp16 paged",
                          "timing": "wall clock at the API, first frame sent -> first generated token received; a step = one prompt "
                                    "(fresh request, fresh KV); inputs far larger than L2 (weights 16 GB + KV)"},
               "e2e": {"value": P / mean, "unit": "tok/s", "h2d_bytes_per_step": 4 * P, "d2h_bytes_per_step": 8,
                       "api": "InferenceManager.generate_stream(prefill_chunk=C) over the ring transport"},
               "gpu_launches": int(lib.dn_launch_count()),
               "roofline": {"bound": "tensor", "achieved": flops / mean / 1e12 / world, "peak": B.tensor_peak_tf(), "unit": "TFLOP/s per GPU",
                            "frac": flops / mean / 1e12 / world / B.tensor_peak_tf(), "traffic": None,
                            "algorithmic_flops_per_prompt": flops,
                            "note": "2 x params x tokens for the projections + 4 x heads x head_dim x P^2 / 2 for causal attention"},
               "transport": {"frames_hop": ad.stats.get("frames_hop"), "frames_bytes": ad.stats.get("frames_bytes")},
               "first_token": firsts[0], "clocks": sampler.summary(tw0, tw1), "cpu_baseline": None}
        B.emit(out)
    sampler.stop()
    barrier()
    if api is not None:
        api.shutdown()
    node.unload_model()
    node.shutdown()
    if world > 1:
        dist.destroy_process_group()
