"""bench_moe.py -- BASELINE configs[4]: Mixtral-8x7B dims (32 layers, 8 experts, top-2), bf16, N shards, several
requests in flight (`bench.py --config moe`, one process per GPU under torchrun).

Every shard is a ShardNode; the model class is ``mixtral`` (dnet_b200/core/models/llama.py MixtralRingModel): attention
as in llama, the FFN a sparse MoE block whose expert selection happens on the device, so the layers run on the per-op
path (router GEMV -> top-k select -> expert gate/up + down GEMVs through a device pointer table; one CUDA graph per
decode step and shard) and the token loop is the reference's host-closed loop: the API sends every token, activations
hop between shards over NVLink as metadata-only frames, the last shard returns tokens over SendToken.  ``--in-flight``
requests (default 8, the config's "bs=8") decode concurrently, each bs=1, so up to N shards work at once.

Reported: aggregate decode tok/s by the API's wall clock over K steps of every request after W warm-up steps (the host is
on the token's critical path in this loop by construction), and the HBM roofline of the bytes a token actually touches:
attention weights + router + 2 of 8 experts per layer (+ lm_head)."""
from __future__ import annotations

import os
import time
import types

MIXTRAL_8X7B = dict(hidden_size=4096, num_attention_heads=32, num_key_value_heads=8, head_dim=128, intermediate_size=14336,
                    vocab_size=32000, num_hidden_layers=32, rms_norm_eps=1e-5, rope_theta=1000000.0, model_type="mixtral",
                    tie_word_embeddings=False, torch_dtype="bfloat16", num_local_experts=8, num_experts_per_tok=2)


def run_moe(args, rank: int, local_rank: int, world: int) -> None:
    import asyncio

    import torch
    import torch.distributed as dist

    import bench as B
    from bench import ClockSampler, log
    from dnet_b200 import _cabi
    from dnet_b200.config import TransportSettings
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

    cfg = dict(MIXTRAL_8X7B)
    L = args.layers or cfg["num_hidden_layers"]
    cfg["num_hidden_layers"] = L
    H, F, E, k, V = cfg["hidden_size"], cfg["intermediate_size"], cfg["num_local_experts"], cfg["num_experts_per_tok"], cfg["vocab_size"]
    split = even_split(L, world)
    mine = split[rank]
    K, Wm = args.steps, args.warmup
    NS = args.in_flight or 8
    lib = _cabi.load()
    qd, kd = cfg["num_attention_heads"] * 128, cfg["num_key_value_heads"] * 128
    attn_bytes = 2 * (qd * H + 2 * kd * H + H * qd) + 4 * H
    layer_token_bytes = attn_bytes + 2 * E * H + k * 3 * F * H * 2          # what one token reads in one layer
    token_bytes = L * layer_token_bytes + 2 * V * H + 2 * H + 2 * H

    base_port = (int(os.environ.get("MASTER_PORT", "29500")) % 20000) + 35000
    ports = [base_port + 7 * r for r in range(world)]
    node = ShardNode(rank, ports[rank], transport_settings=TransportSettings(), queue_size=128).start()
    rt = node.runtime
    rt.kv_cache_config.max_tokens = B.PROMPT_LEN + Wm + K + 64
    nxt = None if world == 1 else types.SimpleNamespace(local_ip="127.0.0.1", shard_port=ports[(rank + 1) % world], instance="n")
    t0 = time.perf_counter()
    res = node.load_model(ShardLoadModelRequest(model_path=SyntheticSource(cfg, seed=0, layers=mine), total_layers=L, layers=mine,
                                                window_size=len(mine), residency_size=len(mine), kv_bits="fp16", next_node=nxt),
                          timeout=3600)
    assert res.success, res.message
    assert rt.use_megakernel is False
    log(f"rank {rank}: layers {mine[0]}..{mine[-1]} resident ({len(mine) * (attn_bytes + 2 * E * H + E * 3 * F * H * 2) / 1e9:.1f} GB) "
        f"in {time.perf_counter() - t0:.1f}s")
    barrier()

    on_api = rank == 0
    api = None
    sampler = ClockSampler(local_rank)
    tw0 = tw1 = time.perf_counter()
    toks = {}
    if on_api:
        if world == 1:
            api = ApiNode(f"127.0.0.1:{ports[0]}", callback="local://")
            node.adapter.token_sink = api.token_sink
        else:
            api = ApiNode(f"127.0.0.1:{ports[0]}", callback="grpc", grpc_port=base_port + 7 * world + 3)
        g = torch.Generator().manual_seed(1234)
        prompts = [torch.randint(0, V, (B.PROMPT_LEN,), generator=g).tolist() for _ in range(NS)]
        marks = {}

        async def one(i):
            out = []
            n = 0
            async for r in api.manager.generate_stream(f"moe{i}", prompts[i], 1 + Wm + K, device_loop=False, logprobs=True):
                out.append(r.token_id)
                n += 1
                if n == 1 + Wm:
                    marks.setdefault("t0", []).append(time.perf_counter())
            marks.setdefault("t1", []).append(time.perf_counter())
            toks[i] = out

        async def run():
            await asyncio.gather(*[one(i) for i in range(NS)])
        sampler.start()
        api.call(run(), timeout=3600)
        tw0, tw1 = min(marks["t0"]), max(marks["t1"])
    barrier()
    rt.compute_stream.synchronize()
    if on_api:
        assert all(len(v) == 1 + Wm + K and min(v) >= 0 for v in toks.values()), {i: len(v) for i, v in toks.items()}
        wall = tw1 - tw0
        tps = NS * K / wall
        peak, peak_src = B.peaks()
        # steady state bound: with NS >= world requests in flight every shard streams its layers' bytes per token
        bound = 1.0 / (max(len(x) for x in split) * layer_token_bytes / (peak * 1e9))
        out = {"metric": "decode tok/s Mixtral-8x7B dims bf16, bs=1 per request (BASELINE configs[4])", "value": tps, "unit": "tok/s",
               "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": wall / K * 1e3, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
               "config": {"workload": f"Mixtral-8x7B dims (32 layers, 8 experts top-2), {world} shards x {'/'.join(str(len(x)) for x in split)} "
                                      f"layers, {NS} requests in flight, each bs=1; sparse MoE layers on the per-op path (CUDA graph per step)",
                          "prompt_len": B.PROMPT_LEN, "wire_dtype": "bf16", "kv": "fp16 paged", "sequences_in_flight": NS,
                          "token_loop": "host-closed (the API sends every token); activations hop over NVLink as metadata-only frames",
                          "timing": "wall clock at the API from the last warm-up token of the first request to the last token of the last; "
                                    "inputs larger than L2 (>= 0.8 GB of weights per layer and token)"},
               "e2e": {"value": tps, "unit": "tok/s", "h2d_bytes_per_step": 4 * NS, "d2h_bytes_per_step": 8 * NS,
                       "api": "InferenceManager.generate_stream(device_loop=False) x in-flight requests over the ring transport"},
               "gpu_launches": int(lib.dn_launch_count()),
               "roofline": {"bound": "hbm", "achieved": token_bytes * tps / 1e9 / world, "peak": peak, "unit": "GB/s per GPU",
                            "frac": token_bytes * tps / 1e9 / world / peak, "traffic": None, "peak_source": peak_src,
                            "algorithmic_bytes_per_token": token_bytes, "layer_bytes_per_token": layer_token_bytes,
                            "pipeline_bound_tok_s": bound,
                            "note": "bytes one token touches: attention weights + router + 2 of 8 experts per layer, + lm_head; "
                                    "frac = aggregate bytes/s per GPU over the copy peak"},
               "check": {"first_tokens": [toks[i][0] for i in range(NS)], "last_tokens": [toks[i][-1] for i in range(NS)]},
               "clocks": sampler.summary(tw0, tw1), "cpu_baseline": None}
        B.emit(out)
    sampler.stop()
    barrier()
    if api is not None:
        api.shutdown()
    node.unload_model()
    node.shutdown()
    if world > 1:
        dist.destroy_process_group()
