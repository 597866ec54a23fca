"""bench_swap.py -- BASELINE configs[3]: Llama-3-70B dims on 2 shards x 40 layers with an HBM cap, i.e. the
layer-swap path end to end (`bench.py --config swap`, one process per GPU under torchrun).

Every shard is a ShardNode (runtime + RingAdapter + gRPC server); the request loads with
``window_size = residency_size = W < 40`` so ``plan_policy`` selects ``offload``: W HBM slots are recycled, the 40
layer records (1.71 GB each) live in pinned host memory and each token swaps every layer in once on the prefetch
stream while the compute stream runs the resident window (OffloadPolicy + WeightCache + LayerManager +
dn_slot_prefetch).  Activations hop between the two shards over NVLink (metadata-only frames); the token loop is the
reference's host-closed loop (the API sends every token), because a shard that swaps layers cannot run the
persistent step kernel.  Reported: decode tok/s over K tokens by the API's wall clock (the host closes the loop here,
so the host IS on the token's critical path; there is no device-only interval that spans a token), host->HBM GB/s
against the box's own pinned-copy rate measured in the same run, and the share of a token not covered by the
compute the layers would need if they were resident."""
from __future__ import annotations

import json
import os
import time
import types

LLAMA3_70B = dict(hidden_size=8192, num_attention_heads=64, num_key_value_heads=8, head_dim=128, intermediate_size=28672,
                  vocab_size=128256, num_hidden_layers=80, rms_norm_eps=1e-5, rope_theta=500000.0, model_type="llama",
                  tie_word_embeddings=False, torch_dtype="bfloat16")


def run_swap(args, rank: int, local_rank: int, world: int) -> None:
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

    assert world >= 1
    gloo = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        gloo = dist.new_group(backend="gloo")

    def barrier():
        if gloo is not None:
            dist.barrier(group=gloo)

    cfg = dict(LLAMA3_70B)
    L = args.layers or cfg["num_hidden_layers"]
    cfg["num_hidden_layers"] = L
    H = cfg["hidden_size"]
    split = even_split(L, world)
    mine = split[rank]
    W = int(args.swap_window)
    K, Wm = args.steps, args.warmup
    lib = _cabi.load()
    layer_bytes = B.layer_bytes(cfg)

    # the box's own pinned host -> HBM copy rate (the roofline of this path)
    hb = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
    db = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    cs = torch.cuda.Stream()
    with torch.cuda.stream(cs):
        db.copy_(hb, non_blocking=True)
        cs.synchronize()
        t0 = time.perf_counter()
        for _ in range(4):
            db.copy_(hb, non_blocking=True)
        cs.synchronize()
    pcie = 4 * (1 << 30) / (time.perf_counter() - t0) / 1e9
    del hb, db

    os.environ["DNET_TOPOLOGY_RESIDENT_WINDOWS"] = str(args.swap_resident_windows)
    get_settings.cache_clear()
    base_port = (int(os.environ.get("MASTER_PORT", "29500")) % 20000) + 31000
    ports = [base_port + 7 * r for r in range(world)]
    node = ShardNode(rank, ports[rank], transport_settings=TransportSettings(), queue_size=64).start()
    rt = node.runtime
    rt.kv_cache_config.max_tokens = B.PROMPT_LEN + Wm + K + 64
    nxt = None if world == 1 else types.SimpleNamespace(local_ip="127.0.0.1", shard_port=ports[(rank + 1) % world], instance="n")
    t0 = time.perf_counter()
    res = node.load_model(ShardLoadModelRequest(model_path=SyntheticSource(cfg, seed=0, layers=mine, share_layers=True),
                                                total_layers=L, layers=mine, window_size=W, residency_size=W, kv_bits="fp16",
                                                next_node=nxt), timeout=3600)
    assert res.success, res.message
    pol = rt.policy
    assert pol._mode == "offload", pol._mode
    load_s = time.perf_counter() - t0
    log(f"rank {rank}: layers {mine[0]}..{mine[-1]} staged in pinned host memory ({len(mine) * layer_bytes / 1e9:.1f} GB) in {load_s:.1f}s; "
        f"HBM slots {pol.weight_cache.max_weights} x {layer_bytes / 1e9:.2f} GB; pinned copy rate {pcie:.1f} GB/s")
    barrier()

    # the API sits beside the head shard (rank 0); the finalising shard returns tokens in process when it is the same
    # process, else over the reference's own SendToken RPC (shardapi proto)
    on_api = rank == 0
    got = []
    api = None
    if on_api:
        api_port = base_port + 7 * world + 3
        if world == 1:
            api = ApiNode(f"127.0.0.1:{ports[0]}", callback="local://")
            node.adapter.token_sink = api.token_sink
        else:
            api = ApiNode(f"127.0.0.1:{ports[0]}", callback="grpc", grpc_port=api_port)
    g = torch.Generator().manual_seed(1234)
    prompt = torch.randint(0, cfg["vocab_size"], (B.PROMPT_LEN,), generator=g).tolist()
    stream = rt.compute_stream
    sampler = ClockSampler(local_rank)
    wall = 0.0
    tw0 = tw1 = time.perf_counter()
    if on_api:
        async def run():
            nonlocal wall, tw0, tw1
            i = 0
            async for r in api.manager.generate_stream("swap", prompt, 1 + Wm + K, device_loop=False, logprobs=True):
                got.append(r.token_id)
                i += 1
                if i == 1 + Wm:
                    tw0 = time.perf_counter()
            tw1 = time.perf_counter()
            wall = tw1 - tw0
        sampler.start()
        api.call(run(), timeout=3600)
    # ranks that do not host the API just serve; completion is signalled by the barrier below
    barrier()
    stream.synchronize()
    # per-rank accounting: every local layer is swapped in once per token
    lm = pol.weight_cache.layer_manager
    ms_tok = wall / K * 1e3 if on_api else 0.0
    if gloo is not None:
        t = torch.tensor([ms_tok], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=gloo)
        ms_tok = float(t.item())
    swapped = len(mine) if pol.weight_cache.max_weights < len(mine) else 0
    per_rank_gbs = swapped * layer_bytes / (ms_tok / 1e3) / 1e9 if ms_tok else 0.0
    if rank == 0:
        # one sequence: the two shards swap one after the other, so each GPU's copy engine is busy ~1/world of a token
        out = {"metric": "decode tok/s Llama-3-70B bf16 bs=1, layer swap (BASELINE configs[3])", "value": 1e3 / ms_tok, "unit": "tok/s",
               "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms_tok, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
               "config": {"workload": f"Llama-3-70B dims, {world} shards x {'/'.join(str(len(x)) for x in split)} layers, HBM cap: "
                                      f"window {W} x resident windows {args.swap_resident_windows} -> {pol.weight_cache.max_weights} HBM layer slots "
                                      f"of {layer_bytes / 1e9:.2f} GB per shard, every layer swapped in from pinned host memory each token "
                                      f"(policy '{pol._mode}')", "prompt_len": B.PROMPT_LEN, "wire_dtype": "bf16", "kv": "fp16 paged",
                          "token_loop": "host-closed (API sends every token); activations hop over NVLink as metadata-only frames",
                          "timing": "wall clock at the API over K tokens after W warm-up tokens (host-closed loop)"},
               "e2e": {"value": 1e3 / ms_tok, "unit": "tok/s", "h2d_bytes_per_step": world * len(mine) * layer_bytes + 4,
                       "d2h_bytes_per_step": 8, "api": "InferenceManager.generate_stream(device_loop=False) over the ring transport"},
               "gpu_launches": int(lib.dn_launch_count()),
               "roofline": {"bound": "pcie", "achieved": world * len(mine) * layer_bytes / (ms_tok / 1e3) / 1e9, "peak": pcie,
                            "unit": "GB/s host->HBM (one GPU's copy engine active at a time: one sequence, shards take turns)",
                            "frac": world * len(mine) * layer_bytes / (ms_tok / 1e3) / 1e9 / pcie, "traffic": None,
                            "peak_source": "pinned cudaMemcpyAsync host->device measured in this run (4 x 1 GiB)",
                            "algorithmic_bytes_per_token": world * len(mine) * layer_bytes},
               "swap": {"layer_bytes": layer_bytes, "hbm_slots_per_shard": pol.weight_cache.max_weights, "layers_per_shard": len(mine),
                        "load_s": load_s, "record_reads": dict(lm.record_reads),
                        "compute_ms_per_token_if_resident": world * len(mine) * layer_bytes / 6572.2e6,
                        "stall_fraction": 1.0 - (world * len(mine) * layer_bytes / 6572.2e6) / ms_tok},
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
