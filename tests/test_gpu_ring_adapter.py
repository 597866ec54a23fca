"""-m gpu: the product-side ring transport on a real device -- ShardNode (ShardRuntime + RingAdapter +
Shard + gRPC server) driven by the API-side token loop (ApiNode / InferenceManager).

* one shard, device-closed loop: leases -> schedule frames -> dn_shard_step_hop with the token loop
  closed through the shard's own lane slot, tokens observed through the TokenTap; bit-identical to the
  host-closed loop and to the oracle's golden greedy ids;
* several nonces in flight on one shard interleave without disturbing each other;
* two shards in one process: activations travel as device hops (metadata-only frames) and the result
  equals the one-shard run.  (Fused step-hop kernels of two shards cannot share ONE GPU -- each is a
  cooperative launch that fills the device and would spin on a flag its peer can never set -- so the
  multi-shard device-closed loop is exercised by `bench.py --gpus N`, one process per GPU.)"""
import socket
import time

import pytest
import torch

from tests.helpers import load_golden, oracle_weights

pytestmark = pytest.mark.gpu


def free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def make_node(cfgd, w, layers, shard_id, port, next_port=None, total=None, max_tokens=512):
    import types

    from dnet_b200.shard.models import ShardLoadModelRequest
    from dnet_b200.shard.node import ShardNode
    from dnet_b200.utils.model import HostDictSource

    node = ShardNode(shard_id, port).start()
    node.runtime.kv_cache_config.max_tokens = max_tokens
    nxt = None if next_port is None else types.SimpleNamespace(local_ip="127.0.0.1", shard_port=next_port, instance="next")
    req = ShardLoadModelRequest(model_path=HostDictSource(w, cfgd), total_layers=total or cfgd["num_hidden_layers"],
                                layers=list(layers), window_size=len(layers), residency_size=len(layers), kv_bits="fp16",
                                next_node=nxt)
    return node, req


@pytest.fixture(scope="module")
def tiny(cuda_lib):
    g = load_golden("tiny_llama")
    return g, oracle_weights(g["config"], g["wseed"])


def test_single_shard_device_closed_loop_matches_host_loop_and_golden(tiny):
    from dnet_b200.shard.node import ApiNode

    g, w = tiny
    cfgd = g["config"]
    port = free_port()
    node, req = make_node(cfgd, w, range(cfgd["num_hidden_layers"]), "s0", port)
    api = None
    try:
        assert node.load_model(req).success
        ad = node.adapter
        assert ad.hop is not None and ad.is_head and ad.is_tail
        api = ApiNode(f"127.0.0.1:{port}")
        ad.token_sink = api.token_sink
        steps = int(g["steps"])
        prompt = g["prompt"].tolist()
        dev = api.generate("dev", prompt, steps, device_loop=True, lease_steps=4, lease_ahead=2, logprobs=True)
        host = api.generate("host", prompt, steps, device_loop=False, logprobs=True)
        want = [int(t) for t in g["tokens"][:steps]]
        assert [r.token_id for r in host] == want
        assert [r.token_id for r in dev] == want, "device-closed loop must produce the host loop's tokens"
        assert [r.logprob for r in dev] == [r.logprob for r in host]      # same kernels, same bits
        assert ad.stats["frames_sched"] >= 1 and node.runtime.token_tap.delivered >= steps - 1
        assert node.runtime.step_errors == 0
    finally:
        if api is not None:
            api.shutdown()
        node.unload_model()
        node.shutdown()


def test_many_nonces_in_flight_do_not_disturb_each_other(tiny):
    from dnet_b200.shard.node import ApiNode

    g, w = tiny
    cfgd = g["config"]
    port = free_port()
    node, req = make_node(cfgd, w, range(cfgd["num_hidden_layers"]), "s0", port)
    api = None
    try:
        assert node.load_model(req).success
        api = ApiNode(f"127.0.0.1:{port}")
        node.adapter.token_sink = api.token_sink
        gen = torch.Generator().manual_seed(7)
        prompts = [g["prompt"].tolist()] + [torch.randint(0, cfgd["vocab_size"], (5 + 3 * i,), generator=gen).tolist() for i in range(5)]
        solo = [[r.token_id for r in api.generate(f"solo{i}", p, 10, device_loop=True)] for i, p in enumerate(prompts)]
        many = api.generate_many(prompts, 10, prefix="many", device_loop=True, lease_steps=3, lease_ahead=1)
        assert [[r.token_id for r in seq] for seq in many] == solo
        assert solo[0] == [int(t) for t in g["tokens"][:10]]
        # the end_of_request frames are processed by the shard after the API's generate returns
        t0 = time.time()
        while node.adapter._streams.lanes_in_use() and time.time() - t0 < 5.0:
            time.sleep(0.01)
        lanes = node.adapter._streams.lanes_in_use()
        assert not lanes, f"every request ended: no lane may stay claimed, got {lanes}"
    finally:
        if api is not None:
            api.shutdown()
        node.unload_model()
        node.shutdown()


def test_two_shards_one_process_activations_travel_as_device_hops(tiny):
    import asyncio

    from dnet_b200.shard.node import ApiNode

    g, w = tiny
    cfgd = g["config"]
    L = cfgd["num_hidden_layers"]
    p0, p1 = free_port(), free_port()
    n0, r0 = make_node(cfgd, w, range(0, L // 2), "s0", p0, next_port=p1, max_tokens=1024)
    n1, r1 = make_node(cfgd, w, range(L // 2, L), "s1", p1, next_port=p0, max_tokens=1024)
    api = None
    try:
        # in-process hop exchange: each adapter is handed its successor's HopLink object (no CUDA IPC)
        def exchange_for(me, peer):
            async def ex(own, k=0):          # the ring census asks for the endpoint k hops ahead
                node = peer if k % 2 == 0 else me
                for _ in range(600):
                    hop = node.runtime.hop or node.runtime.hop_pending
                    if hop is not None:
                        return hop
                    await asyncio.sleep(0.05)
                return None
            return ex
        n0.adapter.hop_exchange = exchange_for(n0, n1)
        n1.adapter.hop_exchange = exchange_for(n1, n0)
        import concurrent.futures as cf
        with cf.ThreadPoolExecutor(2) as ex:
            f0, f1 = ex.submit(n0.load_model, r0), ex.submit(n1.load_model, r1)
            assert f0.result().success and f1.result().success
        assert n0.adapter.hop is not None and n1.adapter.hop is not None
        api = ApiNode(f"127.0.0.1:{p0}")
        n1.adapter.token_sink = api.token_sink
        steps = int(g["steps"])
        out = api.generate("two", g["prompt"].tolist(), steps, device_loop=False, logprobs=True)
        assert [r.token_id for r in out] == [int(t) for t in g["tokens"][:steps]]
        assert n0.adapter.stats["frames_hop"] == steps and n0.adapter.stats["frames_bytes"] == 0
        # a longer prompt than one bulk slot falls back to bytes for that frame only
        long_prompt = torch.randint(0, cfgd["vocab_size"], (n0.adapter.bulk_tokens + 8,), generator=torch.Generator().manual_seed(3)).tolist()
        out2 = api.generate("long", long_prompt, 3, device_loop=False)
        assert len(out2) == 3 and n0.adapter.stats["frames_bytes"] == 1
        # chunked prefill: the same prompt as 64-token chunks (every chunk a device hop, only the last one sampled)
        hops_before = n0.adapter.stats["frames_hop"]
        out3 = api.generate("chunked", long_prompt, 3, device_loop=False, prefill_chunk=64)
        assert [r.token_id for r in out3] == [r.token_id for r in out2]
        n_chunks = -(-len(long_prompt) // 64)
        assert n0.adapter.stats["frames_hop"] - hops_before == n_chunks + 2 and n0.adapter.stats["frames_bytes"] == 1
    finally:
        if api is not None:
            api.shutdown()
        for n in (n0, n1):
            n.unload_model()
            n.shutdown()
