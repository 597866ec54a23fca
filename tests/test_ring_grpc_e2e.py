"""BASELINE config 1 end to end through the real transport classes, on CPU: two shards (stub layers)
and an API node in one process, three localhost gRPC servers, the rebuilt Shard / GrpcServicer /
RingAdapter / StreamManager on the shards and RingApiAdapter / ShardApiServicer / InferenceManager on
the API side.  No GPU -> no hop link -> tensor bytes ride the frames (the reference's path); the
device-hop variant of the same flow is tests/test_gpu_ring_adapter.py."""
import asyncio
import socket
import types

import numpy as np

from dnet_b200.api.grpc_servicer import ShardApiServer
from dnet_b200.api.inference import InferenceManager
from dnet_b200.api.strategies.ring import RingApiAdapter
from dnet_b200.config import TransportSettings
from dnet_b200.core.decoding.config import DecodingConfig
from dnet_b200.protos import dnet_ring_pb2 as pb
from dnet_b200.protos.dnet_ring_pb2_grpc import DnetRingServiceStub
from dnet_b200.shard.adapters.ring import RingAdapter
from dnet_b200.shard.grpc_servicer import GrpcServer
from dnet_b200.shard.shard import Shard
from tests.test_ring_plumbing import make_stub_shard


def free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_shard_ring_over_localhost_grpc_host_closed_loop():
    async def main():
        p0, p1, papi = free_port(), free_port(), free_port()
        rt0, rt1 = make_stub_shard("s0", [0, 1], 4), make_stub_shard("s1", [2, 3], 4)
        shards, servers = [], []
        for rt, port in ((rt0, p0), (rt1, p1)):
            ad = RingAdapter(rt, discovery=None, transport_settings=TransportSettings())
            sh = Shard(rt.shard_id, ad)
            srv = GrpcServer(port, sh, host="127.0.0.1")
            await sh.start(asyncio.get_running_loop())
            await srv.start()
            shards.append(sh)
            servers.append(srv)
        # topology: s0 -> s1 -> (ring closes at s0); the tail delivers tokens to the API callback
        await shards[0].adapter.configure_topology(types.SimpleNamespace(
            next_node=types.SimpleNamespace(local_ip="127.0.0.1", shard_port=p1), total_layers=4,
            api_callback_address=f"127.0.0.1:{papi}"))
        await shards[1].adapter.configure_topology(types.SimpleNamespace(
            next_node=types.SimpleNamespace(local_ip="127.0.0.1", shard_port=p0), total_layers=4,
            api_callback_address=f"127.0.0.1:{papi}"))
        api = RingApiAdapter()
        await api.start()
        await api.connect_first_shard("127.0.0.1", p0)
        im = InferenceManager(api, f"127.0.0.1:{papi}", request_timeout_s=10.0)
        api_srv = ShardApiServer(papi, im, host="127.0.0.1")
        await api_srv.start()

        toks = []
        async for res in im.generate_stream("req-1", [4, 5, 6], max_tokens=4, decoding=DecodingConfig(temperature=0.0),
                                            device_loop=False):
            toks.append(res.token_id)
        assert toks == [3, 3, 3, 3]                       # argmax(arange(4)) every step, 4 host-closed round trips
        a0, a1 = shards[0].adapter, shards[1].adapter
        assert a0.stats["frames_bytes"] == 4 and a0.stats["frames_hop"] == 0      # no hop link on CPU: bytes path
        assert a1.stats["tokens"] == 4
        # health / latency RPCs of the unchanged wire contract answer on the shard's server
        from grpc import aio as aio_grpc
        ch = aio_grpc.insecure_channel(f"127.0.0.1:{p1}")
        stub = DnetRingServiceStub(ch)
        h = await stub.HealthCheck(pb.HealthRequest(requester_id="t"))
        assert h.healthy and list(h.assigned_layers) == [2, 3]
        lat = await stub.MeasureLatency(pb.LatencyMeasureRequest(requester_id="t", payload_size=0))
        assert lat.success and lat.node_id == "s1"
        opened = await stub.SendActivation(pb.ActivationRequest(nonce="", activation=pb.Activation(dtype="b200.hop.open", layer_id=-1)))
        assert opened.success is False                     # no hop lanes on a CPU shard
        await ch.close()
        # a second request interleaves with nothing left over from the first (streams ended by end_request)
        toks2 = [r.token_id async for r in im.generate_stream("req-2", [1], max_tokens=2, device_loop=False)]
        assert toks2 == [3, 3]
        await api_srv.shutdown()
        await api.shutdown()
        for srv in servers:
            await srv.shutdown()
        for sh in shards:
            await sh.shutdown()

    asyncio.run(main())
