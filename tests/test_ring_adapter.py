"""RingAdapter / StreamManager / Shard / servicer on CPU (no GPU).

The first half restates the assertions the reference holds for this layer
(tests/subsystems/test_ring_adapter.py:43-513, tests/test_stream_manager.py:13-59,
tests/subsystems/test_shard.py) against the rebuilt classes, mlx-free.  The second half covers
what the rebuild adds: metadata frames, lane bookkeeping, the head shard's decode schedule, and a
real two-shard ring over localhost gRPC driven by the API-side adapter (BASELINE config 1)."""
import asyncio
import queue as pyq
import types
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest
import torch

from dnet_b200.config import TransportSettings
from dnet_b200.core.stream_manager import StreamManager
from dnet_b200.core.types.messages import ActivationMessage, TokenResult
from dnet_b200.protos import dnet_ring_pb2 as pb
from dnet_b200.shard import frames as fr
from dnet_b200.shard.adapters import ring as ring_mod
from dnet_b200.shard.adapters.ring import RingAdapter
from dnet_b200.shard.models import ShardLoadModelRequest
from dnet_b200.shard.shard import Shard


# ---------------------------------------------------------------------------- fakes
class FakeRuntimeForAdapter:
    def __init__(self, shard_id="S1", max_queue_size=8, assigned_next=()):
        self.shard_id = shard_id
        self.max_queue_size = max_queue_size
        self.executor = ThreadPoolExecutor(max_workers=1)
        self.activation_recv_queue = pyq.Queue(maxsize=max_queue_size)
        self.activation_send_queue = pyq.Queue(maxsize=max_queue_size)
        self._assigned_set = set(assigned_next)
        self._wire_dtype_str = "float16"
        self._wire_mx_dtype = torch.float16
        self._kv_calls = []
        self.input_pool = None
        self.output_pool = None

    def get_or_make_kv(self, nonce):
        self._kv_calls.append(nonce)
        return []


class FakeChannel:
    def __init__(self, addr):
        self.addr, self.closed = addr, False

    async def close(self):
        self.closed = True


class FakeRingStub:
    def __init__(self, ch):
        self._ch = ch

    def StreamActivations(self, it):
        async def gen():
            if False:
                yield None
        return gen()


class FakeApiStub:
    def __init__(self, ch):
        self._ch, self.sent = ch, []

    async def SendToken(self, req, timeout=3.0):
        self.sent.append(req)
        return types.SimpleNamespace(success=True, message="ok")


class FakeStreamAck:
    def __init__(self, *, accepted=True, message=""):
        self.accepted, self.message = bool(accepted), str(message)


class FakeStreamCall:
    def __init__(self, acks):
        self._acks, self._closed = list(acks), False

    def __aiter__(self):
        return self

    async def __anext__(self):
        if self._acks:
            return self._acks.pop(0)
        raise StopAsyncIteration

    async def aclose(self):
        self._closed = True


def node(port=9002, ip="10.0.0.2"):
    return types.SimpleNamespace(instance="S2", local_ip=ip, shard_port=port, server_port=8002)


async def wait_until(cond, timeout=0.5, interval=0.01):
    end = asyncio.get_running_loop().time() + timeout
    while True:
        try:
            if cond():
                return True
        except Exception:
            pass
        if asyncio.get_running_loop().time() >= end:
            return False
        await asyncio.sleep(interval)


@pytest.fixture
def grpc_ok(monkeypatch):
    seen = {}
    monkeypatch.setattr(ring_mod, "aio_grpc", types.SimpleNamespace(
        insecure_channel=lambda addr, options=None: (seen.__setitem__("addr", addr) or FakeChannel(addr))))
    monkeypatch.setattr(ring_mod, "_make_ring_stub", lambda ch: FakeRingStub(ch))
    monkeypatch.setattr(ring_mod, "_make_api_stub", lambda ch: FakeApiStub(ch))
    return seen


def make_adapter(assigned_next=(), streaming=True):
    rt = FakeRuntimeForAdapter(assigned_next=assigned_next)
    ad = RingAdapter(runtime=rt, discovery=None, transport_settings=TransportSettings(streaming=streaming))
    return ad, rt


def load_req(next_node=None, layers=(0,), total=2):
    return ShardLoadModelRequest(model_path="m", total_layers=total, layers=list(layers), warmup=False, next_node=next_node,
                                 window_size=1, residency_size=1, kv_bits="8bit", api_callback_address="")


def act_request(nonce, layer_id=0, dtype="float32", data=b"\x00\x00\x80?"):
    return pb.ActivationRequest(nonce=nonce, activation=pb.Activation(data=data, batch_size=1, shape=[1], dtype=dtype,
                                                                      layer_id=layer_id),
                                timestamp=0, node_origin="S", callback_url="cb")


def amsg(nonce, final=False, cb=""):
    m = ActivationMessage(nonce=nonce, pool_id=0, batch_size=1, shape=(1,), dtype="float32", layer_id=0, timestamp=0,
                          node_origin="S", callback_url=cb)
    m.is_final = final
    return m


# ---------------------------------------------------------------------------- StreamManager (reference tests/test_stream_manager.py)
def test_stream_backpressure_disables_and_reenables():
    sm = StreamManager(idle_timeout_s=10.0, backoff_s=0.05)

    async def main():
        factory = lambda _it: FakeStreamCall([FakeStreamAck(accepted=False, message="backpressure: slow")])
        ctx = await sm.get_or_create_stream("n1", factory)
        assert ctx is not None and ctx.open is True
        await asyncio.sleep(0)
        assert ctx.disabled is True and ctx.disabled_until > 0
        ctx2 = await sm.get_or_create_stream("n1", factory)
        assert ctx2 is ctx and ctx2.disabled is True
        await asyncio.sleep(0.06)
        ctx3 = await sm.get_or_create_stream("n1", factory)
        assert ctx3 is ctx and ctx3.disabled is False
        await sm.end_stream("n1")
        assert sm.get_ctx("n1") is None

    asyncio.run(main())


def test_stream_idle_cleanup_closes_context():
    sm = StreamManager(idle_timeout_s=0.01, backoff_s=0.01)

    async def main():
        ctx = await sm.get_or_create_stream("n2", lambda _it: FakeStreamCall([]))
        assert ctx is not None
        sm.get_ctx("n2").last_activity_t = asyncio.get_running_loop().time() - 1.0
        assert await sm.cleanup_idle_streams() == 1
        assert sm.get_ctx("n2") is None

    asyncio.run(main())


def test_stream_manager_lanes_are_exclusive_and_drain_before_reuse():
    sm = StreamManager(n_lanes=2)
    a, b = sm.claim_lane("a"), sm.claim_lane("b")
    assert {a.lane, b.lane} == {0, 1} and sm.claim_lane("c") is None          # all lanes taken
    assert sm.claim_lane("a") is a                                            # idempotent per nonce
    busy = types.SimpleNamespace(done=False)
    busy.query = lambda: busy.done
    sm.note_scheduled(a, 7, tail_event=busy)
    lane_a = sm.release_lane("a")
    assert sm.claim_lane("c") is None, "a released lane is not reusable while its last kernel is still queued"
    busy.done = True
    c = sm.claim_lane("c")
    assert c is not None and c.lane == lane_a and c.base_seq == 7             # flags only grow: new owner starts above
    # a follower shard claims exactly the lane the head named, and refuses a lane another nonce holds
    sm2 = StreamManager(n_lanes=4)
    assert sm2.claim_lane("x", 3).lane == 3 and sm2.claim_lane("y", 3) is None
    assert sm2.lanes_in_use() == {"x": 3}


# ---------------------------------------------------------------------------- adapter: reference assertions
def test_workers_start_and_shutdown():
    ad, rt = make_adapter()

    async def main():
        await ad.start()
        assert ad.running is True and len(ad._tasks) >= 4
        await ad.shutdown()
        assert ad.running is False

    asyncio.run(main())


def test_configure_topology_connects_next_node(grpc_ok):
    ad, rt = make_adapter()

    async def main():
        await ad.configure_topology(load_req(next_node=node()))
        assert ad.next_node_stub is not None and isinstance(ad.next_node_channel, FakeChannel)
        assert ad.next_node_channel.addr.endswith(":9002") and grpc_ok["addr"] == "10.0.0.2:9002"
        assert ad.hop is None                       # no model / no CUDA here: the bytes path stays in force

    asyncio.run(main())


def test_configure_topology_without_next_node():
    ad, rt = make_adapter()

    async def main():
        await ad.configure_topology(load_req(next_node=None))
        assert ad.next_node is None and ad.next_node_stub is None

    asyncio.run(main())


def test_reset_topology_closes_channels():
    ad, rt = make_adapter()
    ch1, ch2 = FakeChannel("10.0.0.2:9002"), FakeChannel("127.0.0.1:5050")
    ad.next_node_channel, ad.api_channel, ad.api_stub, ad.api_address = ch1, ch2, FakeApiStub(ch2), "127.0.0.1:5050"

    async def main():
        await ad.reset_topology()
        assert ad.next_node_channel is None and ad.next_node_stub is None and ch1.closed
        assert ad.api_channel is None and ad.api_stub is None and ad.api_address is None and ch2.closed

    asyncio.run(main())


def test_ingress_local_layer_is_deserialised_and_queued(monkeypatch):
    ad, rt = make_adapter(assigned_next={1})
    fake = amsg("n")
    monkeypatch.setattr(ad.codec, "deserialize", lambda req: fake)

    async def main():
        await ad.start()
        await ad.ingress_q.put(act_request("n", layer_id=0))
        assert await wait_until(lambda: not rt.activation_recv_queue.empty())
        assert rt.activation_recv_queue.get_nowait() is fake and rt._kv_calls == ["n"]
        await ad.shutdown()

    asyncio.run(main())


def test_ingress_forwards_frames_for_other_shards(grpc_ok):
    ad, rt = make_adapter(assigned_next=set())
    ad.next_node = node()

    async def main():
        await ad.start()
        await ad.ingress_q.put(act_request("n2"))
        assert await wait_until(lambda: ad._streams.get_ctx("n2") is not None and ad._streams.get_ctx("n2").last_seq >= 1)
        frame = await ad._streams.get_ctx("n2").queue.get()
        assert frame.request.nonce == "n2" and frame.seq == 1 and rt.activation_recv_queue.empty()
        await ad.shutdown()

    asyncio.run(main())


def test_ingress_deserialize_exception_is_handled(monkeypatch):
    ad, rt = make_adapter(assigned_next={1})

    def boom(req):
        raise RuntimeError("decode")

    monkeypatch.setattr(ad.codec, "deserialize", boom)

    async def main():
        await ad.start()
        await ad.ingress_q.put(act_request("n3"))
        assert await wait_until(lambda: ad.ingress_q.empty()) and rt.activation_recv_queue.empty()
        await ad.shutdown()

    asyncio.run(main())


def test_egress_routes_final_and_non_final(monkeypatch):
    ad, rt = make_adapter()
    calls = {"act": 0, "tok": 0}

    async def fake_act(msg):
        calls["act"] += 1

    async def fake_tok(msg):
        calls["tok"] += 1

    monkeypatch.setattr(ad, "_send_activation", fake_act)
    monkeypatch.setattr(ad, "_send_token", fake_tok)

    async def main():
        await ad.start()
        rt.activation_send_queue.put(amsg("a", final=False))
        rt.activation_send_queue.put(amsg("b", final=True))
        assert await wait_until(lambda: calls == {"act": 1, "tok": 1}, timeout=2.0)
        await ad.shutdown()

    asyncio.run(main())


def test_send_activation_serialises_and_enqueues(grpc_ok):
    ad, rt = make_adapter()
    ad.next_node = node()

    async def main():
        await ad._connect_next_node()
        ctx = await ad._streams.get_or_create_stream("x", ad.next_node_stub.StreamActivations)
        msg = amsg("x")
        msg.tensor = torch.tensor([1.0], dtype=torch.float32)
        await ad._send_activation(msg)
        frame = await ctx.queue.get()
        assert frame is not None and frame.request.activation.dtype == rt._wire_dtype_str
        assert frame.request.activation.data == np.float16(1.0).tobytes()        # cast to the wire dtype, raw bytes
        assert msg.tensor is None and msg.dtype == rt._wire_dtype_str

    asyncio.run(main())


def test_send_activation_without_stub_is_a_noop():
    ad, rt = make_adapter()

    async def main():
        await ad._send_activation(amsg("x"))
        assert ad._streams.get_ctx("x") is None

    asyncio.run(main())


def test_forward_with_streaming_disabled():
    ad, rt = make_adapter(streaming=False)

    async def main():
        await ad._forward_activation(act_request("q1"))
        assert ad._streams.get_ctx("q1") is None

    asyncio.run(main())


def test_send_token_callback_url_then_fallback_address(grpc_ok):
    ad, rt = make_adapter()

    async def main():
        m1 = amsg("t1", final=True, cb="grpc://127.0.0.1:5050")
        m1.token_id = 7
        await ad._send_token(m1)
        assert ad.api_address == "127.0.0.1:5050" and ad.api_stub.sent[-1].token_id == 7
        m2 = amsg("t2", final=True)
        m2.token_id = 8
        ad.api_callback_address = "10.0.0.9:5055"
        await ad._send_token(m2)
        assert ad.api_address == "10.0.0.9:5055" and ad.api_stub.sent[-1].nonce == "t2"

    asyncio.run(main())


def test_send_token_invalid_url_and_no_fallback():
    ad, rt = make_adapter()

    async def main():
        await ad._send_token(amsg("b1", final=True, cb="http://bad"))
        assert ad.api_address is None

    asyncio.run(main())


def test_send_token_channel_creation_failure(monkeypatch):
    ad, rt = make_adapter()

    def boom(addr, options=None):
        raise RuntimeError("boom")

    monkeypatch.setattr(ring_mod, "aio_grpc", types.SimpleNamespace(insecure_channel=boom))

    async def main():
        ad.api_callback_address = "10.0.0.1:5050"
        await ad._send_token(amsg("b2", final=True))
        assert ad.api_channel is None

    asyncio.run(main())


def test_admit_frame_drops_when_not_running():
    ad, rt = make_adapter()

    async def main():
        await ad.admit_frame(act_request("z"))
        assert ad.ingress_q.empty()

    asyncio.run(main())


def test_reconnect_next_node_replaces_channel(grpc_ok):
    ad, rt = make_adapter()
    old = FakeChannel("10.0.0.2:9002")
    ad.next_node_channel, ad.next_node_stub, ad.next_node = old, FakeRingStub(old), node()

    async def main():
        assert await ad._reconnect_next_node() is True
        assert old.closed is True and isinstance(ad.next_node_channel, FakeChannel) and ad.next_node_channel is not old

    asyncio.run(main())


# ---------------------------------------------------------------------------- frames
def test_frame_headers_round_trip_and_reject_garbage():
    m = fr.HopMeta(lane=5, seq=77, seq0=12, flags=1)
    assert fr.unpack_hop(fr.pack_hop(m)) == m
    ent = [(0, 3), (7, 9), (2, 4000000000)]
    assert fr.unpack_sched(fr.pack_sched(ent)) == ent and fr.unpack_sched(fr.pack_sched([])) == []
    assert fr.unpack_lease(fr.pack_lease(16)) == (16, None) and fr.unpack_lease(fr.pack_lease(3, 42)) == (3, 42)
    assert fr.is_hop(fr.hop_dtype("bfloat16")) and fr.hop_wire_dtype(fr.hop_dtype("bfloat16")) == "bfloat16"
    for bad in (b"", b"XXXX" + b"\0" * 16, fr.pack_hop(m)[:-1]):
        with pytest.raises(ValueError):
            fr.unpack_hop(bad)
    with pytest.raises(ValueError):
        fr.unpack_sched(fr.pack_sched(ent)[:-4])
    # a metadata frame is an ordinary ActivationRequest on the wire (the proto is unchanged)
    req = pb.ActivationRequest(nonce="n", activation=pb.Activation(data=fr.pack_hop(m), batch_size=1, shape=[1, 4, 8],
                                                                 dtype=fr.hop_dtype("bfloat16"), layer_id=3))
    back = pb.ActivationRequest.FromString(req.SerializeToString())
    assert fr.unpack_hop(back.activation.data) == m and list(back.activation.shape) == [1, 4, 8]


# ---------------------------------------------------------------------------- decode schedule (head shard)
def test_head_scheduler_merges_leases_round_robin_in_lane_order(grpc_ok):
    ad, rt = make_adapter(assigned_next={0})
    rt._assigned_set = {0, 1}
    ad.total_layers = 4                       # layers 2,3 live on the next shard: head but not tail
    ad.next_node = node()
    ad.rounds_per_frame, ad.sched_depth = 2, 8
    ad._streams.configure_lanes(4)
    for nonce in ("a", "b"):
        ctx = ad._streams.claim_lane(nonce)
        ctx.params["seq0"] = 1
    lane = {n: ad._streams.lane_ctx(n).lane for n in ("a", "b")}

    async def main():
        await ad.start()
        ad.lease("a", 3)
        ad.lease("b", 2)
        assert await wait_until(lambda: rt.activation_recv_queue.qsize() >= 2, timeout=2.0)
        got = []
        while not rt.activation_recv_queue.empty():
            m = rt.activation_recv_queue.get_nowait()
            assert m.dtype == fr.SCHED_DTYPE
            got += m.sched
            m.sched_done.record(None)
        order = sorted(lane.values())
        la, lb = lane["a"], lane["b"]
        # rounds: (a1 b1) (a2 b2) (a3): one step per leased nonce per round, lanes ascending, seq grows per lane
        exp = [(l, 1) for l in order] + [(l, 2) for l in order] + [(la, 3)]
        assert got == exp and not ad._leases
        # the same frames, byte-identical entries, were forwarded to the next shard in the same order
        ctx = ad._streams.get_ctx(f"sched-{rt.shard_id}")
        fwd = []
        while not ctx.queue.empty():
            fwd += fr.unpack_sched((await ctx.queue.get()).request.activation.data)
        assert fwd == exp
        await ad.shutdown()

    asyncio.run(main())


def test_scheduler_bounds_frames_in_flight(grpc_ok):
    ad, rt = make_adapter(assigned_next={0})
    rt._assigned_set = {0}
    rt.activation_recv_queue = pyq.Queue(maxsize=64)
    ad.total_layers = 1                       # single shard: head and tail
    ad.rounds_per_frame, ad.sched_depth = 1, 2
    ad._streams.configure_lanes(2)
    ad._streams.claim_lane("a").params["seq0"] = 5

    async def main():
        await ad.start()
        ad.lease("a", 6)
        assert await wait_until(lambda: rt.activation_recv_queue.qsize() == 2, timeout=2.0)
        await asyncio.sleep(0.05)
        assert rt.activation_recv_queue.qsize() == 2, "only sched_depth frames may be outstanding"
        first = rt.activation_recv_queue.get_nowait()
        assert first.sched == [(ad._streams.lane_ctx("a").lane, 5)]
        first.sched_done.record(None)         # the compute thread finished that frame
        assert await wait_until(lambda: rt.activation_recv_queue.qsize() == 2, timeout=2.0)
        await ad.shutdown()

    asyncio.run(main())


def test_follower_enqueues_schedule_and_forwards_unless_tail(grpc_ok):
    for total, expect_forward in ((4, True), (2, False)):
        ad, rt = make_adapter(assigned_next={1})
        rt._assigned_set = {1}
        ad.total_layers = total
        ad.next_node = node()
        ent = [(1, 4), (0, 9)]

        async def main():
            await ad.start()
            await ad.ingress_q.put(pb.ActivationRequest(nonce="sched-S0", activation=pb.Activation(
                data=fr.pack_sched(ent), batch_size=1, shape=[2], dtype=fr.SCHED_DTYPE, layer_id=-1)))
            assert await wait_until(lambda: not rt.activation_recv_queue.empty())
            assert rt.activation_recv_queue.get_nowait().sched == ent
            await asyncio.sleep(0.02)
            assert (ad._streams.get_ctx("sched-S0") is not None) == expect_forward
            await ad.shutdown()

        asyncio.run(main())


# ---------------------------------------------------------------------------- Shard glue (reference tests/subsystems/test_shard.py)
def test_shard_admit_frame_and_load_unload():
    class Rt(FakeRuntimeForAdapter):
        def __init__(self):
            super().__init__()
            self.assigned_layers, self.model_path, self.model = [], None, None
            self.started = False

        def attach_loop(self, loop):
            self._loop = loop

        def start(self):
            self.started = True

        def shutdown(self):
            self.started = False

        def queue_size(self):
            return self.activation_recv_queue.qsize()

        def load_model_core(self, req):
            self.assigned_layers, self.model_path = list(req.layers), req.model_path
            self._assigned_set = set(req.layers)

        def unload_model_core(self):
            from dnet_b200.shard.models import ShardUnloadModelResponse
            return ShardUnloadModelResponse(success=True, message="ok")

    rt = Rt()
    ad = RingAdapter(runtime=rt, discovery=None, transport_settings=TransportSettings())
    sh = Shard("S1", ad)

    async def main():
        await sh.admit_frame(act_request("dropped"))
        assert ad.ingress_q.empty()                       # adapter not running: dropped
        await sh.start(asyncio.get_running_loop())
        assert rt.started and ad.running
        res = await sh.load_model(load_req(layers=(0, 1), total=2))
        assert res.success and res.layers_loaded == [0, 1] and ad.total_layers == 2 and ad.is_head and ad.is_tail
        un = await sh.unload_model()
        assert un.success and ad.total_layers == 0
        await sh.shutdown()
        assert not rt.started and not ad.running

    asyncio.run(main())


# ---------------------------------------------------------------------------- tensor-parallel lm_head schedule
def test_tp_head_schedule_keeps_a_requests_steps_S_plus_1_entries_apart_and_flushes(grpc_ok):
    """With the lm_head tensor-parallel over an S-shard ring the token of entry j exists after entry j + S, so the
    head shard may schedule the same request again at entry j + S + 1 at the earliest: bubbles fill the gap when
    fewer than S + 1 requests are leased, none are needed with S + 1 or more, and S + 1 trailing bubbles let the
    last tokens' head parts run."""
    S = 4
    for n_req, steps in ((1, 3), (2, 4), (5, 3), (7, 2)):
        ad, rt = make_adapter(assigned_next={0})
        rt._assigned_set = {0}
        ad.total_layers, ad.ring_size, ad.head_tp, ad.rounds_per_frame, ad.head_tp_lag = 8, S, True, 3, 0
        ad._streams.configure_lanes(8)
        for i in range(n_req):
            ad._streams.claim_lane(f"r{i}").params["seq0"] = 1
            ad._leases[f"r{i}"] = steps
        stream = []
        while True:
            e = ad._next_schedule()
            if not e:
                break
            stream += e
        real = [(i, lane, seq) for i, (lane, seq) in enumerate(stream) if lane != fr.BUBBLE]
        assert len(real) == n_req * steps
        last = {}
        for i, lane, seq in real:
            if lane in last:
                assert i - last[lane][0] >= S + 1, f"lane {lane}: entries {last[lane][0]} and {i} too close ({n_req} requests)"
                assert seq == last[lane][1] + 1
            last[lane] = (i, seq)
        bubbles = sum(1 for lane, _ in stream if lane == fr.BUBBLE)
        if n_req >= S + 1:
            assert bubbles == S + 1, "a full ring needs no padding, only the final flush"
        assert all(lane == fr.BUBBLE for lane, _ in stream[-(S + 1):])          # the flush
        # every real entry's token is merged by the head at entry j + S, which exists in the stream
        assert all(i + S < len(stream) for i, _, _ in real)
        assert fr.unpack_sched(fr.pack_sched(stream)) == stream                 # bubbles survive the wire format


def test_lease_is_refused_with_an_error_token_when_the_model_cannot_run_in_the_step_kernel(grpc_ok):
    """Sparse-MoE models run on the per-op path (runtime.use_megakernel False): the device-closed loop does not exist
    for them, so a lease is answered at once with a negative token instead of leaving the API waiting."""
    ad, rt = make_adapter(assigned_next={0})
    rt.use_megakernel = False
    ad._streams.configure_lanes(2)
    ctx = ad._streams.claim_lane("m")
    ctx.params.update(seq0=1, callback_url="grpc://127.0.0.1:9")

    async def main():
        await ad.start()
        ad.lease("m", 4)
        assert "m" not in ad._leases
        msg = rt.activation_send_queue.get_nowait()
        assert msg.is_final and msg.token_id <= -1000 and msg.nonce == "m" and msg.callback_url == "grpc://127.0.0.1:9"
        await ad.shutdown()

    asyncio.run(main())


def test_single_request_is_scheduled_without_the_collection_grace(grpc_ok):
    """The head waits lease_grace_s after the first lease only while another lane-holding request has none yet."""
    ad, rt = make_adapter(assigned_next={0})
    ad.lease_grace_s = 1.0
    ad.rounds_per_frame, ad.sched_depth = 1, 4
    ad._streams.configure_lanes(4)
    ad._streams.claim_lane("solo").params["seq0"] = 1

    async def main():
        await ad.start()
        t0 = asyncio.get_running_loop().time()
        ad.lease("solo", 1)
        assert await wait_until(lambda: rt.activation_recv_queue.qsize() >= 1, timeout=0.8), "scheduled only after the grace"
        assert asyncio.get_running_loop().time() - t0 < 0.8
        rt.activation_recv_queue.get_nowait().sched_done.record(None)
        # a second request holds a lane but has no lease yet: now the grace applies
        ad._streams.claim_lane("other").params["seq0"] = 1
        await asyncio.sleep(0.05)
        t1 = asyncio.get_running_loop().time()
        ad.lease("solo", 1)
        assert await wait_until(lambda: rt.activation_recv_queue.qsize() >= 1, timeout=5.0)
        assert asyncio.get_running_loop().time() - t1 >= 0.9
        await ad.shutdown()

    asyncio.run(main())
