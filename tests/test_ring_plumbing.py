"""BASELINE config 1 -- two-shard pipelined ring with stub layers, on CPU (no GPU): the wire
contract (hand-built protos == reference field numbers), ActivationCodec branches, routing by
layer_id, the final-token contract, and a world_size-2 gloo run of the same ring.

Stub algebra is the reference's FakeComputeModel (tests/fakes/policies.py:41-70):
embed = cast, apply_single_layer(l, x) = x + l, lm_project = arange(4)."""
import os
import queue

import numpy as np
import pytest
import torch

from dnet_b200.core.memory.memory_pool import LayerAwareMemoryPool
from dnet_b200.core.types.messages import ActivationMessage
from dnet_b200.protos import dnet_ring_pb2 as pb
from dnet_b200.protos import shard_api_comm_pb2 as sp
from dnet_b200.shard.codec import ActivationCodec
from dnet_b200.shard.policies.base import ComputePolicy
from dnet_b200.shard.runtime import ShardRuntime
from dnet_b200.utils.serialization import tensor_to_bytes


class StubPolicy(ComputePolicy):
    """fit-policy control flow with the fake model; CPU tensors in a cpu-placed pool"""

    def configure_policy_for_model(self, req):
        pass

    def clear(self):
        pass

    def process(self, msg):
        rt = self.runtime
        buf = rt.input_pool.get_buffer(msg.pool_id)
        n = int(np.prod(msg.shape))
        x = buf[:n].reshape(msg.shape)
        x = x.to(torch.float32)[None] if msg.dtype == "tokens" else x.to(torch.float32)
        cur = msg.layer_id + 1
        while cur in rt._assigned_set:
            x = x + cur
            cur += 1
        last = cur - 1
        common = dict(nonce=msg.nonce, layer_id=last, pool_id=-1, shape=tuple(x.shape), batch_size=msg.batch_size,
                      timestamp=1, node_origin=f"shard_{rt.shard_id}", dtype=rt._wire_dtype_str, callback_url=msg.callback_url)
        if cur >= rt.total_layers:
            logits = torch.arange(4, dtype=torch.float32)
            out = ActivationMessage(**common, is_final=True, token_id=int(torch.argmax(logits)), logprob=0.0, top_logprobs={})
        else:
            out = ActivationMessage(**common, tensor=x.to(rt._wire_mx_dtype))
        rt.emit_result(out)
        rt.input_pool.release(msg.pool_id)


def make_stub_shard(shard_id, layers, total):
    rt = ShardRuntime(shard_id=shard_id, queue_size=8)
    rt.input_pool = LayerAwareMemoryPool(total_memory_mb=4, placement="cpu")
    rt.output_pool = LayerAwareMemoryPool(total_memory_mb=4, placement="cpu")
    rt.assigned_layers = list(layers)
    rt._assigned_sorted = sorted(layers)
    rt._assigned_set = set(layers)
    rt.total_layers = total
    rt.policy = StubPolicy(rt, 1)
    return rt


def tokens_frame(nonce, ids, seq=0):
    req = pb.ActivationRequest(nonce=nonce, activation=pb.Activation(data=np.asarray(ids, np.int32).tobytes(), batch_size=1,
                                                                    shape=[1], dtype="tokens", layer_id=-1),
                               timestamp=0, node_origin="api", callback_url="grpc://api:1", temperature=0.0)
    return pb.ActivationFrame(request=req, seq=seq).SerializeToString()


def shard_step(rt, frame_bytes):
    """servicer.StreamActivations -> admit_frame -> ingress (route by layer_id+1) -> compute -> egress"""
    frame = pb.ActivationFrame.FromString(frame_bytes)
    target = frame.request.activation.layer_id + 1
    assert target in rt._assigned_set, "a real shard would forward this frame untouched"
    msg = ActivationCodec(rt).deserialize(frame.request)
    assert msg is not None
    rt.policy.process(msg)
    out = rt.activation_send_queue.get_nowait()
    if out.is_final:
        return "token", sp.TokenRequest(nonce=out.nonce, token_id=out.token_id, timestamp=out.timestamp,
                                        logprob=out.logprob, top_logprobs=out.top_logprobs or {}).SerializeToString()
    data = ActivationCodec(rt).serialize(out)
    out.dtype = rt._wire_dtype_str
    return "frame", pb.ActivationFrame(request=out.to_proto(data), seq=frame.seq).SerializeToString()


def test_wire_contract_field_numbers():
    r = pb.ActivationRequest(nonce="n", activation=pb.Activation(data=b"\x01\x02", batch_size=1, shape=[1, 2], dtype="tokens",
                                                                layer_id=-1), timestamp=5, temperature=0.0)
    # tag bytes: nonce=0x0a, activation=0x12, timestamp=0x18, temperature (field 8, fixed32)=0x45
    assert r.SerializeToString().hex() == ("0a016e121d0a02010210011a0201022206746f6b656e7328ffffffffffffffffff01"
                                           "18054500000000")
    back = pb.ActivationRequest.FromString(r.SerializeToString())
    assert back.HasField("temperature") and not back.HasField("top_p") and back.activation.layer_id == -1
    t = sp.TokenRequest(nonce="n", token_id=7, logprob=-0.5, top_logprobs={3: -1.0})
    assert t.SerializeToString().hex() == "0a016e100725000000bf2a07080315000080bf"
    assert pb.METHODS["StreamActivations"] == "/dnetring.DnetRingService/StreamActivations"
    assert sp.METHODS["SendToken"] == "/shardapi.ShardApiService/SendToken"
    m = ActivationMessage.from_proto(back, pool_id=3)
    assert m.temperature == 0.0 and m.top_p == 1.0 and m.top_k == -1 and m.pool_id == 3 and m.dtype == "tokens"


def test_codec_branches():
    rt = make_stub_shard("s", [0, 1], 4)
    codec = ActivationCodec(rt)
    f = pb.ActivationFrame.FromString(tokens_frame("n", [5, 6, 7]))
    m = codec.deserialize(f.request)
    assert m.dtype == "tokens" and m.shape == (3,) and rt.input_pool.get_buffer(m.pool_id)[:3].tolist() == [5, 6, 7]
    x = torch.tensor([[1.0, 0.5, 2.0, -1.0]]).to(torch.bfloat16)
    raw = pb.ActivationRequest(nonce="n", activation=pb.Activation(data=tensor_to_bytes(x), batch_size=1, shape=[1, 1, 4],
                                                                  dtype="bfloat16", layer_id=1))
    m2 = codec.deserialize(raw)
    assert m2.shape == (1, 1, 4) and rt.input_pool.get_buffer(m2.pool_id)[:4].tolist() == [1.0, 0.5, 2.0, -1.0]
    bad = pb.ActivationRequest(nonce="n", activation=pb.Activation(data=b"\x00" * 6, batch_size=1, shape=[1, 1, 4],
                                                                  dtype="bfloat16", layer_id=1))
    assert codec.deserialize(bad) is None                      # payload size mismatch
    comp = pb.ActivationRequest(nonce="n", activation=pb.Activation(data=b"", shape=[1], dtype="float16|sparse_v1"))
    assert codec.deserialize(comp) is None                     # compressed branch: dead path, rejected
    out = ActivationMessage(nonce="n", pool_id=-1, batch_size=1, shape=(1, 1, 4), dtype="x", layer_id=1, timestamp=0,
                            node_origin="s", callback_url="", tensor=x.view(1, 1, 4).float())
    assert len(codec.serialize(out)) == 4 * 2 and out.tensor is None   # cast to the (fp16-width) wire dtype


def test_two_shard_ring_stub_layers_in_process():
    a, b = make_stub_shard("a", [0, 1], 4), make_stub_shard("b", [2, 3], 4)
    kind, payload = shard_step(a, tokens_frame("n0", [10, 20, 30]))
    assert kind == "frame"
    fr = pb.ActivationFrame.FromString(payload)
    assert fr.request.activation.layer_id == 1 and list(fr.request.activation.shape) == [1, 3]
    assert fr.request.activation.dtype == a._wire_dtype_str and len(fr.request.activation.data) == 3 * 2
    kind, payload = shard_step(b, payload)
    assert kind == "token"
    tok = sp.TokenRequest.FromString(payload)
    assert tok.token_id == 3 and tok.nonce == "n0"            # argmax(arange(4))
    assert a.activation_send_queue.empty() and b.activation_send_queue.empty()
    assert a.input_pool.get_stats()["pool"]["allocated_buffers"] == 0


def _gloo_worker(rank, world, port, results):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rt = make_stub_shard(f"r{rank}", [0, 1] if rank == 0 else [2, 3], 4)

    def send_bytes(b, dst):
        t = torch.frombuffer(bytearray(b), dtype=torch.uint8)
        dist.send(torch.tensor([t.numel()]), dst)
        dist.send(t, dst)

    def recv_bytes(src):
        n = torch.zeros(1, dtype=torch.int64)
        dist.recv(n, src)
        t = torch.zeros(int(n.item()), dtype=torch.uint8)
        dist.recv(t, src)
        return t.numpy().tobytes()

    toks = []
    ids = [4, 5, 6]
    for step in range(3):
        if rank == 0:
            _, frame = shard_step(rt, tokens_frame("g", ids, seq=step))
            send_bytes(frame, 1)
            tok = sp.TokenRequest.FromString(recv_bytes(1))   # the API callback of the last shard
            toks.append(tok.token_id)
            ids = [tok.token_id]
        else:
            kind, payload = shard_step(rt, recv_bytes(0))
            assert kind == "token"
            send_bytes(payload, 0)
    if rank == 0:
        results.put(toks)
    dist.barrier()
    dist.destroy_process_group()


def test_two_shard_ring_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    results = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, results)) for r in range(2)]
    for p in procs:
        p.start()
    toks = results.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert toks == [3, 3, 3]


def test_even_split_and_hop_endpoint_shapes():
    from dnet_b200.shard.ring import even_split
    assert even_split(32, 8) == [list(range(i * 4, i * 4 + 4)) for i in range(8)]
    assert [len(x) for x in even_split(32, 5)] == [7, 7, 6, 6, 6]
    assert sum(even_split(80, 2), []) == list(range(80))
    from dnet_b200.shard.ring import balanced_split
    lb, head = 436_224_000, 1_050_673_152 + 8_192                 # Llama-3-8B layer / lm_head + final norm
    for world in (1, 2, 4, 8):
        sp = balanced_split(32, world, lb, first_extra=8_192, last_extra=head)
        assert sum(sp, []) == list(range(32)) and all(len(x) >= 1 for x in sp)
        loads = [len(x) * lb + (8_192 if i == 0 else 0) + (head if i == world - 1 else 0) for i, x in enumerate(sp)]
        eq = even_split(32, world)
        eq_loads = [len(x) * lb + (8_192 if i == 0 else 0) + (head if i == world - 1 else 0) for i, x in enumerate(eq)]
        assert max(loads) <= max(eq_loads)
    s8 = balanced_split(32, 8, lb, 8_192, head)
    assert sorted(len(x) for x in s8[:-1]) == [4, 4, 4, 4, 4, 5, 5] and len(s8[-1]) == 2      # busiest shard: 5 layers, not 4 + head
    assert [len(x) for x in balanced_split(32, 2, lb, 8_192, head)] == [17, 15]
    assert [len(x) for x in balanced_split(6, 3, 10)] == [2, 2, 2]
    with pytest.raises(ValueError):
        balanced_split(2, 3, 10)
