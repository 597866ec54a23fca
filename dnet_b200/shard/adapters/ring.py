"""RingAdapter: ring transport + topology glue around the topology-agnostic ShardRuntime
(reference src/dnet/shard/adapters/ring.py:40-487), rebuilt for a ring of B200 shards.

Same seam as the reference -- ingress queue filled by the gRPC servicer, routing by
``layer_id + 1``, egress split into ring / token queues, per-nonce streams owned by a
``StreamManager``, ``configure_topology`` / ``reset_topology`` -- so ``Shard`` and the servicer are
transport-agnostic.  What changed underneath:

* **tensor bytes never enter a frame when a hop link exists.**  ``configure_topology`` gives the shard
  receive lanes in HBM (``HopLink``: per in-flight nonce a decode slot, a bulk slot and their
  sequence flags), publishes their CUDA-IPC endpoint and maps the successor's.  A computed activation
  is stored into the successor's slot over NVLink on a second CUDA stream the moment the policy emits
  it (``_device_egress``, on the compute thread), and the frame that follows on the gRPC stream is
  metadata only (``b200.hop/<dtype>``: lane + sequence number).  Without a link (a reference peer, no
  CUDA IPC) the reference's bytes path is used unchanged.
* **decode runs with the token loop closed on the device.**  The API leases decode steps to the head
  shard (``b200.lease``); the head's scheduler merges all active leases into an ordered schedule of
  (lane, seq) pairs (``b200.sched``) that travels once around the ring; every shard launches one fused
  wait+step+hop kernel per pair, in that order, on its single compute stream
  (``dn_shard_step_hop``) -- N nonces in flight keep N shards busy, no host, protobuf or TCP per
  token.  The finalising shard's kernel hands the token to the head's lane slot and, through a pinned
  ring (``TokenTap``), to the host, which forwards it to the API exactly like the reference
  (``SendToken``).  The host-closed loop of the reference (API sends every token) still works and
  uses the same hop for its activations.
"""
from __future__ import annotations

import asyncio
import queue
import time
from collections import deque
from typing import Any, Awaitable, Callable, Deque, Dict, List, Optional, Tuple
from urllib.parse import urlparse

import numpy as np

from dnet_b200.config import TransportSettings, get_settings
from dnet_b200.core.stream_manager import StreamManager
from dnet_b200.core.types.messages import ActivationMessage
from dnet_b200.protos import dnet_ring_pb2 as pb2
from dnet_b200.protos import shard_api_comm_pb2
from dnet_b200.utils.logger import logger
from dnet_b200.utils.time import utc_epoch_now
from .. import frames as fr
from ..codec import ActivationCodec
from .base import TopologyAdapter

try:  # grpcio is present in the image; the adapter also runs with injected stubs (tests)
    from grpc import aio as aio_grpc
except Exception:  # pragma: no cover
    aio_grpc = None

HopExchange = Callable[[dict], Awaitable[Optional[dict]]]


def _make_ring_stub(channel):
    from dnet_b200.protos.dnet_ring_pb2_grpc import DnetRingServiceStub

    return DnetRingServiceStub(channel)


def _make_api_stub(channel):
    from dnet_b200.protos.shard_api_comm_pb2_grpc import ShardApiServiceStub

    return ShardApiServiceStub(channel)


class RingAdapter(TopologyAdapter):
    def __init__(self, runtime, discovery=None, transport_settings: Optional[TransportSettings] = None) -> None:
        super().__init__(runtime, discovery)
        self.transport_settings: TransportSettings = transport_settings or get_settings().transport
        self.codec = ActivationCodec(runtime)
        self.running = False
        self._active_nonce: Optional[str] = None
        self._streaming_enabled = bool(self.transport_settings.streaming)
        self._streams = StreamManager(idle_timeout_s=self.transport_settings.stream_idle_s,
                                      backoff_s=getattr(self.transport_settings, "stream_backoff_s", 0.5))
        # topology
        self.next_node: Optional[Any] = None
        self.next_node_channel: Optional[Any] = None
        self.next_node_stub: Optional[Any] = None
        self.total_layers: int = 0
        self.api_callback_address: Optional[str] = None
        # queues (reference names)
        self.queue_size = runtime.max_queue_size
        self._ingress_q: asyncio.Queue = asyncio.Queue(maxsize=self.queue_size)
        self.ring_tx_q: asyncio.Queue = asyncio.Queue(maxsize=self.queue_size)
        self.token_tx_q: asyncio.Queue = asyncio.Queue(maxsize=self.queue_size)
        # API callback
        self.api_channel: Optional[Any] = None
        self.api_stub: Optional[Any] = None
        self.api_address: Optional[str] = None
        self.token_sink: Optional[Callable[[ActivationMessage], None]] = None   # in-process API (callback_url local://)
        self._tasks: List[asyncio.Task] = []
        self._loop: Optional[asyncio.AbstractEventLoop] = None
        # device hop + on-device decode schedule
        self.hop = None                                   # HopLink once configure_topology built it
        self.hop_exchange: Optional[HopExchange] = None   # how the successor's endpoint is obtained (default: gRPC)
        self.n_lanes = int(getattr(self.transport_settings, "hop_lanes", 16))
        self.bulk_tokens = int(getattr(self.transport_settings, "hop_bulk_tokens", 512))
        self.rounds_per_frame = int(getattr(self.transport_settings, "sched_rounds_per_frame", 4))
        self.sched_depth = int(getattr(self.transport_settings, "sched_frames_in_flight", 3))
        self.lease_grace_s = float(getattr(self.transport_settings, "lease_grace_s", 5e-4))
        self._leases: Dict[str, int] = {}                 # head: nonce -> decode steps still to schedule
        self._lease_evt: Optional[asyncio.Event] = None
        self._bulk_seq: Dict[int, int] = {}               # lane -> last bulk-flag value this shard sent
        self._dseq: Dict[int, int] = {}                   # head: lane -> next decode-flag value to schedule
        self._sched_seq = 0
        self.stats = {"frames_hop": 0, "frames_bytes": 0, "frames_sched": 0, "tokens": 0}
        # ring census (position / size) and the tensor-parallel lm_head it enables
        self.head_tp_mode = str(getattr(self.transport_settings, "head_tp", "auto")).strip().lower()
        self.ring_size = 1
        self.ring_pos = 0
        self.head_tp = False                  # lm_head split over the ring's shards during on-device decode
        self.head_tp_lag = max(0, int(getattr(self.transport_settings, "head_tp_lag", 1)))
        self._tp_flush = 0                    # head: bubble entries still owed so the last tokens' head parts run
        self.advertise_addr: Optional[str] = None   # this shard's gRPC address (host:port), set by whoever starts its server
        self._ring_addrs: List[Optional[str]] = []  # gRPC address of every ring position (from the census), None = unknown
        self._fan_channels: Dict[int, Any] = {}
        self._fan_stubs: Dict[int, Any] = {}
        self._sched_index = 0                 # head: ring-wide index of the next schedule entry
        self._lane_last_idx: Dict[int, int] = {}

    # ------------------------------------------------------------------ reference surface
    @property
    def ingress_q(self) -> asyncio.Queue:
        return self._ingress_q

    @property
    def activation_computed_queue(self) -> asyncio.Queue:
        return self.ring_tx_q

    @property
    def activation_token_queue(self) -> asyncio.Queue:
        return self.token_tx_q

    @property
    def is_head(self) -> bool:
        return 0 in self.runtime._assigned_set

    @property
    def is_tail(self) -> bool:
        return self.total_layers > 0 and (self.total_layers - 1) in self.runtime._assigned_set

    async def ingress(self):
        pass

    async def egress(self):
        pass

    async def start(self):
        self.running = True
        self._loop = asyncio.get_running_loop()
        self._lease_evt = asyncio.Event()
        self._tasks = [asyncio.create_task(w()) for w in (self._ingress_worker, self._egress_worker,
                                                          self._ring_tx_worker, self._token_tx_worker,
                                                          self._sched_worker)]
        if self._streaming_enabled:
            self._tasks.append(asyncio.create_task(self._stream_sweeper()))

    async def shutdown(self) -> None:
        self.running = False
        for t in self._tasks:
            t.cancel()
        if self._tasks:
            await asyncio.gather(*self._tasks, return_exceptions=True)
        self._tasks.clear()
        for nonce in list(self._streams._streams.keys()):
            await self._streams.end_stream(nonce)
        for ch in (self.next_node_channel, self.api_channel):
            if ch is not None:
                try:
                    await ch.close()
                except Exception:
                    pass
        self.next_node_channel = self.next_node_stub = None
        self.api_channel = self.api_stub = None
        self._teardown_hop()
        logger.info("Shard %s: ring adapter shutdown complete", self.runtime.shard_id)

    async def configure_topology(self, req) -> None:
        self.next_node = req.next_node
        self.total_layers = req.total_layers
        self.api_callback_address = req.api_callback_address
        if self.next_node:
            await self._connect_next_node()
        else:
            logger.warning("Node %s: No next node configured", self.runtime.shard_id)
        await self._setup_hop()

    async def reset_topology(self) -> None:
        self.next_node = None
        self.total_layers = 0
        self.api_callback_address = None
        for attr in ("next_node_channel", "api_channel"):
            ch = getattr(self, attr)
            if ch is not None:
                try:
                    await ch.close()
                except Exception:
                    pass
                setattr(self, attr, None)
        self.next_node_stub = None
        self.api_stub = None
        self.api_address = None
        self._leases.clear()
        for ch in self._fan_channels.values():
            try:
                await ch.close()
            except Exception:
                pass
        self._fan_channels.clear()
        self._fan_stubs.clear()
        self._teardown_hop()

    async def admit_frame(self, request) -> None:
        """Servicer entry: queue a frame for the ingress worker; spin (yielding) while the queue is
        full; drop silently once the adapter is stopped."""
        while self.running:
            try:
                self.ingress_q.put_nowait(request)
                return
            except asyncio.QueueFull:
                await asyncio.sleep(0)

    # ------------------------------------------------------------------ device hop plumbing
    async def _setup_hop(self) -> None:
        """Receiver-owned lanes + endpoint exchange with the ring successor.  Needs a loaded model (for
        the hidden size) and a CUDA device; otherwise the adapter stays on the bytes path."""
        rt = self.runtime
        model = getattr(rt, "model", None)
        if model is None or getattr(rt, "compute_stream", None) is None:
            return
        try:
            from ..ring import HopLink
            from ..token_tap import TokenTap

            bind = getattr(rt, "_bind_thread_device", None)
            if bind is not None:
                bind()                      # this runs on the event-loop thread: bind it to the shard's device
            self._teardown_hop()
            layers = sorted(rt._assigned_set)
            hop = HopLink(self.n_lanes, int(model.hidden_size), self.bulk_tokens, shard_id=rt.shard_id,
                          first_layer=layers[0] if layers else -1, n_layers=len(layers))
            self._streams.configure_lanes(self.n_lanes)
            self._bulk_seq.clear()
            self._dseq.clear()
            self._sched_index, self._tp_flush = 0, 0
            self._lane_last_idx.clear()
            single = self.next_node is None
            self.ring_size, self.ring_pos, self.head_tp = 1, 0, False
            if single:
                hop.connect(None)
            else:
                exchange = self.hop_exchange or self._grpc_hop_exchange
                rt.hop_pending = hop              # lets the servicer answer the predecessors' b200.hop.open
                # ring census: ask for the endpoint k hops ahead until our own comes back
                ring: list = []
                for k in range(64):
                    ep = await exchange(self._endpoint(hop), k)
                    if ep is None:
                        break
                    if self._ep_id(ep) == str(rt.shard_id):
                        break
                    ring.append(ep)
                if not ring:
                    logger.warning("Shard %s: no hop endpoint from the next node; tensor bytes ride the gRPC stream",
                                   rt.shard_id)
                    hop.close()
                    rt.hop_pending = None
                    return
                ep = ring[0]
                if isinstance(ep, HopLink):      # the successor lives in this process: no IPC
                    hop.connect_local(ep)
                else:
                    hop.connect(ep)
                self._census(hop, ring)
            self.hop = hop
            rt.hop = hop
            rt.on_emit = self._device_egress
            rt.tp_head = None
            if self.head_tp:
                row0, row1 = await asyncio.get_running_loop().run_in_executor(rt.executor, rt.load_head_slice,
                                                                               self.ring_pos, self.ring_size)
                rt.tp_head = {"S": self.ring_size, "r": self.ring_pos, "rows": (row0, row1), "lag": self.head_tp_lag}
                logger.info("Shard %s: lm_head tensor-parallel over %d shards, this shard rows %d..%d", rt.shard_id,
                            self.ring_size, row0, row1 - 1)
            # tokens surface where they are finalised: the tail shard, or the head shard with a tensor-parallel head
            if (self.is_tail and not self.head_tp) or single or (self.is_head and self.head_tp):
                rt.token_tap = TokenTap(self.n_lanes, on_token=self._tap_token)
                rt.token_tap.start()
            logger.info("Shard %s: device hop link up (%d lanes, head=%s tail=%s)", rt.shard_id, self.n_lanes,
                        self.is_head, self.is_tail)
        except Exception as e:
            logger.warning("Shard %s: device hop unavailable (%s); falling back to bytes over gRPC", rt.shard_id, e)
            self.hop = None
            rt.hop = None

    def _teardown_hop(self) -> None:
        rt = self.runtime
        tap = getattr(rt, "token_tap", None)
        if tap is not None:
            tap.stop()
            rt.token_tap = None
        if self.hop is not None:
            try:
                if getattr(rt, "compute_stream", None) is not None:
                    rt.compute_stream.synchronize()
                self.hop.close()
            except Exception:
                pass
        self.hop = None
        if hasattr(rt, "hop"):
            rt.hop = None
            rt.on_emit = None

    def _endpoint(self, hop) -> dict:
        ep = hop.endpoint()
        ep["grpc_addr"] = self.advertise_addr
        return ep

    @staticmethod
    def _ep_id(ep) -> str:
        return str(ep.shard_id if hasattr(ep, "shard_id") else ep.get("shard_id"))

    def _census(self, hop, ring: list) -> None:
        """``ring`` = endpoints of the successors in ring order (ours excluded).  Derive the ring size, our
        position counted from the shard that owns layer 0, and -- when every shard runs one contiguous block and
        the setting allows -- map every peer's head area (tensor-parallel lm_head)."""
        from ..ring import HopLink

        def first_layer(ep):
            return int(ep.first_layer if isinstance(ep, HopLink) else ep.get("first_layer", -1))

        order = ring + [hop]                                   # successors..., self
        heads = [i for i, ep in enumerate(order) if first_layer(ep) == 0]
        S = len(order)
        if len(heads) != 1:
            self.ring_size, self.ring_pos, self.head_tp = S, 0, False
            return
        h = heads[0]
        order = order[h:] + order[:h]                          # ring positions 0..S-1
        self._ring_addrs = [(self.advertise_addr if ep is hop else (None if isinstance(ep, HopLink) else ep.get("grpc_addr")))
                            for ep in order]
        self.ring_size = S
        self.ring_pos = next(i for i, ep in enumerate(order) if ep is hop)
        want = self.head_tp_mode in ("on", "1", "true") or (self.head_tp_mode == "auto" and S >= 4)
        ordered = all(first_layer(order[i]) < first_layer(order[i + 1]) for i in range(S - 1))
        self.head_tp = bool(want and S >= 2 and S <= hop.mesh.MAX_SHARDS and ordered)
        if self.head_tp:
            hop.mesh.connect([None if ep is hop else (ep.mesh if isinstance(ep, HopLink) else ep["head"]) for ep in order],
                             self.ring_pos)

    async def _grpc_hop_exchange(self, own: dict, k: int = 0) -> Optional[dict]:
        """Ask the successor for its endpoint with a unary SendActivation whose dtype is ``b200.hop.open``;
        a dnet_b200 servicer answers with the endpoint as JSON in ``ActivationResponse.message``."""
        import json

        if self.next_node_stub is None:
            return None
        req = pb2.ActivationRequest(nonce="", activation=pb2.Activation(data=b"", batch_size=0, shape=[int(k)], dtype="b200.hop.open",
                                                                      layer_id=-1),
                                    timestamp=utc_epoch_now(), node_origin=f"shard_{self.runtime.shard_id}", callback_url="")
        for _ in range(200):        # the successor may still be loading its model
            try:
                resp = await self.next_node_stub.SendActivation(req, timeout=10.0)
                if resp.success and resp.message.startswith("{"):
                    return json.loads(resp.message)
            except Exception as e:
                logger.debug("hop endpoint request failed: %s", e)
            await asyncio.sleep(0.1)
        return None

    def _device_egress(self, msg: ActivationMessage) -> None:
        """runtime.emit_result hook, on the compute thread: move the tensor (or the request's first token)
        to the successor's lane over NVLink *now*, on the comm stream, so only metadata is left to frame."""
        hop, rt = self.hop, self.runtime
        if hop is None or not hop.connected or msg.lane < 0:
            return
        lib = rt.lib
        if msg.is_final:
            if msg.seq0 > 0 and msg.token_id >= 0:
                ns = rt._kv_by_nonce.get(msg.nonce)
                if ns is not None:   # the sampled token sits in the nonce's device step state
                    rt.check(lib.dn_hop_send(hop.tx_slot(msg.lane), ns.kv.token_ptr, 4, hop.tx_flag(msg.lane), msg.seq0,
                                             rt.compute_stream_ptr))
            return
        t = msg.tensor
        if t is None or not getattr(t, "is_cuda", False):
            return
        T = int(t.numel() // hop.hidden)
        if T > hop.bulk_tokens or str(t.dtype) != "torch.bfloat16":
            return
        lane = msg.lane
        seq = self._bulk_seq.get(lane, 0) + 1
        self._bulk_seq[lane] = seq
        cs = rt.comm_stream
        if msg.ready_event is not None:
            cs.wait_event(msg.ready_event)
        else:
            cs.wait_stream(rt.compute_stream)
        hop.tx_bulk.wait_consumed(lane, seq - 1, int(cs.cuda_stream))     # credit: previous chunk copied out
        hop.tx_bulk.send(lane, t.data_ptr(), T * hop.hidden * 2, seq, int(cs.cuda_stream))
        ns = rt._kv_by_nonce.get(msg.nonce)
        if ns is not None:       # the nonce's activation buffer may only be rewritten after the copy
            import torch

            ev = torch.cuda.Event()
            ev.record(cs)
            ns.hop_sent_event = ev
        msg.hop_meta = fr.HopMeta(lane=lane, seq=seq, seq0=msg.seq0)
        msg.tensor = None

    def _tap_token(self, info: Tuple[str, int, dict], token: int, logprob: float) -> None:
        """TokenTap consumer (watcher thread): one final message per decoded token, as the reference's
        end shard emits (fit_in_memory.py:168-181)."""
        nonce, step_seq, params = info
        rt = self.runtime
        if token <= -1000:
            logger.error("Shard %s: step kernel reported error %d for nonce %s", rt.shard_id, -token - 1000, nonce)
        msg = ActivationMessage(nonce=nonce, pool_id=-1, batch_size=1, shape=(1, 1, int(rt.model.hidden_size)),
                                dtype=rt._wire_dtype_str, layer_id=int(rt._assigned_sorted[-1]), timestamp=utc_epoch_now(),
                                node_origin=f"shard_{rt.shard_id}", callback_url=params.get("callback_url", ""),
                                is_final=True, token_id=int(token), logprob=float(logprob) if params.get("logprobs") else 0.0,
                                top_logprobs={})
        rt.activation_send_queue.put(msg)

    # ------------------------------------------------------------------ leases + schedule (head shard)
    def lease(self, nonce: str, steps: int) -> None:
        """Head shard: allow ``steps`` more on-device decode steps for ``nonce`` (thread-safe)."""
        if not getattr(self.runtime, "use_megakernel", True):
            # the device-closed loop lives in the persistent step kernel; models that run on the per-op path (sparse
            # MoE) are driven by the host-closed loop.  Answer with an error token so the API does not wait.
            logger.error("Shard %s: lease for nonce %s refused: this model does not run in the step kernel "
                         "(use the host-closed token loop)", self.runtime.shard_id, nonce)
            ctx = self._streams.lane_ctx(nonce)
            cb = ctx.params.get("callback_url", "") if ctx is not None else ""
            rt = self.runtime
            rt.activation_send_queue.put(ActivationMessage(
                nonce=nonce, pool_id=-1, batch_size=1, shape=(1,), dtype=rt._wire_dtype_str, layer_id=-1,
                timestamp=utc_epoch_now(), node_origin=f"shard_{rt.shard_id}", callback_url=cb, is_final=True,
                token_id=-1099, logprob=0.0, top_logprobs={}))
            return

        def _add():
            self._leases[nonce] = self._leases.get(nonce, 0) + int(steps)
            if self._lease_evt is not None:
                self._lease_evt.set()
        if self._loop is not None and self._loop.is_running():
            self._loop.call_soon_threadsafe(_add)
        else:
            _add()

    async def end_request(self, nonce: str) -> None:
        """End of a request (an ``end_of_request`` frame): stop scheduling it, free its lane and KV here
        (behind whatever is still queued for it on the compute stream) and tell the next shard."""
        self._leases.pop(nonce, None)
        self._streams.release_lane(nonce)
        rel = getattr(self.runtime, "release_nonce_deferred", None)
        if rel is not None:
            rel(nonce)
        if self.next_node_stub is not None and not self.is_tail:
            ctx = self._streams.get_ctx(nonce)
            if ctx is not None and ctx.open:
                ctx.last_seq += 1
                eor = pb2.ActivationFrame(request=pb2.ActivationRequest(nonce=nonce), seq=ctx.last_seq, end_of_request=True)
                await ctx.queue.put(eor)
        await self._streams.end_stream(nonce)

    def _next_schedule(self) -> List[Tuple[int, int]]:
        """Up to ``rounds_per_frame`` rounds; one round = one decode step of every leased nonce, in lane
        order.  The order is what every shard launches in, so it is decided exactly once, here.
        With a tensor-parallel lm_head the token of entry j is finalised by the head shard at the END of its kernel
        for entry j + S, so the same request's next step may be entry j + S + 1 at the
        earliest: bubble entries are inserted wherever that distance would be violated, and S + 1 bubbles flush the
        head parts of the last tokens once nothing is leased any more."""
        entries: List[Tuple[int, int]] = []
        gap = self.ring_size + 1 + self.head_tp_lag if self.head_tp else 0
        for _ in range(self.rounds_per_frame):
            live = [(self._streams.lane_ctx(n), n) for n, left in self._leases.items() if left > 0]
            live = sorted(((c.lane, n, c) for c, n in live if c is not None and c.lane >= 0), key=lambda x: x[0])
            if not live:
                break
            for lane, nonce, ctx in live:
                if gap:
                    short = gap - (self._sched_index - self._lane_last_idx.get(lane, -(1 << 30)))
                    if short > 0:
                        entries += [(fr.BUBBLE, 0)] * short
                        self._sched_index += short
                    self._lane_last_idx[lane] = self._sched_index
                    self._tp_flush = gap
                seq = self._dseq.get(lane, ctx.params.get("seq0", 1))
                self._dseq[lane] = seq + 1
                entries.append((lane, seq))
                self._sched_index += 1
                ctx.steps_enqueued += 1
                self._streams.note_scheduled(ctx, seq + 1)
                self._leases[nonce] -= 1
        for n in [n for n, left in self._leases.items() if left <= 0]:
            self._leases.pop(n, None)
        if not entries and self._tp_flush:
            entries = [(fr.BUBBLE, 0)] * self._tp_flush
            self._sched_index += self._tp_flush
            self._tp_flush = 0
        return entries

    async def _sched_worker(self):
        rt = self.runtime
        pending: Deque[Any] = deque()
        while self.running:
            try:
                if not self._leases and not self._tp_flush:
                    self._lease_evt.clear()
                    await self._lease_evt.wait()
                    # leases of concurrent requests arrive as separate frames within a few hundred microseconds: collect
                    # them before fixing the order, or the first request would be scheduled alone (bubble-padded rounds)
                    # (nothing to wait for when every request that holds a lane here has its lease already)
                    waiting = [n for n in self._streams.lanes_in_use() if self._leases.get(n, 0) <= 0]
                    if waiting:
                        await asyncio.sleep(self.lease_grace_s)
                if not self.is_head:
                    self._leases.clear()
                    continue
                while pending and pending[0].is_set():
                    pending.popleft()
                if len(pending) >= self.sched_depth:
                    await asyncio.sleep(1e-4)
                    continue
                entries = self._next_schedule()
                if not entries:
                    await asyncio.sleep(0)
                    continue
                self._sched_seq += 1
                msg = self._sched_message(entries)
                pending.append(msg.sched_done)
                # FIFO with prompts: through the same compute queue, and on to the successor
                await self._enqueue_compute(msg)
                if not self.is_tail:
                    await self._publish_schedule(entries)
                self.stats["frames_sched"] += 1
            except asyncio.CancelledError:
                break
            except Exception as e:
                logger.error("schedule worker error: %s", e)
                await asyncio.sleep(0.01)

    async def _publish_schedule(self, entries) -> None:
        """Head shard: get a schedule frame to every other shard.  When the census told us every shard's gRPC address
        the frame is sent to all of them directly (batch_size = 2 marks it "do not relay"), so a shard S hops away
        starts launching one RPC after the head instead of S relays later; otherwise it is relayed round the ring."""
        req = self._sched_request(entries)
        targets = [i for i in range(self.ring_size) if i != self.ring_pos]
        if len(self._ring_addrs) == self.ring_size and all(self._ring_addrs[i] for i in targets) and aio_grpc is not None:
            req.activation.batch_size = 2
            for i in targets:
                stub = self._fan_stubs.get(i)
                if stub is None:
                    ch = aio_grpc.insecure_channel(self._ring_addrs[i])
                    self._fan_channels[i] = ch
                    stub = self._fan_stubs[i] = _make_ring_stub(ch)
                await self._stream_put(f"{req.nonce}@{i}", req, stub)
            return
        await self._connect_next_node()
        await self._forward_activation(req)

    def _sched_message(self, entries) -> ActivationMessage:
        msg = ActivationMessage(nonce="", pool_id=-1, batch_size=1, shape=(len(entries),), dtype=fr.SCHED_DTYPE, layer_id=-1,
                                timestamp=utc_epoch_now(), node_origin=f"shard_{self.runtime.shard_id}", callback_url="",
                                sched=list(entries))
        msg.sched_done = fr.SchedTicket()
        return msg

    def _sched_request(self, entries):
        return pb2.ActivationRequest(nonce=f"sched-{self.runtime.shard_id}",
                                     activation=pb2.Activation(data=fr.pack_sched(entries), batch_size=1, shape=[len(entries)],
                                                               dtype=fr.SCHED_DTYPE, layer_id=-1),
                                     timestamp=utc_epoch_now(), node_origin=f"shard_{self.runtime.shard_id}", callback_url="")

    # ------------------------------------------------------------------ ingress
    async def _ingress_worker(self):
        """Drain the ingress queue: frames for a local layer are staged (heavy work on the executor) and
        queued for the compute thread, everything else is forwarded untouched."""
        loop = asyncio.get_running_loop()
        rt = self.runtime
        while self.running:
            try:
                req = await self.ingress_q.get()
            except asyncio.CancelledError:
                break
            try:
                await self._connect_next_node()
                activation = req.activation
                dtype = activation.dtype
                if dtype == fr.SCHED_DTYPE:
                    entries = fr.unpack_sched(activation.data)
                    await self._enqueue_compute(self._sched_message(entries))
                    if not self.is_tail and activation.batch_size != 2:      # 2 = sent to every shard directly by the head
                        await self._forward_activation(req)
                    continue
                if dtype == fr.LEASE_DTYPE:
                    steps, token = fr.unpack_lease(activation.data)
                    if token is not None:   # seed the lane's token on the device, in order with the compute queue
                        m = ActivationMessage.from_proto(req, pool_id=-1)
                        ctx = self._streams.lane_ctx(req.nonce)
                        m.lane = ctx.lane if ctx else -1
                        m.seq0 = self._dseq.get(m.lane, ctx.params.get("seq0", 1) if ctx else 1)
                        m.token_id = int(token)
                        await self._enqueue_compute(m)
                    self.lease(req.nonce, steps)
                    continue
                target_layer = activation.layer_id + 1
                if req.nonce != self._active_nonce:     # new sequence on this node
                    self._active_nonce = req.nonce
                    # the reference builds the nonce's KV here; CUDA state is created on the compute thread
                    # (bound to the shard's device), so the runtime is only told about the nonce
                    note = getattr(rt, "note_new_nonce", None)
                    (note or rt.get_or_make_kv)(req.nonce)
                if target_layer not in rt._assigned_set:
                    await self._forward_activation(req)
                    continue
                if fr.is_hop(dtype):
                    msg = self._hop_message(req)
                else:
                    try:
                        msg = await loop.run_in_executor(rt.executor, self.codec.deserialize, req)
                    except Exception as e:
                        logger.error("Codec deserialize failed for nonce %s: %s", req.nonce, e)
                        continue
                    if msg is not None and self.hop is not None and self.is_head and msg.dtype == "tokens":
                        self._claim_head_lane(msg)
                if msg is not None:
                    await self._enqueue_compute(msg)
            except Exception as e:
                logger.error("Ingress worker error: %s", e)

    async def _enqueue_compute(self, msg: ActivationMessage) -> None:
        """Hand a message to the compute thread without ever blocking the event loop."""
        q = self.runtime.activation_recv_queue
        while self.running:
            try:
                q.put_nowait(msg)
                return
            except queue.Full:
                await asyncio.sleep(0)

    def _claim_head_lane(self, msg: ActivationMessage) -> None:
        """Head shard, first frame of a request: pick the lane every shard will use for this nonce and the
        decode-flag value its first on-device step will wait for."""
        ctx = self._streams.claim_lane(msg.nonce)
        if ctx is None:
            logger.warning("no free hop lane for nonce %s: serving it on the bytes path", msg.nonce)
            return
        lane = ctx.lane
        if "seq0" not in ctx.params:
            seq0 = max(self._dseq.get(lane, 1), ctx.base_seq + 1)
            ctx.params.update(seq0=seq0, callback_url=msg.callback_url, logprobs=msg.req_logprobs)
            self._dseq[lane] = seq0
            self._streams.note_scheduled(ctx, seq0)
        msg.lane, msg.seq0 = lane, ctx.params["seq0"]

    def _hop_message(self, req) -> Optional[ActivationMessage]:
        """A metadata-only frame: the tensor already sits in OUR bulk slot (or is on its way)."""
        from ..ring import device_view
        import torch

        hop = self.hop
        if hop is None:
            logger.error("hop frame for nonce %s but no hop link on shard %s", req.nonce, self.runtime.shard_id)
            return None
        meta = fr.unpack_hop(req.activation.data)
        msg = ActivationMessage.from_proto(req, pool_id=-1)
        msg.dtype = fr.hop_wire_dtype(req.activation.dtype)
        shape = tuple(req.activation.shape)
        ctx = self._streams.claim_lane(req.nonce, meta.lane)
        if ctx is None:
            logger.error("lane %d of nonce %s is owned by another nonce here", meta.lane, req.nonce)
            return None
        ctx.params.setdefault("callback_url", req.callback_url)
        ctx.params.setdefault("logprobs", bool(req.logprobs))
        ctx.params.setdefault("seq0", meta.seq0)
        msg.lane, msg.seq0 = meta.lane, meta.seq0
        msg.tensor = device_view(hop.rx_bulk.slot(meta.lane), shape, torch.bfloat16)
        msg.hop_wait = (hop.rx_bulk.flag(meta.lane), meta.seq, hop.rx_bulk.consumed_flag(meta.lane))
        return msg

    # ------------------------------------------------------------------ egress
    async def _egress_worker(self):
        loop = asyncio.get_running_loop()
        q = self.runtime.activation_send_queue
        while self.running:
            try:
                msg = await loop.run_in_executor(self.runtime.executor, lambda: q.get(timeout=0.5))
            except asyncio.CancelledError:
                break
            except queue.Empty:
                continue
            await (self.token_tx_q if msg.is_final else self.ring_tx_q).put(msg)

    async def _pump(self, q: asyncio.Queue, deliver) -> None:
        """drain one egress queue; keeps going after stop until the queue is empty"""
        while self.running or not q.empty():
            await deliver(await q.get())

    async def _ring_tx_worker(self):
        await self._pump(self.ring_tx_q, self._send_activation)

    async def _token_tx_worker(self):
        await self._pump(self.token_tx_q, self._send_token)

    async def _stream_sweeper(self):
        while self.running:
            await asyncio.sleep(1.0)
            await self._streams.cleanup_idle_streams()

    async def _stream_put(self, nonce: str, request, stub=None) -> bool:
        """One frame onto the stream keyed ``nonce`` (created on first use) to the next node, or to ``stub``."""
        stub = stub or self.next_node_stub
        if not (self._streaming_enabled and stub):
            return False
        ctx = await self._streams.get_or_create_stream(nonce, lambda it: stub.StreamActivations(it))
        if not ctx or not ctx.open or ctx.disabled:
            logger.error("Stream not available for nonce %s", nonce)
            return False
        ctx.last_seq += 1
        await ctx.queue.put(pb2.ActivationFrame(request=request, seq=ctx.last_seq, end_of_request=False))
        ctx.touch()
        return True

    async def _forward_activation(self, request) -> None:
        if not (self._streaming_enabled and self.next_node_stub):
            logger.error("Streaming disabled or next node not connected; cannot forward")
            return
        await self._stream_put(request.nonce, request)

    async def _send_activation(self, msg: ActivationMessage) -> None:
        if not (self._streaming_enabled and self.next_node_stub):
            logger.error("Streaming disabled or next node not connected; cannot send")
            return
        meta = getattr(msg, "hop_meta", None)
        if meta is not None:           # the tensor already travelled over NVLink (_device_egress)
            data = fr.pack_hop(meta)
            msg.dtype = fr.hop_dtype(self.runtime._wire_dtype_str)
            self.stats["frames_hop"] += 1
        else:
            try:
                data = self.codec.serialize(msg, self.transport_settings)
            except Exception as e:
                logger.error("Serialization failed for nonce %s: %s", msg.nonce, e)
                return
            msg.dtype = self.runtime._wire_dtype_str
            self.stats["frames_bytes"] += 1
        request = msg.to_proto(data)
        request.timestamp = int(time.time() * 1000)
        await self._stream_put(msg.nonce, request)
        msg.tensor = None

    async def _send_token(self, msg: ActivationMessage) -> None:
        """Final-hop delivery of a sampled token to the API (transport only)."""
        self.stats["tokens"] += 1
        cb = msg.callback_url or ""
        if self.token_sink is not None:          # the API lives in this process: no RPC to ourselves
            self.token_sink(msg)
            return
        if cb.startswith("local://"):
            return
        if cb:
            parsed = urlparse(cb)
            if parsed.scheme != "grpc" or not parsed.netloc:
                logger.error("Shard %s: invalid gRPC callback URL for token: %s", self.runtime.shard_id, cb)
                return
            addr = parsed.netloc
        elif self.api_callback_address:
            addr = self.api_callback_address
        else:
            logger.error("Shard %s: no callback URL for final token; nonce=%s", self.runtime.shard_id, msg.nonce)
            return
        try:
            if self.api_channel is None or addr != self.api_address:
                if self.api_channel is not None:
                    try:
                        await self.api_channel.close()
                    except Exception:
                        pass
                self.api_address = addr
                self.api_channel = aio_grpc.insecure_channel(addr)
                self.api_stub = _make_api_stub(self.api_channel)
        except Exception as e:
            logger.error("Shard %s: failed to create API channel for %s: %s", self.runtime.shard_id, addr, e)
            self.api_channel = None
            self.api_stub = None
            return
        try:
            req = shard_api_comm_pb2.TokenRequest(nonce=msg.nonce, token_id=int(msg.token_id), timestamp=utc_epoch_now(),
                                                  logprob=float(msg.logprob), top_logprobs=msg.top_logprobs or {})
            if self.api_stub is None:
                logger.error("Shard %s: API stub not available for nonce=%s", self.runtime.shard_id, msg.nonce)
                return
            resp = await self.api_stub.SendToken(req, timeout=3.0)
            if resp is None or not resp.success:
                logger.error("Shard %s: API SendToken failed for nonce=%s token=%s: %s", self.runtime.shard_id, msg.nonce,
                             msg.token_id, getattr(resp, "message", ""))
        except Exception as e:
            logger.exception("Shard %s: error sending token via gRPC for nonce=%s: %s", self.runtime.shard_id, msg.nonce, e)

    # ------------------------------------------------------------------ next-node channel
    def _next_address(self) -> str:
        node = self.next_node
        ip = getattr(node, "local_ip", None) or getattr(node, "ip", "127.0.0.1")
        return f"{ip}:{getattr(node, 'shard_port', getattr(node, 'port', 0))}"

    async def _connect_next_node(self) -> bool:
        """True if connected or there is no next node (final shard)."""
        if not self.next_node:
            return True
        if self.next_node_channel:
            return True
        address = ""
        try:
            address = self._next_address()
            self.next_node_channel = aio_grpc.insecure_channel(address)
            self.next_node_stub = _make_ring_stub(self.next_node_channel)
            return True
        except Exception as e:
            logger.warning("Shard %s failed to connect to next node %s: %s", self.runtime.shard_id, address, e)
            self.next_node_channel = None
            self.next_node_stub = None
            return False

    async def _reconnect_next_node(self) -> bool:
        if self.next_node_channel:
            try:
                await self.next_node_channel.close()
            except Exception:
                pass
        self.next_node_channel = None
        self.next_node_stub = None
        return await self._connect_next_node()
