"""Per-nonce stream lifecycle (reference src/dnet/core/stream_manager.py:17-130), rebuilt for a
ring whose tensor bytes move over NVLink.

A nonce owns TWO things on a shard, and one ``StreamContext`` tracks both:

* the **control stream** the reference has: a request-scoped bidirectional RPC to the next node
  (``get_or_create_stream(nonce, call_factory)``), its ACK reader with the "backpressure" back-off,
  ``end_stream`` and the idle sweep -- same public surface and semantics, so the reference's own
  assertions (tests/test_stream_manager.py:13-59) hold;
* the **device lane**: the index of the nonce's hop slot + sequence flag (the same index on every
  shard of the ring, chosen by the head shard), the flag value the lane started from, how many
  decode steps were enqueued on the CUDA compute stream for it and a CUDA event recorded behind the
  last of them.  A lane is only handed to another nonce once that event has completed, i.e. when no
  kernel that spins on the lane's flag is still queued (``release_lane`` / ``reap_lanes``).

Sequence numbers on a lane only grow: step s of a request waits for ``base_seq + s + 1`` and the
finalising shard publishes the next token with ``base_seq + s + 2`` (see shard/adapters/ring.py).
"""
from __future__ import annotations

import asyncio
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional

from dnet_b200.utils.logger import logger

CallFactory = Callable[[Any], Any]
_CLOSE = None   # sentinel that ends a request iterator


@dataclass
class StreamContext:
    nonce: str
    queue: asyncio.Queue
    call: Optional[Any] = None
    ack_task: Optional[asyncio.Task] = None
    open: bool = False
    disabled: bool = False
    disabled_until: float = 0.0
    last_seq: int = 0
    last_activity_t: float = 0.0
    # ---- device lane (dnet_b200) ----
    lane: int = -1                 # hop slot / flag index on every shard, -1 = none claimed
    base_seq: int = 0              # flag value when the lane was claimed
    steps_enqueued: int = 0        # decode steps launched on the compute stream for this nonce
    tail_event: Optional[Any] = None   # object with .query(): True once the last enqueued step has run
    params: Dict[str, Any] = field(default_factory=dict)   # decoding parameters of the request

    def touch(self) -> None:
        try:
            self.last_activity_t = asyncio.get_running_loop().time()
        except RuntimeError:
            pass

    def quiescent(self) -> bool:
        ev = self.tail_event
        if ev is None:
            return True
        try:
            return bool(ev.query())
        except Exception:
            return True


class StreamManager:
    """Owns the per-nonce contexts: control stream + ACK reader + idle cleanup + device lanes."""

    def __init__(self, *, idle_timeout_s: float = 30.0, backoff_s: float = 0.5, n_lanes: int = 0) -> None:
        self._streams: Dict[str, StreamContext] = {}
        self._idle_timeout_s = float(idle_timeout_s)
        self._backoff_s = float(backoff_s)
        self._lanes: Dict[str, StreamContext] = {}      # nonce -> context holding a lane (may have no control stream)
        self._free_lanes: List[int] = list(range(int(n_lanes)))[::-1]
        self._draining: List[StreamContext] = []        # released lanes whose last kernel is still queued
        self._lane_seq: Dict[int, int] = {}             # lane -> highest flag value ever scheduled on it

    # ------------------------------------------------------------------ control streams
    def get_ctx(self, nonce: str) -> Optional[StreamContext]:
        return self._streams.get(nonce)

    async def get_or_create_stream(self, nonce: str, call_factory: CallFactory) -> Optional[StreamContext]:
        loop = asyncio.get_running_loop()
        ctx = self._streams.get(nonce)
        if ctx is not None and ctx.open:
            if ctx.disabled and loop.time() >= ctx.disabled_until:
                ctx.disabled = False          # back-off elapsed: the next frame may go out again
            return ctx
        lane_ctx = self._lanes.get(nonce)
        if lane_ctx is not None and not lane_ctx.open:
            ctx = lane_ctx                    # the nonce already holds a lane: attach the stream to the same record
            ctx.queue = asyncio.Queue(maxsize=64)
        else:
            ctx = StreamContext(nonce=nonce, queue=asyncio.Queue(maxsize=64))
        self._streams[nonce] = ctx

        async def frames():
            while True:
                item = await ctx.queue.get()
                if item is _CLOSE:
                    return
                yield item

        ctx.call = call_factory(frames())
        ctx.open = True
        ctx.last_activity_t = loop.time()
        ctx.ack_task = asyncio.create_task(self._read_acks(ctx))
        return ctx

    async def _read_acks(self, ctx: StreamContext) -> None:
        """Negative ACKs are logged; an ACK whose message mentions backpressure pauses the stream for
        ``backoff_s``; a broken response stream closes and disables the context."""
        try:
            async for ack in ctx.call:
                if not getattr(ack, "accepted", True):
                    logger.debug("[STREAM][ACK] nonce=%s seq=%s rejected: %s", getattr(ack, "nonce", ""),
                                 getattr(ack, "seq", -1), getattr(ack, "message", ""))
                if "backpressure" in str(getattr(ack, "message", "")).lower():
                    ctx.disabled = True
                    ctx.disabled_until = asyncio.get_running_loop().time() + self._backoff_s
        except asyncio.CancelledError:
            raise
        except Exception as e:
            logger.error("[STREAM] ack reader error: %s", e)
            ctx.open = False
            ctx.disabled = True

    async def end_stream(self, nonce: str) -> None:
        ctx = self._streams.pop(nonce, None)
        if ctx is None:
            return
        if ctx.ack_task is not None:
            ctx.ack_task.cancel()
        try:
            await ctx.queue.put(_CLOSE)
            closer = getattr(ctx.call, "aclose", None)
            if ctx.open and closer is not None:
                await closer()
        except Exception:
            pass
        ctx.open = False

    async def cleanup_idle_streams(self) -> int:
        now = asyncio.get_running_loop().time()
        stale = [n for n, c in self._streams.items() if (now - c.last_activity_t) > self._idle_timeout_s]
        for nonce in stale:
            await self.end_stream(nonce)
        self.reap_lanes()
        return len(stale)

    # ------------------------------------------------------------------ device lanes
    def configure_lanes(self, n_lanes: int) -> None:
        """(Re)initialise the lane free list; flag values persist per lane across reconfiguration of the
        same hop buffers only if the caller keeps them (see RingAdapter.configure_topology)."""
        self._free_lanes = list(range(int(n_lanes)))[::-1]
        self._lanes.clear()
        self._draining.clear()
        self._lane_seq = {}

    def lane_ctx(self, nonce: str) -> Optional[StreamContext]:
        return self._lanes.get(nonce)

    def lanes_in_use(self) -> Dict[str, int]:
        return {n: c.lane for n, c in self._lanes.items()}

    def claim_lane(self, nonce: str, lane: Optional[int] = None) -> Optional[StreamContext]:
        """Give ``nonce`` a hop lane.  The head shard passes ``lane=None`` and takes the next free index;
        every other shard is told the index by the head's frame and claims exactly that one."""
        ctx = self._lanes.get(nonce)
        if ctx is not None:
            return ctx
        self.reap_lanes()
        if lane is None:
            if not self._free_lanes:
                return None
            lane = self._free_lanes.pop()
        else:
            lane = int(lane)
            if lane in self._free_lanes:
                self._free_lanes.remove(lane)
            elif any(c.lane == lane for c in self._lanes.values()):
                return None            # the head would never hand out a lane twice: refuse
            else:
                # still draining here (its previous owner's last kernel has not run yet): kernels of the
                # new owner are queued behind it on the same compute stream, so taking it over is safe
                self._draining = [c for c in self._draining if c.lane != lane]
        ctx = self._streams.get(nonce) or StreamContext(nonce=nonce, queue=asyncio.Queue(maxsize=64))
        ctx.lane = lane
        ctx.base_seq = self._lane_seq.get(lane, 0)
        ctx.steps_enqueued = 0
        ctx.tail_event = None
        self._lanes[nonce] = ctx
        return ctx

    def note_scheduled(self, ctx: StreamContext, upto_seq: int, tail_event: Any = None) -> None:
        """Record that kernels waiting for flag values up to ``upto_seq`` were enqueued for the lane."""
        if upto_seq > self._lane_seq.get(ctx.lane, 0):
            self._lane_seq[ctx.lane] = upto_seq
        if tail_event is not None:
            ctx.tail_event = tail_event

    def release_lane(self, nonce: str) -> Optional[int]:
        """End of request: the lane returns to the free list once its last enqueued kernel has run."""
        ctx = self._lanes.pop(nonce, None)
        if ctx is None or ctx.lane < 0:
            return None
        lane = ctx.lane
        if ctx.quiescent():
            self._free_lanes.append(lane)
        else:
            self._draining.append(ctx)
        return lane

    def reap_lanes(self) -> int:
        done = [c for c in self._draining if c.quiescent()]
        for c in done:
            self._draining.remove(c)
            self._free_lanes.append(c.lane)
        return len(done)


__all__ = ["StreamManager", "StreamContext"]
