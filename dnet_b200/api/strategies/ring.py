"""API-side ring adapter: the token loop as the API sees it (reference
src/dnet/api/strategies/ring.py:101-209; the HALDA solver half of that file is consumed unchanged
and not rebuilt).

``send_tokens`` / ``await_token`` / ``resolve_token`` keep the reference's names and meaning.  Two
additions for the device-closed loop: ``lease(nonce, steps)`` lets the ring
decode ``steps`` tokens without the API in the loop -- tokens then arrive unsolicited through
``resolve_token`` and queue up per nonce until ``await_token`` takes them in order -- and
``end_request`` releases the nonce's lanes on every shard (an ``end_of_request`` frame)."""
from __future__ import annotations

import asyncio
from collections import deque
from typing import Any, Deque, Dict, Optional

from dnet_b200.core.stream_manager import StreamManager
from dnet_b200.core.types.messages import ActivationMessage, TokenResult
from dnet_b200.protos import dnet_ring_pb2 as pb2
from dnet_b200.shard import frames as fr
from dnet_b200.utils.logger import logger
from dnet_b200.utils.time import utc_epoch_now


class RingApiAdapter:
    def __init__(self) -> None:
        self.running = False
        self.channel: Optional[Any] = None
        self.stub: Optional[Any] = None
        self._streams = StreamManager(idle_timeout_s=5.0, backoff_s=0.2)
        self._pending: Dict[str, asyncio.Future] = {}
        self._arrived: Dict[str, Deque[TokenResult]] = {}
        self._loop: Optional[asyncio.AbstractEventLoop] = None

    async def start(self) -> None:
        self.running = True
        self._loop = asyncio.get_running_loop()

    async def shutdown(self) -> None:
        self.running = False
        for nonce in list(self._streams._streams.keys()):
            try:
                await self._streams.end_stream(nonce)
            except Exception:
                pass
        if self.channel is not None:
            try:
                await self.channel.close()
            except Exception:
                pass
        self.channel = None
        self.stub = None

    async def connect_first_shard(self, ip: str, port: int) -> None:
        from grpc import aio as aio_grpc

        from dnet_b200.protos.dnet_ring_pb2_grpc import DnetRingServiceStub

        if self.channel is not None:
            try:
                await self.channel.close()
            except Exception:
                pass
        target = f"{ip}:{port}"
        self.channel = aio_grpc.insecure_channel(target)
        self.stub = DnetRingServiceStub(self.channel)
        logger.info("Connected API adapter to first shard at %s", target)

    async def reset_cache(self) -> None:
        if not self.stub:
            raise RuntimeError("API adapter not connected to a shard")
        try:
            await self.stub.ResetCache(pb2.ResetCacheRequest())
        except Exception as e:
            logger.warning("ResetCache RPC failed: %s", e)

    async def _put(self, nonce: str, req, end: bool = False) -> None:
        if not self.stub:
            raise RuntimeError("Ring adapter not connected to first shard")
        stub = self.stub
        ctx = await self._streams.get_or_create_stream(nonce, lambda it: stub.StreamActivations(it))
        if not ctx or not ctx.open:
            raise RuntimeError(f"Failed to create stream for nonce {nonce}")
        ctx.last_seq += 1
        await ctx.queue.put(pb2.ActivationFrame(request=req, seq=ctx.last_seq, end_of_request=end))
        ctx.touch()

    def _request(self, nonce: str, dtype: str, data: bytes, callback_addr: str, logprobs: bool, top_logprobs: int,
                 decoding_config: Optional[Any], more: bool = False):
        d = decoding_config
        msg = ActivationMessage(
            nonce=nonce, pool_id=-1, batch_size=0 if more else 1, shape=(1,), dtype=dtype, layer_id=-1, timestamp=utc_epoch_now(),
            node_origin="api", callback_url=callback_addr if "://" in callback_addr else f"grpc://{callback_addr}",
            req_logprobs=logprobs, req_top_logprobs=top_logprobs,
            temperature=d.temperature if d else 1.0, top_p=d.top_p if d else 1.0, top_k=d.top_k if d else -1,
            repetition_penalty=d.repetition_penalty if d else 1.0, min_p=d.min_p if d else 0.0,
            min_tokens_to_keep=d.min_tokens_to_keep if d else 1)
        return msg.to_proto(data)

    async def send_tokens(self, nonce: str, tokens: bytes, callback_addr: str, logprobs: bool = False,
                          top_logprobs: int = 0, decoding_config: Optional[Any] = None, more: bool = False) -> None:
        """int32 token ids (prompt, or one sampled token in the host-closed loop) to the first shard.
        ``more=True`` marks a prompt chunk that is not the last (chunked prefill): the ring fills the KV cache
        and returns no token for it."""
        await self._put(nonce, self._request(nonce, "tokens", tokens, callback_addr, logprobs, top_logprobs,
                                             decoding_config, more=more))

    async def lease(self, nonce: str, steps: int, callback_addr: str = "", token: Optional[int] = None) -> None:
        """Let the ring decode ``steps`` more tokens of ``nonce`` with the token loop closed on the device."""
        await self._put(nonce, self._request(nonce, fr.LEASE_DTYPE, fr.pack_lease(steps, token), callback_addr, False, 0, None))

    async def end_request(self, nonce: str) -> None:
        try:
            await self._put(nonce, self._request(nonce, "tokens", b"", "", False, 0, None), end=True)
        finally:
            self._arrived.pop(nonce, None)

    async def await_token(self, nonce: str, timeout_s: float) -> TokenResult:
        q = self._arrived.get(nonce)
        if q:
            return q.popleft()
        fut = asyncio.get_running_loop().create_future()
        self._pending[nonce] = fut
        try:
            return await asyncio.wait_for(fut, timeout=timeout_s)
        finally:
            self._pending.pop(nonce, None)

    def resolve_token(self, nonce: str, result: TokenResult) -> None:
        """Called by ShardApiServicer.SendToken (or an in-process token sink, from any thread)."""
        loop = self._loop
        if loop is not None and not _in_loop(loop):
            loop.call_soon_threadsafe(self._resolve, nonce, result)
        else:
            self._resolve(nonce, result)

    def _resolve(self, nonce: str, result: TokenResult) -> None:
        fut = self._pending.get(nonce)
        if fut is not None and not fut.done():
            fut.set_result(result)
        else:
            self._arrived.setdefault(nonce, deque()).append(result)


def _in_loop(loop) -> bool:
    try:
        return asyncio.get_running_loop() is loop
    except RuntimeError:
        return False
