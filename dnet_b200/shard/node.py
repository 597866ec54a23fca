"""Process-level assembly of a shard and of an API client, for synchronous drivers.

``ShardNode`` is what the reference's ``dnet-shard`` entry point wires together
(src/cli/shard.py: ShardRuntime -> RingAdapter -> Shard -> gRPC server on one asyncio loop); the
HTTP control surface (/load_model ...) stays out of scope, so ``load_model`` / ``unload_model`` are
plain method calls that run the same ``Shard.load_model`` coroutine.  ``ApiNode`` is the API
node's token loop (RingApiAdapter + InferenceManager + the SendToken server) without HTTP.  Both
own a background event loop so tests, ``bench.py`` and notebooks can drive them synchronously.
"""
from __future__ import annotations

import asyncio
import threading
from typing import Any, AsyncIterator, Callable, List, Optional, Sequence

from dnet_b200.core.types.messages import ActivationMessage, TokenResult


class _LoopThread:
    def __init__(self, name: str):
        self.loop = asyncio.new_event_loop()
        self._thread = threading.Thread(target=self._run, name=name, daemon=True)
        self._thread.start()

    def _run(self) -> None:
        asyncio.set_event_loop(self.loop)
        self.loop.run_forever()

    def call(self, coro, timeout: Optional[float] = 120.0):
        return asyncio.run_coroutine_threadsafe(coro, self.loop).result(timeout)

    def stop(self) -> None:
        self.loop.call_soon_threadsafe(self.loop.stop)
        self._thread.join(timeout=5)


class ShardNode:
    def __init__(self, shard_id, grpc_port: int, host: str = "127.0.0.1", queue_size: int = 128, transport_settings=None,
                 runtime=None):
        from .adapters.ring import RingAdapter
        from .grpc_servicer import GrpcServer
        from .runtime import ShardRuntime
        from .shard import Shard

        self.grpc_port, self.host = grpc_port, host
        self.runtime = runtime or ShardRuntime(shard_id=shard_id, queue_size=queue_size)
        self._lt = _LoopThread(f"dnet-shard-{shard_id}")
        self.adapter = self._lt.call(self._make(lambda: RingAdapter(self.runtime, None, transport_settings)))
        self.adapter.advertise_addr = f"{host}:{grpc_port}"
        self.shard = Shard(shard_id, self.adapter)
        self.server = GrpcServer(grpc_port, self.shard, host=host)
        self._started = False

    @staticmethod
    async def _make(factory):
        return factory()      # asyncio queues are created on the loop that will use them

    def call(self, coro, timeout: Optional[float] = 120.0):
        return self._lt.call(coro, timeout)

    def start(self) -> "ShardNode":
        async def go():
            await self.shard.start(asyncio.get_running_loop())
            await self.server.start()
        self.call(go())
        self._started = True
        return self

    def load_model(self, req, timeout: float = 1800.0):
        return self.call(self.shard.load_model(req), timeout)

    def unload_model(self):
        return self.call(self.shard.unload_model())

    def shutdown(self) -> None:
        if self._started:
            async def stop():
                await self.server.shutdown()
                await self.shard.shutdown()
            try:
                self.call(stop(), 30)
            except Exception:
                pass
            self._started = False
        self._lt.stop()


class ApiNode:
    """Token-id level API client of a ring: ``generate`` drives one request, ``submit``/``collect`` many
    concurrently (one nonce each -> one hop lane each -> the shards overlap)."""

    def __init__(self, first_shard_addr: str, callback: str = "local://", grpc_port: int = 0, host: str = "127.0.0.1"):
        from dnet_b200.api.inference import InferenceManager
        from dnet_b200.api.strategies.ring import RingApiAdapter

        self._lt = _LoopThread("dnet-api")
        self.adapter = self._lt.call(ShardNode._make(RingApiAdapter))
        self.callback = callback if callback.startswith("local://") else f"{host}:{grpc_port}"
        self.manager = InferenceManager(self.adapter, self.callback)
        self._server = None
        ip, port = first_shard_addr.rsplit(":", 1)

        async def go():
            await self.adapter.start()
            await self.adapter.connect_first_shard(ip, int(port))
            if not callback.startswith("local://"):
                from dnet_b200.api.grpc_servicer import ShardApiServer

                self._server = ShardApiServer(grpc_port, self.manager, host=host)
                await self._server.start()
        self._lt.call(go())

    def token_sink(self, msg: ActivationMessage) -> None:
        """In-process delivery (callback_url local://): hand this to the finalising shard's
        ``RingAdapter.token_sink``; it is what ShardApiServicer.SendToken does after the RPC."""
        self.manager.resolve_request(msg.nonce, TokenResult(token_id=int(msg.token_id), logprob=float(msg.logprob),
                                                            top_logprobs=dict(msg.top_logprobs or {})))

    def call(self, coro, timeout: Optional[float] = 600.0):
        return self._lt.call(coro, timeout)

    def generate(self, nonce: str, prompt: Sequence[int], max_tokens: int, **kw) -> List[TokenResult]:
        async def run():
            return [r async for r in self.manager.generate_stream(nonce, prompt, max_tokens, **kw)]
        return self.call(run())

    def generate_many(self, prompts: Sequence[Sequence[int]], max_tokens: int, prefix: str = "req", **kw) -> List[List[TokenResult]]:
        async def one(i):
            return [r async for r in self.manager.generate_stream(f"{prefix}{i}", prompts[i], max_tokens, **kw)]

        async def run():
            return await asyncio.gather(*[one(i) for i in range(len(prompts))])
        return self.call(run())

    def shutdown(self) -> None:
        async def stop():
            if self._server is not None:
                await self._server.shutdown()
            await self.adapter.shutdown()
        try:
            self.call(stop(), 30)
        except Exception:
            pass
        self._lt.stop()
