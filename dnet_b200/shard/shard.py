"""Topology/strategy-agnostic shard: runtime + adapter glue (reference src/dnet/shard/shard.py:26-71).

    runtime = ShardRuntime(shard_id)
    adapter = RingAdapter(runtime, discovery)
    shard = Shard(shard_id, adapter)
    servicer = GrpcServicer(shard)            # shard/grpc_servicer
"""
from __future__ import annotations

import asyncio

from .adapters.base import TopologyAdapter
from .models import ShardLoadModelResponse, ShardUnloadModelResponse
from .runtime import ShardRuntime


class Shard:
    def __init__(self, shard_id, adapter: TopologyAdapter):
        self.node_id = shard_id
        self.adapter = adapter
        self.runtime: ShardRuntime = adapter.runtime

    async def start(self, loop: asyncio.AbstractEventLoop) -> None:
        self.runtime.attach_loop(loop)
        self.runtime.start()          # compute thread
        await self.adapter.start()

    async def shutdown(self) -> None:
        await self.adapter.shutdown()
        self.runtime.shutdown()

    async def admit_frame(self, request) -> None:
        """Queue one ring frame for the adapter's ingress worker; yields while the queue is full and
        drops the frame once the adapter stopped."""
        q = self.adapter.ingress_q
        while self.adapter.running:
            try:
                q.put_nowait(request)
                return
            except asyncio.QueueFull:
                await asyncio.sleep(0)

    async def end_request(self, nonce: str) -> None:
        ender = getattr(self.adapter, "end_request", None)
        if ender is not None:
            await ender(nonce)

    async def reset_cache(self):
        self.runtime.reset_cache()

    async def load_model(self, req) -> ShardLoadModelResponse:
        loop = asyncio.get_running_loop()
        await loop.run_in_executor(self.runtime.executor, self.runtime.load_model_core, req)
        await self.adapter.configure_topology(req)
        return ShardLoadModelResponse(success=True, message="Model loaded successfully",
                                      layers_loaded=self.runtime.assigned_layers, load_time_ms=100)

    async def unload_model(self) -> ShardUnloadModelResponse:
        await self.adapter.reset_topology()
        model_path = self.runtime.model_path
        response = self.runtime.unload_model_core()
        if response.success and isinstance(model_path, str):
            try:
                from dnet_b200.utils.repack import delete_repacked_layers

                delete_repacked_layers(current_model_path=model_path)
            except Exception:
                pass
        return response

    def queue_size(self) -> int:
        return self.runtime.queue_size()
