"""grpc.aio server hosting the ring service (reference src/dnet/shard/grpc_servicer/server.py:9-36)."""
from __future__ import annotations

from typing import Optional

from grpc import aio as aio_grpc

from dnet_b200.protos.dnet_ring_pb2_grpc import add_DnetRingServiceServicer_to_server
from dnet_b200.utils.logger import logger
from .servicer import GrpcServicer

GRPC_OPTIONS = [("grpc.max_send_message_length", 64 * 1024 * 1024), ("grpc.max_receive_message_length", 64 * 1024 * 1024)]


class GrpcServer:
    def __init__(self, grpc_port: int, shard, host: str = "[::]") -> None:
        self.grpc_port = grpc_port
        self.host = host
        self.shard = shard
        self.server: Optional[aio_grpc.Server] = None
        self.servicer = GrpcServicer(shard)

    async def start(self) -> None:
        self.server = aio_grpc.server(options=GRPC_OPTIONS)
        add_DnetRingServiceServicer_to_server(self.servicer, self.server)
        listen_addr = f"{self.host}:{self.grpc_port}"
        self.server.add_insecure_port(listen_addr)
        await self.server.start()
        logger.info("gRPC server started on %s", listen_addr)

    async def shutdown(self) -> None:
        if self.server:
            await self.server.stop(grace=1)
            self.server = None
