"""shardapi.ShardApiService on the API node: shards deliver sampled tokens here
(reference src/dnet/api/grpc_servicer/servicer.py:10-37, server.py)."""
from __future__ import annotations

from typing import Optional

from dnet_b200.core.types.messages import TokenResult
from dnet_b200.protos import shard_api_comm_pb2 as pb2
from dnet_b200.protos.shard_api_comm_pb2_grpc import ShardApiServiceServicer, add_ShardApiServiceServicer_to_server
from dnet_b200.utils.logger import logger


class ShardApiServicer(ShardApiServiceServicer):
    def __init__(self, inference_manager) -> None:
        self.inference_manager = inference_manager

    async def SendToken(self, request, context):
        try:
            self.inference_manager.resolve_request(
                request.nonce, TokenResult(token_id=int(request.token_id), logprob=float(request.logprob),
                                           top_logprobs=dict(request.top_logprobs)))
            return pb2.TokenResponse(success=True, message="Token received")
        except Exception as e:
            logger.error("Error handling token: %s", e)
            return pb2.TokenResponse(success=False, message=str(e))

    async def SendFinalActivation(self, request, context):
        return pb2.FinalActivationResponse(success=False, message="not supported: shards sample the token", token_id=-1)


class ShardApiServer:
    def __init__(self, grpc_port: int, inference_manager, host: str = "[::]") -> None:
        self.grpc_port, self.host = grpc_port, host
        self.servicer = ShardApiServicer(inference_manager)
        self.server: Optional[object] = None

    async def start(self) -> None:
        from grpc import aio as aio_grpc

        self.server = aio_grpc.server()
        add_ShardApiServiceServicer_to_server(self.servicer, self.server)
        self.server.add_insecure_port(f"{self.host}:{self.grpc_port}")
        await self.server.start()

    async def shutdown(self) -> None:
        if self.server is not None:
            await self.server.stop(grace=1)
            self.server = None
