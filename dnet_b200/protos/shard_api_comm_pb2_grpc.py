"""gRPC stub + server registration for shardapi.ShardApiService (reference
src/dnet/protos/shard_api_comm.proto:6-13), hand-written over grpcio's generic API."""
from __future__ import annotations

import grpc

from . import shard_api_comm_pb2 as pb

_UNARY = {
    "SendFinalActivation": (pb.FinalActivationRequest, pb.FinalActivationResponse),
    "SendToken": (pb.TokenRequest, pb.TokenResponse),
}


class ShardApiServiceStub:
    def __init__(self, channel):
        for name, (req, resp) in _UNARY.items():
            setattr(self, name, channel.unary_unary(pb.METHODS[name], request_serializer=req.SerializeToString,
                                                    response_deserializer=resp.FromString))


class ShardApiServiceServicer:
    async def SendFinalActivation(self, request, context):
        raise NotImplementedError

    async def SendToken(self, request, context):
        raise NotImplementedError


def add_ShardApiServiceServicer_to_server(servicer, server) -> None:
    handlers = {
        name: grpc.unary_unary_rpc_method_handler(getattr(servicer, name), request_deserializer=req.FromString,
                                                  response_serializer=resp.SerializeToString)
        for name, (req, resp) in _UNARY.items()
    }
    server.add_generic_rpc_handlers((grpc.method_handlers_generic_handler(pb.SERVICE, handlers),))
