"""gRPC client stub + server registration for dnetring.DnetRingService, written by hand against
grpcio's generic API (the reference generates this file with grpc_tools, which is not installed;
service and method paths are the ones src/dnet/protos/dnet_ring.proto:6-21 declares, so a
reference peer can call this server and this stub can call a reference shard)."""
from __future__ import annotations

import grpc

from . import dnet_ring_pb2 as pb

_UNARY = {
    "SendActivation": (pb.ActivationRequest, pb.ActivationResponse),
    "HealthCheck": (pb.HealthRequest, pb.HealthResponse),
    "ResetCache": (pb.ResetCacheRequest, pb.ResetCacheResponse),
    "MeasureLatency": (pb.LatencyMeasureRequest, pb.LatencyMeasureResponse),
}


class DnetRingServiceStub:
    """Client side: one callable per RPC, bound to a (sync or aio) channel."""

    def __init__(self, channel):
        for name, (req, resp) in _UNARY.items():
            setattr(self, name, channel.unary_unary(pb.METHODS[name], request_serializer=req.SerializeToString,
                                                    response_deserializer=resp.FromString))
        self.StreamActivations = channel.stream_stream(pb.METHODS["StreamActivations"],
                                                       request_serializer=pb.ActivationFrame.SerializeToString,
                                                       response_deserializer=pb.StreamAck.FromString)


class DnetRingServiceServicer:
    """Server side base: override the RPCs you serve."""

    async def SendActivation(self, request, context):
        raise NotImplementedError

    async def HealthCheck(self, request, context):
        raise NotImplementedError

    async def ResetCache(self, request, context):
        raise NotImplementedError

    async def MeasureLatency(self, request, context):
        raise NotImplementedError

    def StreamActivations(self, request_iterator, context):
        raise NotImplementedError


def add_DnetRingServiceServicer_to_server(servicer, server) -> None:
    handlers = {
        name: grpc.unary_unary_rpc_method_handler(getattr(servicer, name), request_deserializer=req.FromString,
                                                  response_serializer=resp.SerializeToString)
        for name, (req, resp) in _UNARY.items()
    }
    handlers["StreamActivations"] = grpc.stream_stream_rpc_method_handler(
        servicer.StreamActivations, request_deserializer=pb.ActivationFrame.FromString,
        response_serializer=pb.StreamAck.SerializeToString)
    server.add_generic_rpc_handlers((grpc.method_handlers_generic_handler(pb.SERVICE, handlers),))
