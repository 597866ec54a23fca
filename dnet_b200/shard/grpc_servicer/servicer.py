"""dnetring.DnetRingService on a shard (reference src/dnet/shard/grpc_servicer/servicer.py:21-161).

Same five RPCs and response contracts.  Two additions ride inside the unchanged wire contract:
``SendActivation`` with dtype ``b200.hop.open`` answers with this shard's hop endpoint (CUDA IPC
handles as JSON in ``ActivationResponse.message``) so the predecessor can map our receive lanes,
and an ``end_of_request`` frame releases the nonce's lane before the stream closes."""
from __future__ import annotations

import asyncio
import json
import time

from dnet_b200.protos import dnet_ring_pb2 as pb2
from dnet_b200.protos.dnet_ring_pb2_grpc import DnetRingServiceServicer
from dnet_b200.utils.logger import logger


class GrpcServicer(DnetRingServiceServicer):
    def __init__(self, shard):
        self.shard = shard

    async def SendActivation(self, request, context):
        try:
            if request.activation.dtype == "b200.hop.open":
                rt = self.shard.runtime
                hops = int(request.activation.shape[0]) if len(request.activation.shape) else 0
                if hops > 0:      # ring census: relay the question `hops` shards further along the ring
                    stub = getattr(self.shard.adapter, "next_node_stub", None)
                    if stub is None:
                        return pb2.ActivationResponse(success=False, message="no next node", node_id=str(self.shard.node_id))
                    fwd = pb2.ActivationRequest()
                    fwd.CopyFrom(request)
                    del fwd.activation.shape[:]
                    fwd.activation.shape.append(hops - 1)
                    return await stub.SendActivation(fwd, timeout=10.0)
                hop = getattr(rt, "hop", None) or getattr(rt, "hop_pending", None)
                if hop is None:
                    return pb2.ActivationResponse(success=False, message="no hop lanes on this shard (yet)",
                                                  node_id=str(self.shard.node_id))
                ep = hop.endpoint()
                ep["grpc_addr"] = getattr(self.shard.adapter, "advertise_addr", None)
                return pb2.ActivationResponse(success=True, message=json.dumps(ep), node_id=str(self.shard.node_id))
            await self.shard.admit_frame(request)
            return pb2.ActivationResponse(success=True, message="Activation processed successfully",
                                          node_id=str(self.shard.node_id))
        except Exception as e:
            logger.error("Error processing activation request: %s", e)
            return pb2.ActivationResponse(success=False, message=f"Error: {e}", node_id=str(self.shard.node_id))

    async def HealthCheck(self, request, context):
        rt = self.shard.runtime
        return pb2.HealthResponse(healthy=bool(self.shard.adapter.running), node_id=str(self.shard.node_id),
                                  assigned_layers=list(rt.assigned_layers), queue_size=rt.activation_recv_queue.qsize(),
                                  active_requests=0)

    async def ResetCache(self, request, context):
        try:
            await self.shard.reset_cache()
            return pb2.ResetCacheResponse(success=True, message="Activation processed successfully")
        except Exception as e:
            logger.error("Error processing reset-cache request: %s", e)
            return pb2.ResetCacheResponse(success=False, message=f"Error: {e}")

    async def MeasureLatency(self, request, context):
        now_ms = int(time.time() * 1000)
        try:
            return pb2.LatencyMeasureResponse(success=True, message="Latency measurement response",
                                              node_id=str(self.shard.node_id), timestamp=now_ms)
        except Exception as e:
            return pb2.LatencyMeasureResponse(success=False, message=f"Error: {e}", node_id=str(self.shard.node_id),
                                              timestamp=now_ms)

    async def StreamActivations(self, request_iterator, context):
        """One ACK per frame; a frame without a nonce is refused; end_of_request is acknowledged ("eor"),
        releases the nonce's hop lane and closes the stream."""
        try:
            async for frame in request_iterator:
                if frame.end_of_request:
                    try:
                        await self.shard.end_request(frame.request.nonce)
                    except Exception:
                        pass
                    yield pb2.StreamAck(nonce=frame.request.nonce, seq=frame.seq, accepted=True, message="eor")
                    break
                req = frame.request
                if not req.nonce:
                    yield pb2.StreamAck(nonce="", seq=frame.seq, accepted=False, message="missing nonce")
                    continue
                await self.shard.admit_frame(req)
                yield pb2.StreamAck(nonce=req.nonce, seq=frame.seq, accepted=True)
        except asyncio.CancelledError:
            logger.debug("[STREAM][RX] cancelled")
            return
        except Exception as e:
            logger.error("[STREAM][RX] error: %s", e)
            import grpc

            await context.abort(grpc.StatusCode.INTERNAL, str(e))
