"""Core message DTOs (reference src/dnet/core/types/messages.py:16-126), mlx-free."""
from __future__ import annotations

from dataclasses import dataclass, field
from enum import Enum
from typing import Any, NamedTuple, Optional, Tuple


@dataclass(slots=True)
class ActivationMessage:
    nonce: str
    pool_id: int
    batch_size: int
    shape: Tuple[int, ...]
    dtype: str
    layer_id: int
    timestamp: int
    node_origin: str
    callback_url: str
    # Optional direct tensor reference (device tensor) to avoid staging/copies
    tensor: Optional[Any] = None
    recv_perf_t: float = 0.0
    enq_perf_t: float = 0.0
    tx_enq_perf_t: float = 0.0
    # Final token path (end-shard sampling)
    is_final: bool = False
    token_id: int = -1
    logprob: float = 0.0
    top_logprobs: Optional[dict[int, float]] = None
    # Request control
    req_logprobs: bool = False
    req_top_logprobs: int = 0
    # Decoding parameters
    temperature: float = 1.0
    top_p: float = 1.0
    top_k: int = -1
    repetition_penalty: float = 1.0
    min_p: float = 0.0
    min_tokens_to_keep: int = 1
    # dnet_b200 extension: torch.cuda.Event recorded after the kernels that produce ``tensor``
    # were enqueued; a consumer on another stream orders itself after it (device hand-off)
    ready_event: Optional[Any] = None
    # dnet_b200 device-hop transport (shard/frames.py, shard/adapters/ring.py):
    lane: int = -1                       # hop lane of the nonce (same index on every shard), -1 = none
    seq0: int = 0                        # decode-flag value the request's first on-device decode step waits for
    hop_wait: Optional[Tuple[int, int, int]] = None   # (arrival flag ptr, seq, consumed flag ptr): tensor is a hop slot
    sched: Optional[Any] = None          # b200.sched: ordered [(lane, seq)] decode entries
    sched_done: Optional[Any] = None     # SchedTicket: lets the head's scheduler bound the frames in flight
    hop_meta: Optional[Any] = None       # set by the egress hook once the tensor went out over NVLink

    @classmethod
    def from_proto(cls, proto_msg, pool_id: int = 0):
        """Create from a dnetring.ActivationRequest (reference messages.py:51-77)."""
        a = proto_msg.activation
        return cls(
            nonce=proto_msg.nonce,
            pool_id=pool_id,
            batch_size=a.batch_size,
            shape=tuple(a.shape),
            dtype=a.dtype,
            layer_id=a.layer_id,
            timestamp=proto_msg.timestamp,
            node_origin=proto_msg.node_origin,
            callback_url=proto_msg.callback_url,
            req_logprobs=proto_msg.logprobs,
            req_top_logprobs=proto_msg.top_logprobs,
            temperature=proto_msg.temperature if proto_msg.HasField("temperature") else 1.0,
            top_p=proto_msg.top_p if proto_msg.HasField("top_p") else 1.0,
            top_k=proto_msg.top_k if proto_msg.HasField("top_k") else -1,
            repetition_penalty=proto_msg.repetition_penalty if proto_msg.HasField("repetition_penalty") else 1.0,
            min_p=proto_msg.min_p if proto_msg.HasField("min_p") else 0.0,
            min_tokens_to_keep=proto_msg.min_tokens_to_keep if proto_msg.HasField("min_tokens_to_keep") else 1,
        )

    def to_proto(self, data: bytes):
        """Convert to a dnetring.ActivationRequest (reference messages.py:79-101)."""
        from dnet_b200.protos import dnet_ring_pb2 as pb

        return pb.ActivationRequest(
            nonce=self.nonce,
            activation=pb.Activation(
                data=data, batch_size=self.batch_size, shape=list(self.shape),
                layer_id=self.layer_id, dtype=self.dtype),
            timestamp=self.timestamp,
            node_origin=self.node_origin,
            callback_url=self.callback_url,
            logprobs=self.req_logprobs,
            top_logprobs=self.req_top_logprobs,
            temperature=self.temperature,
            top_p=self.top_p,
            top_k=self.top_k,
            repetition_penalty=self.repetition_penalty,
            min_p=self.min_p,
            min_tokens_to_keep=self.min_tokens_to_keep,
        )


@dataclass(slots=True)
class WeightRequest:
    weight_id: str
    layer_id: int
    priority: int = 0


class PoolStatus(str, Enum):
    FREE = "free"
    ALLOCATED = "allocated"
    IN_USE = "in_use"


class StopCondition(NamedTuple):
    stop_met: bool
    trim_length: int


@dataclass
class TokenResult:
    token_id: int
    logprob: float = 0.0
    top_logprobs: dict[int, float] = field(default_factory=dict)


__all__ = ["ActivationMessage", "WeightRequest", "PoolStatus", "StopCondition", "TokenResult"]
