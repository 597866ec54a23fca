"""Metadata-only ring frames of the device-hop transport.

The wire contract (src/dnet/protos/dnet_ring.proto:23-68) is unchanged: these are ordinary
``ActivationRequest`` messages whose ``activation.dtype`` names a dnet_b200 frame kind and whose
``activation.data`` carries a small little-endian header instead of tensor bytes.  A reference
shard that receives one rejects it in its codec (unknown dtype) exactly like any other malformed
frame; a dnet_b200 shard that talks to a reference peer never emits them (no hop link -> bytes path).

  b200.hop/<dtype>   the tensor of this frame already sits in the receiver's lane slot (device hop);
                     header = lane, arrival seq, decode seq0 of the request
  b200.sched         ordered decode schedule made by the head shard: (lane, seq) pairs; every shard
                     launches one fused step-hop kernel per pair, in this order
  b200.lease         API -> head shard: decode ``steps`` more tokens of the nonce with the token loop
                     closed on the device (optionally seeding the token)
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

MAGIC = b"DNB2"
HOP_PREFIX = "b200.hop/"
SCHED_DTYPE = "b200.sched"
LEASE_DTYPE = "b200.lease"

BUBBLE = 0xFFFFFFFF            # schedule entry without a request: shards only serve due head parts (tensor-parallel lm_head)

_HOP = struct.Struct("<4sIIII")      # magic, lane, seq (bulk arrival), seq0 (first decode seq), flags
_LEASE = struct.Struct("<4sIIiI")    # magic, steps, has_token, token, lane_hint(unused)


@dataclass
class HopMeta:
    lane: int
    seq: int          # bulk-flag value that announces this frame's tensor
    seq0: int         # decode-flag value the request's first on-device decode step waits for
    flags: int = 0


def is_hop(dtype: str) -> bool:
    return dtype.startswith(HOP_PREFIX)


def hop_dtype(wire_dtype: str) -> str:
    return HOP_PREFIX + wire_dtype


def hop_wire_dtype(dtype: str) -> str:
    return dtype[len(HOP_PREFIX):]


def pack_hop(m: HopMeta) -> bytes:
    return _HOP.pack(MAGIC, m.lane, m.seq, m.seq0, m.flags)


def unpack_hop(data: bytes) -> HopMeta:
    if len(data) != _HOP.size:
        raise ValueError(f"hop frame header must be {_HOP.size} bytes, got {len(data)}")
    magic, lane, seq, seq0, flags = _HOP.unpack(data)
    if magic != MAGIC:
        raise ValueError("bad hop frame magic")
    return HopMeta(lane, seq, seq0, flags)


def pack_sched(entries: Sequence[Tuple[int, int]]) -> bytes:
    flat: List[int] = []
    for lane, seq in entries:
        flat += [int(lane), int(seq)]
    return MAGIC + struct.pack(f"<I{len(flat)}I", len(entries), *flat)


def unpack_sched(data: bytes) -> List[Tuple[int, int]]:
    if data[:4] != MAGIC or len(data) < 8:
        raise ValueError("bad schedule frame")
    (n,) = struct.unpack_from("<I", data, 4)
    if len(data) != 8 + 8 * n:
        raise ValueError(f"schedule frame length {len(data)} does not match {n} entries")
    flat = struct.unpack_from(f"<{2 * n}I", data, 8)
    return [(flat[2 * i], flat[2 * i + 1]) for i in range(n)]


def pack_lease(steps: int, token: Optional[int] = None) -> bytes:
    return _LEASE.pack(MAGIC, int(steps), 0 if token is None else 1, 0 if token is None else int(token), 0)


def unpack_lease(data: bytes) -> Tuple[int, Optional[int]]:
    if len(data) != _LEASE.size:
        raise ValueError("bad lease frame")
    magic, steps, has_token, token, _ = _LEASE.unpack(data)
    if magic != MAGIC:
        raise ValueError("bad lease frame magic")
    return steps, (token if has_token else None)


class SchedTicket:
    """Completion handle of one schedule frame on the head shard.  The compute thread records a CUDA
    event behind the frame's last launch (``record``); the scheduler polls ``is_set()`` to keep at most
    ``sched_frames_in_flight`` frames of kernels queued on the GPU."""

    __slots__ = ("_event", "_recorded")

    def __init__(self):
        self._event = None
        self._recorded = False

    def record(self, event=None) -> None:
        self._event = event
        self._recorded = True

    def is_set(self) -> bool:
        if not self._recorded:
            return False
        ev = self._event
        return True if ev is None else bool(ev.query())
