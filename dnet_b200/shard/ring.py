"""Ring hop data plane over NVLink 5 / NVSwitch (replaces the tensor bytes of
RingAdapter._send_activation + DnetRingService.StreamActivations, reference
src/dnet/shard/adapters/ring.py:265-299, shard/grpc_servicer/servicer.py:129-161).

One process per GPU, like dnet-shard.  The RECEIVING shard owns, per in-flight nonce slot,
an HBM activation slot and a 32-bit sequence flag; both are exported with CUDA IPC once at
``configure_topology`` time.  A hop is ``dn_hop_send``: one peer ``cudaMemcpyAsync`` into the
next shard's slot on the sender's stream followed by a system-scope release store of the
sequence number; the receiver's compute stream runs ``dn_hop_wait`` (a one-thread acquire
spin, bounded by a timeout so a dead peer can never hang the GPU) and then computes straight
out of the slot.  No host thread, protobuf, HTTP/2 or TCP on the tensor path; the gRPC
frame of the reference still carries nonce / decoding parameters / ACKs (control plane).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch

from dnet_b200 import _cabi


class _CudaView:
    """Expose a raw device pointer through __cuda_array_interface__ so torch can alias it."""

    def __init__(self, ptr: int, shape: Tuple[int, ...], typestr: str):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}


def device_view(ptr: int, shape: Tuple[int, ...], dtype: torch.dtype) -> torch.Tensor:
    """A torch tensor aliasing ``ptr`` (no ownership)."""
    if dtype == torch.bfloat16:
        n = 1
        for s in shape:
            n *= s
        raw = torch.as_tensor(_CudaView(ptr, (n,), "<i2"), device="cuda")
        return raw.view(torch.bfloat16).view(shape)
    typestr = {torch.int32: "<i4", torch.uint8: "|u1", torch.float32: "<f4", torch.int16: "<i2"}[dtype]
    return torch.as_tensor(_CudaView(ptr, shape, typestr), device="cuda")


@dataclass
class HopEndpoint:
    """What a shard publishes to its ring predecessor (picklable: two 64-byte IPC handles)."""
    data_handle: bytes
    flag_handle: bytes
    n_slots: int
    slot_bytes: int


class HopReceiver:
    """Receiver-owned slots: ``n_slots`` x ``slot_bytes`` of HBM, one uint32 *arrival* flag and one
    *consumed* flag each (one 64-byte line per flag), + 1 error flag.  The consumed flag is written
    by the receiver's compute stream once it has copied a bulk slot out; the sender polls it over
    NVLink before overwriting the slot (credit for single-buffered prefill chunks)."""

    def __init__(self, n_slots: int, slot_bytes: int):
        self.lib = _cabi.load()
        self.n_slots, self.slot_bytes = n_slots, slot_bytes
        d, f = C.c_void_p(), C.c_void_p()
        _cabi.check(self.lib.dn_hop_alloc(n_slots * slot_bytes, C.byref(d)))
        _cabi.check(self.lib.dn_hop_alloc((2 * n_slots + 1) * 64, C.byref(f)))   # one 64-byte line per flag
        self.data_ptr, self.flag_ptr = d.value, f.value

    def slot(self, i: int) -> int:
        return self.data_ptr + i * self.slot_bytes

    def flag(self, i: int) -> int:
        return self.flag_ptr + i * 64

    def consumed_flag(self, i: int) -> int:
        return self.flag_ptr + (self.n_slots + i) * 64

    @property
    def err_flag(self) -> int:
        return self.flag_ptr + 2 * self.n_slots * 64

    def mark_consumed(self, i: int, seq: int, stream: int) -> None:
        """publish on the compute stream that slot i's contents up to ``seq`` were copied out"""
        _cabi.check(self.lib.dn_hop_send(self.slot(i), self.slot(i), 0, self.consumed_flag(i), seq, stream))

    def endpoint(self) -> HopEndpoint:
        hd, hf = (C.c_uint8 * 64)(), (C.c_uint8 * 64)()
        _cabi.check(self.lib.dn_hop_export(self.data_ptr, hd))
        _cabi.check(self.lib.dn_hop_export(self.flag_ptr, hf))
        return HopEndpoint(bytes(hd), bytes(hf), self.n_slots, self.slot_bytes)

    def wait(self, i: int, seq: int, stream: int, timeout_ms: int = 20000) -> None:
        _cabi.check(self.lib.dn_hop_wait(self.flag(i), seq, timeout_ms, self.err_flag, stream))

    def set_local(self, i: int, seq: int, stream: int) -> None:
        """publish a sequence number on one of OUR OWN flags (e.g. prefilled inputs)."""
        _cabi.check(self.lib.dn_hop_send(self.slot(i), self.slot(i), 0, self.flag(i), seq, stream))

    def timed_out(self) -> bool:
        return bool(int(device_view(self.err_flag, (1,), torch.int32).item()))

    def free(self) -> None:
        self.lib.dn_hop_free(self.data_ptr)
        self.lib.dn_hop_free(self.flag_ptr)


class HopSender:
    """The sending side of one ring link: the peer's slots mapped into this process."""

    def __init__(self, ep: HopEndpoint, same_process_ptrs: Optional[Tuple[int, int]] = None):
        self.lib = _cabi.load()
        self.ep = ep
        self._imported = same_process_ptrs is None
        if same_process_ptrs is not None:          # self-loop / single process: no IPC needed
            self.data_ptr, self.flag_ptr = same_process_ptrs
        else:
            d, f = C.c_void_p(), C.c_void_p()
            hd = (C.c_uint8 * 64).from_buffer_copy(ep.data_handle)
            hf = (C.c_uint8 * 64).from_buffer_copy(ep.flag_handle)
            _cabi.check(self.lib.dn_hop_import(hd, C.byref(d)))
            _cabi.check(self.lib.dn_hop_import(hf, C.byref(f)))
            self.data_ptr, self.flag_ptr = d.value, f.value

    def send(self, i: int, src_ptr: int, nbytes: int, seq: int, stream: int) -> None:
        _cabi.check(self.lib.dn_hop_send(self.data_ptr + i * self.ep.slot_bytes, src_ptr, nbytes,
                                         self.flag_ptr + i * 64, seq, stream))

    def wait_consumed(self, i: int, seq: int, stream: int, timeout_ms: int = 20000) -> None:
        """make ``stream`` wait until the peer has copied out everything up to ``seq`` from slot i"""
        if seq <= 0:
            return
        _cabi.check(self.lib.dn_hop_wait(self.flag_ptr + (self.ep.n_slots + i) * 64, seq, timeout_ms, None, stream))

    def close(self) -> None:
        if self._imported:
            self.lib.dn_hop_close(self.data_ptr)
            self.lib.dn_hop_close(self.flag_ptr)


_PARKED_RECEIVERS: list = []


class HopLink:
    """One shard's device-hop state: its own receive lanes and the ring successor's lanes.

    Per lane (= per in-flight nonce) the receiver owns a *decode* slot (one activation row, H bf16;
    on the head shard its first 4 bytes double as the token slot) and a *bulk* slot (a prefill chunk of
    up to ``bulk_tokens`` rows), each with its own 32-bit sequence flag.  ``endpoint()`` is what the
    predecessor needs to map them (four CUDA IPC handles, picklable / hex-serialisable);
    ``connect(endpoint)`` maps the successor's lanes.  In a single-shard ring the shard is its own
    successor and no IPC is involved.
    """

    def __init__(self, n_lanes: int, hidden: int, bulk_tokens: int = 512, shard_id=None, first_layer: int = -1,
                 n_layers: int = 0):
        self.n_lanes, self.hidden, self.bulk_tokens = int(n_lanes), int(hidden), int(bulk_tokens)
        self.shard_id, self.first_layer, self.n_local_layers = shard_id, int(first_layer), int(n_layers)
        self.rx = HopReceiver(self.n_lanes, hidden * 2)
        self.rx_bulk = HopReceiver(self.n_lanes, self.bulk_tokens * hidden * 2)
        self.mesh = HeadMesh(self.n_lanes, hidden)     # tensor-parallel lm_head area (used when the ring enables it)
        self.tx: Optional[HopSender] = None
        self.tx_bulk: Optional[HopSender] = None

    def endpoint(self) -> dict:
        a, b = self.rx.endpoint(), self.rx_bulk.endpoint()
        return {"n_lanes": self.n_lanes, "hidden": self.hidden, "bulk_tokens": self.bulk_tokens,
                "shard_id": str(self.shard_id), "first_layer": self.first_layer, "n_layers": self.n_local_layers,
                "decode": [a.data_handle.hex(), a.flag_handle.hex()], "bulk": [b.data_handle.hex(), b.flag_handle.hex()],
                "head": self.mesh.endpoint()}

    def connect(self, ep: Optional[dict]) -> None:
        """ep=None: self-loop (single shard, or a successor living in this process passes its HopLink)."""
        if ep is None:
            self.connect_local(self)
            return
        if ep["n_lanes"] != self.n_lanes or ep["hidden"] != self.hidden:
            raise ValueError(f"hop endpoint mismatch: peer lanes/hidden {ep['n_lanes']}/{ep['hidden']} vs "
                             f"{self.n_lanes}/{self.hidden}")
        d = HopEndpoint(bytes.fromhex(ep["decode"][0]), bytes.fromhex(ep["decode"][1]), self.n_lanes, self.hidden * 2)
        b = HopEndpoint(bytes.fromhex(ep["bulk"][0]), bytes.fromhex(ep["bulk"][1]), self.n_lanes,
                        int(ep["bulk_tokens"]) * self.hidden * 2)
        self.tx, self.tx_bulk = HopSender(d), HopSender(b)

    def connect_local(self, peer: "HopLink") -> None:
        d = HopEndpoint(b"", b"", peer.n_lanes, peer.hidden * 2)
        b = HopEndpoint(b"", b"", peer.n_lanes, peer.bulk_tokens * peer.hidden * 2)
        self.tx = HopSender(d, (peer.rx.data_ptr, peer.rx.flag_ptr))
        self.tx_bulk = HopSender(b, (peer.rx_bulk.data_ptr, peer.rx_bulk.flag_ptr))

    @property
    def connected(self) -> bool:
        return self.tx is not None

    def tx_slot(self, lane: int) -> int:
        return self.tx.data_ptr + lane * self.tx.ep.slot_bytes

    def tx_flag(self, lane: int) -> int:
        return self.tx.flag_ptr + lane * 64

    def close(self) -> None:
        """Unmap the successor's lanes.  Our own receive buffers were exported with CUDA IPC and the
        predecessor may still have them mapped (shards of a ring stop in no particular order), so they are
        parked instead of freed -- a few MB that the process returns at exit."""
        for t in (self.tx, self.tx_bulk):
            if t is not None:
                t.close()
        self.tx = self.tx_bulk = None
        self.mesh.close()
        _PARKED_RECEIVERS.extend((self.rx, self.rx_bulk))


class HeadMesh:
    """Tensor-parallel lm_head over the ring: this shard's receive area + every peer's area mapped.

    Per lane the area holds the final hidden state of the lane's token (written by the LAST shard's
    broadcast, one arrival flag) and -- used on the HEAD shard only -- a table of up to 16 partial
    results (max, sum-exp, argmax, pad: 16 bytes each; one flag per shard).  Flag values are the
    lane's decode sequence numbers, so they only grow."""

    MAX_SHARDS = 16

    def __init__(self, n_lanes: int, hidden: int):
        self.lib = _cabi.load()
        self.n_lanes, self.hidden = int(n_lanes), int(hidden)
        self.slot_bytes = hidden * 2 + self.MAX_SHARDS * 16
        d, f = C.c_void_p(), C.c_void_p()
        _cabi.check(self.lib.dn_hop_alloc(self.n_lanes * self.slot_bytes, C.byref(d)))
        _cabi.check(self.lib.dn_hop_alloc(self.n_lanes * (1 + self.MAX_SHARDS) * 64, C.byref(f)))
        self.data_ptr, self.flag_ptr = d.value, f.value
        self.peers: List[Tuple[int, int]] = []        # ring position -> (data_ptr, flag_ptr) in THIS process
        self._imported: List[int] = []

    # -- addresses inside an area (ours or a peer's) ---------------------------------------------
    def x_slot(self, lane: int, base: Optional[int] = None) -> int:
        return (self.data_ptr if base is None else base) + lane * self.slot_bytes

    def partial(self, lane: int, shard: int, base: Optional[int] = None) -> int:
        return self.x_slot(lane, base) + self.hidden * 2 + shard * 16

    def x_flag(self, lane: int, fbase: Optional[int] = None) -> int:
        return (self.flag_ptr if fbase is None else fbase) + lane * (1 + self.MAX_SHARDS) * 64

    def p_flag(self, lane: int, shard: int, fbase: Optional[int] = None) -> int:
        return self.x_flag(lane, fbase) + (1 + shard) * 64

    # -- exchange ----------------------------------------------------------------------------------
    def endpoint(self) -> list:
        hd, hf = (C.c_uint8 * 64)(), (C.c_uint8 * 64)()
        _cabi.check(self.lib.dn_hop_export(self.data_ptr, hd))
        _cabi.check(self.lib.dn_hop_export(self.flag_ptr, hf))
        return [bytes(hd).hex(), bytes(hf).hex()]

    def connect(self, ring: list, own_position: int) -> None:
        """``ring``: per ring position either this process's HeadMesh (same process), ``None`` at
        ``own_position``, or the peer's endpoint [data_handle_hex, flag_handle_hex]."""
        self.peers = []
        for pos, ep in enumerate(ring):
            if pos == own_position or ep is None:
                self.peers.append((self.data_ptr, self.flag_ptr))
            elif isinstance(ep, HeadMesh):
                self.peers.append((ep.data_ptr, ep.flag_ptr))
            else:
                d, f = C.c_void_p(), C.c_void_p()
                hd = (C.c_uint8 * 64).from_buffer_copy(bytes.fromhex(ep[0]))
                hf = (C.c_uint8 * 64).from_buffer_copy(bytes.fromhex(ep[1]))
                _cabi.check(self.lib.dn_hop_import(hd, C.byref(d)))
                _cabi.check(self.lib.dn_hop_import(hf, C.byref(f)))
                self.peers.append((d.value, f.value))
                self._imported += [d.value, f.value]

    def close(self) -> None:
        for ptr in self._imported:
            self.lib.dn_hop_close(ptr)
        self._imported = []
        self.peers = []
        _PARKED_RECEIVERS.append(self)      # exported memory: parked, not freed (see HopLink.close)


def even_split(num_layers: int, world: int) -> List[List[int]]:
    """contiguous equal splits, k=1 (the manual topology the benchmarks use; HALDA does not
    balance identical devices)."""
    base, rem = divmod(num_layers, world)
    out, s = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append(list(range(s, s + n)))
        s += n
    return out


def balanced_split(num_layers: int, world: int, layer_bytes: int, first_extra: int = 0, last_extra: int = 0) -> List[List[int]]:
    """contiguous split, k=1, that minimises the bytes of the busiest shard: with several sequences in
    flight the ring's throughput is set by the shard that streams the most weight bytes per token, and
    the shard with the last layer also streams the lm_head (`last_extra`; Llama-3-8B: 1.05 GB = 2.4
    layers).  Any contiguous `LayerAssignment` is executed unchanged; this is the assignment an operator
    would post to /v1/prepare_topology_manual instead of equal counts.  Ties: smallest sum of squares."""
    if world <= 0 or num_layers < world:
        raise ValueError("need at least one layer per shard")
    INF = (float("inf"), float("inf"))
    # best[r][l] = (max bytes, sum of squares) for the first r shards covering layers [0, l)
    best = [[INF] * (num_layers + 1) for _ in range(world + 1)]
    cut = [[0] * (num_layers + 1) for _ in range(world + 1)]
    best[0][0] = (0, 0)
    for r in range(1, world + 1):
        for l in range(r, num_layers - (world - r) + 1):
            for k in range(r - 1, l):
                prev = best[r - 1][k]
                if prev == INF:
                    continue
                load = (l - k) * layer_bytes + (first_extra if r == 1 else 0) + (last_extra if r == world else 0)
                cand = (max(prev[0], load), prev[1] + load * load)
                if cand < best[r][l]:
                    best[r][l] = cand
                    cut[r][l] = k
    out, l = [], num_layers
    for r in range(world, 0, -1):
        k = cut[r][l]
        out.append(list(range(k, l)))
        l = k
    return out[::-1]
