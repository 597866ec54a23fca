"""Activation buffer pools (row a2; contract of the reference's
src/dnet/core/memory/memory_pool.py:17-394 as pinned by tests/test_memory_pool.py and
tests/test_layer_aware_memory_pool.py:19-119).

Contract: ``allocate(size_bytes, dtype) -> id | None`` reuses a FREE buffer of exactly the same byte
size (and dtype) before creating a new one; when the byte budget would be exceeded, FREE buffers are
dropped least-recently-used first and the allocation fails (None) if that is not enough;
``get_buffer(id)`` returns the flat buffer unless it is FREE; ``get_buffer_view(id, shape)`` a reshaped
prefix; ``release(id)`` drops one reference and frees at zero; ``get_stats()``.  ``buffers``,
``buffer_info`` (``BufferInfo`` records) and ``size_to_buffers`` stay inspectable, as the reference's
tests read them.  ``LayerAwareMemoryPool`` adds per-layer size statistics (median = "typical" size).

The buffers are torch tensors on a chosen *placement*: ``"pinned"`` host memory for the ingress pool
(wire bytes land there and leave with one cudaMemcpyAsync), ``"cuda"`` for device staging, ``"cpu"``
for the GPU-less plumbing tests.
"""
from __future__ import annotations

import math
import threading
import time
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

from dnet_b200.core.types.messages import PoolStatus
from dnet_b200.utils.logger import logger


@dataclass
class BufferInfo:
    """Metadata for a memory pool buffer."""

    buffer_id: int
    size: int
    status: PoolStatus
    last_used: float
    ref_count: int = 0


_ITEMSIZE: Dict[torch.dtype, int] = {}


def _itemsize(dtype: torch.dtype) -> int:
    n = _ITEMSIZE.get(dtype)
    if n is None:
        n = _ITEMSIZE[dtype] = torch.empty((), dtype=dtype).element_size()
    return n


class DynamicMemoryPool:
    """Byte-budgeted pool of flat buffers with exact-size reuse."""

    def __init__(self, total_memory_mb: int = 512, min_buffer_size: int = 1024, placement: str = "cpu") -> None:
        self.total_memory_bytes = total_memory_mb * 1024 * 1024
        self.min_buffer_size = min_buffer_size
        self.placement = placement
        self.used_memory = 0
        self.next_buffer_id = 0
        self.buffers: Dict[int, torch.Tensor] = {}
        self.buffer_info: Dict[int, BufferInfo] = {}
        self.size_to_buffers: Dict[int, List[int]] = {}
        self.lock = threading.Lock()

    # -- storage ----------------------------------------------------------------------------
    def _new(self, n: int, dtype: torch.dtype) -> torch.Tensor:
        if self.placement == "cuda":
            return torch.zeros(n, dtype=dtype, device="cuda")
        t = torch.zeros(n, dtype=dtype)
        return t.pin_memory() if (self.placement == "pinned" and torch.cuda.is_available()) else t

    def _forget(self, buffer_id: int) -> int:
        """lock held: drop a buffer entirely; returns the bytes given back to the budget."""
        info = self.buffer_info.pop(buffer_id, None)
        self.buffers.pop(buffer_id, None)
        if info is None:
            return 0
        same = self.size_to_buffers.get(info.size)
        if same is not None:
            if buffer_id in same:
                same.remove(buffer_id)
            if not same:
                del self.size_to_buffers[info.size]
        self.used_memory -= info.size
        return info.size

    # -- allocation -------------------------------------------------------------------------
    def allocate(self, size_bytes: int, dtype: torch.dtype) -> Optional[int]:
        item = _itemsize(dtype)
        nbytes = -(-size_bytes // item) * item               # whole elements
        with self.lock:
            reuse = self._find_free_buffer(nbytes, dtype)
            if reuse is not None:
                info = self.buffer_info[reuse]
                info.status, info.ref_count = PoolStatus.ALLOCATED, 1
                return reuse
            over = self.used_memory + nbytes - self.total_memory_bytes
            if over > 0 and not self._evict_unused_buffers(nbytes):
                logger.warning(f"Cannot allocate {nbytes} bytes - insufficient memory")
                return None
            try:
                storage = self._new(nbytes // item, dtype)
            except (MemoryError, RuntimeError):
                logger.error(f"Failed to allocate {nbytes} bytes - out of memory")
                return None
            bid = self.next_buffer_id
            self.next_buffer_id += 1
            self.buffers[bid] = storage
            self.buffer_info[bid] = BufferInfo(buffer_id=bid, size=nbytes, status=PoolStatus.ALLOCATED,
                                               last_used=time.time(), ref_count=1)
            self.size_to_buffers.setdefault(nbytes, []).append(bid)
            self.used_memory += nbytes
            return bid

    def _find_free_buffer(self, size: int, dtype: Optional[torch.dtype] = None) -> Optional[int]:
        for bid in self.size_to_buffers.get(size, ()):
            info = self.buffer_info.get(bid)
            if info is not None and info.status == PoolStatus.FREE and (dtype is None or self.buffers[bid].dtype == dtype):
                return bid
        return None

    def _evict_unused_buffers(self, needed_bytes: int) -> bool:
        """Drop FREE buffers, least recently used first, until ``needed_bytes`` were given back."""
        idle = sorted((info.last_used, bid) for bid, info in self.buffer_info.items() if info.status == PoolStatus.FREE)
        freed = 0
        for _, bid in idle:
            if freed >= needed_bytes:
                break
            freed += self._forget(bid)
        return freed >= needed_bytes

    # -- access -----------------------------------------------------------------------------
    def get_buffer(self, buffer_id: int) -> Optional[torch.Tensor]:
        with self.lock:
            info = self.buffer_info.get(buffer_id)
            if info is None or info.status == PoolStatus.FREE or buffer_id not in self.buffers:
                return None
            info.last_used = time.time()
            return self.buffers[buffer_id]

    def get_buffer_view(self, buffer_id: int, shape: Tuple[int, ...]) -> Optional[torch.Tensor]:
        flat = self.get_buffer(buffer_id)
        if flat is None:
            return None
        want = math.prod(shape)
        if want > flat.numel():
            logger.error(f"Buffer {buffer_id} too small for shape {shape}")
            return None
        try:
            return flat[:want].reshape(shape)
        except (ValueError, RuntimeError) as e:
            logger.error(f"Cannot reshape buffer {buffer_id} to {shape}: {e}")
            return None

    def release(self, buffer_id: int) -> None:
        with self.lock:
            info = self.buffer_info.get(buffer_id)
            if info is None:
                return
            info.ref_count = max(0, info.ref_count - 1)
            if info.ref_count == 0:
                info.status = PoolStatus.FREE

    def get_stats(self) -> Dict:
        mb = 1024 * 1024
        with self.lock:
            states = [i.status for i in self.buffer_info.values()]
            return {
                "total_memory_mb": self.total_memory_bytes // mb,
                "used_memory_mb": self.used_memory // mb,
                "free_memory_mb": (self.total_memory_bytes - self.used_memory) // mb,
                "total_buffers": len(states),
                "free_buffers": states.count(PoolStatus.FREE),
                "allocated_buffers": states.count(PoolStatus.ALLOCATED),
                "buffer_sizes": list(self.size_to_buffers),
            }


class LayerAwareMemoryPool:
    """DynamicMemoryPool + what each layer typically asks for (the reference keeps the last 50-100 request
    sizes per layer and reports their median)."""

    _KEEP, _TRIM_AT = 50, 100

    def __init__(self, total_memory_mb: int = 512, placement: str = "cpu") -> None:
        self.pool = DynamicMemoryPool(total_memory_mb, placement=placement)
        self.layer_stats: Dict[int, Dict] = {}
        self.lock = threading.Lock()

    def allocate_for_layer(self, layer_id: int, shape: Tuple[int, ...], dtype: torch.dtype) -> Optional[int]:
        nbytes = math.prod(shape) * _itemsize(dtype)
        with self.lock:
            st = self.layer_stats.setdefault(layer_id, {"sizes": [], "shapes": [], "allocations": 0})
            st["sizes"].append(nbytes)
            st["shapes"].append(shape)
            st["allocations"] += 1
            if len(st["sizes"]) > self._TRIM_AT:
                st["sizes"], st["shapes"] = st["sizes"][-self._KEEP:], st["shapes"][-self._KEEP:]
        return self.pool.allocate(nbytes, dtype)

    def get_layer_buffer(self, buffer_id: int, shape: Tuple[int, ...]) -> Optional[torch.Tensor]:
        return self.pool.get_buffer_view(buffer_id, shape)

    def get_typical_size(self, layer_id: int) -> Optional[int]:
        with self.lock:
            sizes = self.layer_stats.get(layer_id, {}).get("sizes")
            return sorted(sizes)[len(sizes) // 2] if sizes else None

    def get_buffer(self, buffer_id: int) -> Optional[torch.Tensor]:
        return self.pool.get_buffer(buffer_id)

    def release(self, buffer_id: int) -> None:
        self.pool.release(buffer_id)

    def get_stats(self) -> Dict:
        with self.lock:
            per_layer = {lid: {"allocations": st["allocations"],
                               "avg_size_mb": sum(st["sizes"]) / len(st["sizes"]) / (1024 * 1024),
                               "recent_shapes": list(set(st["shapes"][-10:]))}
                         for lid, st in self.layer_stats.items() if st["sizes"]}
        return {"pool": self.pool.get_stats(), "layer_stats": per_layer}
