"""Activation buffer pools (reference src/dnet/core/memory/memory_pool.py:17-394).

Same API and semantics -- exact-size reuse keyed by byte size, LRU eviction of FREE
buffers, ref counts, per-layer size statistics -- but the buffers are torch tensors on
a chosen placement: "pinned" host memory for the ingress pool (wire bytes land there and
are copied to HBM with cudaMemcpyAsync), "cuda" for device-side staging, "cpu" for the
GPU-less plumbing tests.
"""
from __future__ import annotations

import threading
import time
from dataclasses import dataclass
from functools import reduce
from operator import mul
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


def _itemsize(dtype: torch.dtype) -> int:
    return torch.empty((), dtype=dtype).element_size()


class DynamicMemoryPool:
    """Dynamic memory pool that allocates buffers of varying sizes."""

    def __init__(self, total_memory_mb: int = 512, min_buffer_size: int = 1024, placement: str = "cpu") -> None:
        self.total_memory_bytes = total_memory_mb * 1024 * 1024
        self.min_buffer_size = min_buffer_size
        self.used_memory = 0
        self.next_buffer_id = 0
        self.placement = placement
        self.buffers: Dict[int, torch.Tensor] = {}
        self.buffer_info: Dict[int, BufferInfo] = {}
        self.size_to_buffers: Dict[int, List[int]] = {}
        self.lock = threading.Lock()

    def _new(self, n: int, dtype: torch.dtype) -> torch.Tensor:
        if self.placement == "cuda":
            return torch.zeros(n, dtype=dtype, device="cuda")
        if self.placement == "pinned" and torch.cuda.is_available():
            return torch.zeros(n, dtype=dtype).pin_memory()
        return torch.zeros(n, dtype=dtype)

    def allocate(self, size_bytes: int, dtype: torch.dtype) -> Optional[int]:
        with self.lock:
            dsize = _itemsize(dtype)
            aligned_size = ((size_bytes + (dsize - 1)) // dsize) * dsize
            buffer_id = self._find_free_buffer(aligned_size, dtype)
            if buffer_id is not None:
                self.buffer_info[buffer_id].status = PoolStatus.ALLOCATED
                self.buffer_info[buffer_id].ref_count = 1
                return buffer_id
            if self.used_memory + aligned_size > self.total_memory_bytes:
                if not self._evict_unused_buffers(aligned_size):
                    logger.warning(f"Cannot allocate {aligned_size} bytes - insufficient memory")
                    return None
            try:
                buffer = self._new(aligned_size // dsize, dtype)
            except (MemoryError, RuntimeError):
                logger.error(f"Failed to allocate {aligned_size} bytes - out of memory")
                return None
            buffer_id = self.next_buffer_id
            self.next_buffer_id += 1
            self.buffers[buffer_id] = buffer
            self.buffer_info[buffer_id] = BufferInfo(
                buffer_id=buffer_id, size=aligned_size, status=PoolStatus.ALLOCATED,
                last_used=time.time(), ref_count=1)
            self.size_to_buffers.setdefault(aligned_size, []).append(buffer_id)
            self.used_memory += aligned_size
            return buffer_id

    def get_buffer(self, buffer_id: int) -> Optional[torch.Tensor]:
        with self.lock:
            if buffer_id in self.buffers and buffer_id in self.buffer_info:
                info = self.buffer_info[buffer_id]
                if info.status != PoolStatus.FREE:
                    info.last_used = time.time()
                    return self.buffers[buffer_id]
        return None

    def get_buffer_view(self, buffer_id: int, shape: Tuple[int, ...]) -> Optional[torch.Tensor]:
        buffer = self.get_buffer(buffer_id)
        if buffer is not None:
            try:
                required_size = reduce(mul, shape)
                if required_size <= len(buffer):
                    return buffer[:required_size].reshape(shape)
                logger.error(f"Buffer {buffer_id} too small for shape {shape}")
            except (ValueError, RuntimeError) as e:
                logger.error(f"Cannot reshape buffer {buffer_id} to {shape}: {e}")
        return None

    def release(self, buffer_id: int) -> None:
        with self.lock:
            if buffer_id in self.buffer_info:
                info = self.buffer_info[buffer_id]
                info.ref_count -= 1
                if info.ref_count <= 0:
                    info.status = PoolStatus.FREE
                    info.ref_count = 0

    def _find_free_buffer(self, size: int, dtype: Optional[torch.dtype] = None) -> Optional[int]:
        if size in self.size_to_buffers:
            for buffer_id in self.size_to_buffers[size]:
                if (buffer_id in self.buffer_info and self.buffer_info[buffer_id].status == PoolStatus.FREE
                        and (dtype is None or self.buffers[buffer_id].dtype == dtype)):
                    return buffer_id
        return None

    def _evict_unused_buffers(self, needed_bytes: int) -> bool:
        free_buffers = [(info.last_used, bid, info.size) for bid, info in self.buffer_info.items()
                        if info.status == PoolStatus.FREE]
        free_buffers.sort()
        freed_bytes = 0
        for _, buffer_id, size in free_buffers:
            if freed_bytes >= needed_bytes:
                break
            self.buffers.pop(buffer_id, None)
            if buffer_id in self.buffer_info:
                bsz = self.buffer_info[buffer_id].size
                if bsz in self.size_to_buffers:
                    self.size_to_buffers[bsz].remove(buffer_id)
                    if not self.size_to_buffers[bsz]:
                        del self.size_to_buffers[bsz]
                del self.buffer_info[buffer_id]
            freed_bytes += size
            self.used_memory -= size
        return freed_bytes >= needed_bytes

    def get_stats(self) -> Dict:
        with self.lock:
            infos = list(self.buffer_info.values())
            return {
                "total_memory_mb": self.total_memory_bytes // (1024 * 1024),
                "used_memory_mb": self.used_memory // (1024 * 1024),
                "free_memory_mb": (self.total_memory_bytes - self.used_memory) // (1024 * 1024),
                "total_buffers": len(infos),
                "free_buffers": sum(1 for i in infos if i.status == PoolStatus.FREE),
                "allocated_buffers": sum(1 for i in infos if i.status == PoolStatus.ALLOCATED),
                "buffer_sizes": list(self.size_to_buffers.keys()),
            }


class LayerAwareMemoryPool:
    """Memory pool that's aware of layer-specific activation patterns."""

    def __init__(self, total_memory_mb: int = 512, placement: str = "cpu") -> None:
        self.pool = DynamicMemoryPool(total_memory_mb, placement=placement)
        self.layer_stats: Dict[int, Dict] = {}
        self.lock = threading.Lock()

    def allocate_for_layer(self, layer_id: int, shape: Tuple[int, ...], dtype: torch.dtype) -> Optional[int]:
        size_bytes = reduce(mul, shape) * _itemsize(dtype)
        with self.lock:
            stats = self.layer_stats.setdefault(layer_id, {"sizes": [], "shapes": [], "allocations": 0})
            stats["sizes"].append(size_bytes)
            stats["shapes"].append(shape)
            stats["allocations"] += 1
            if len(stats["sizes"]) > 100:
                stats["sizes"] = stats["sizes"][-50:]
                stats["shapes"] = stats["shapes"][-50:]
        return self.pool.allocate(size_bytes, dtype)

    def get_layer_buffer(self, buffer_id: int, shape: Tuple[int, ...]) -> Optional[torch.Tensor]:
        return self.pool.get_buffer_view(buffer_id, shape)

    def get_typical_size(self, layer_id: int) -> Optional[int]:
        with self.lock:
            if layer_id in self.layer_stats and self.layer_stats[layer_id]["sizes"]:
                sizes = self.layer_stats[layer_id]["sizes"]
                return sorted(sizes)[len(sizes) // 2]
        return None

    def release(self, buffer_id: int) -> None:
        self.pool.release(buffer_id)

    def get_buffer(self, buffer_id: int) -> Optional[torch.Tensor]:
        return self.pool.get_buffer(buffer_id)

    def get_stats(self) -> Dict:
        pool_stats = self.pool.get_stats()
        with self.lock:
            layer_stats = {}
            for layer_id, stats in self.layer_stats.items():
                if stats["sizes"]:
                    layer_stats[layer_id] = {
                        "allocations": stats["allocations"],
                        "avg_size_mb": sum(stats["sizes"]) / len(stats["sizes"]) / (1024 * 1024),
                        "recent_shapes": list(set(stats["shapes"][-10:])),
                    }
        return {"pool": pool_stats, "layer_stats": layer_stats}
