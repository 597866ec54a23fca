"""Disk/host -> HBM materialisation of one layer
(reference src/dnet/utils/layer_manager.py:37-292, rebuilt for B200).

The reference mmaps safetensors files, madvise()s the byte ranges of the next layer and
copies tensor by tensor into MLX arrays.  Here each assigned layer is packed into a
page-locked host buffer (one contiguous record per layer, tensors at 256-byte aligned
offsets) and a load is a single ``cudaMemcpyAsync`` of that record into an HBM layer slot
on the prefetch stream (``dn_slot_prefetch``), followed by an event the compute stream
waits on -- no host thread ever blocks on the copy.

Where the record comes from:
  * ``use_mxload_fastpath`` (the reference's switch for its per-layer ``mx.load``): the assigned layers are
    repacked once into per-layer safetensors files (utils/repack.py, the reference's on-disk format) whose
    data region IS the record, and a record is one sequential read disk -> pinned memory;
  * otherwise tensor by tensor out of the mmapped checkpoint shards.
How long it stays pinned: ``keep_host_records`` -- offload / sliding_fit keep the records (they are the
backing store of every swap; ``host_record_budget`` > 0 caps how many stay pinned, least recently used
first, the rest is re-read from disk), the fit policy drops a record as soon as its one H2D copy ran.
"""
from __future__ import annotations

import os
import threading
import time
from concurrent.futures import Future, ThreadPoolExecutor
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

from dnet_b200.utils.logger import logger
from dnet_b200.utils.model import MappedFile, ModelMetadata, SyntheticSource, TensorInfo, load_weight
from dnet_b200.utils.serialization import safetensor_torch_dtype

_VALID_PREFETCH_MODES = {"full", "sequential", "off"}
_ALIGN = 256


@dataclass(frozen=True)
class PackedEntry:
    suffix: str
    offset: int
    nbytes: int
    shape: Tuple[int, ...]
    dtype: torch.dtype


class LayerManager:
    """Packs assigned layers into pinned host records and stages them into HBM slots."""

    def __init__(self, model_metadata: ModelMetadata, assigned_layers: List[int], thread_pool_size: int = 2, *,
                 use_mxload_fastpath: bool = False, prefetch_mode: str = "off", stage_host: bool = True,
                 device: Optional[str] = None, keep_host_records: bool = True, host_record_budget: int = 0):
        self.assigned_layers = set(assigned_layers)
        self.weight_info = model_metadata.weight_info
        self.source = model_metadata.source
        self.mapped_files: Dict[str, MappedFile] = {}
        self.executor = ThreadPoolExecutor(max_workers=thread_pool_size)
        self._use_mxload_fastpath = bool(use_mxload_fastpath)
        pm = (prefetch_mode or "off").strip().lower()
        self._prefetch_mode = pm if pm in _VALID_PREFETCH_MODES else "off"
        self._stage_host = stage_host
        self.device = device or ("cuda" if torch.cuda.is_available() else "cpu")
        self._layout: Dict[int, List[PackedEntry]] = {}
        self._layer_bytes: Dict[int, int] = {}
        self._host: Dict[int, torch.Tensor] = {}   # layer -> pinned uint8 record
        self._host_lock = threading.Lock()
        self._prefetch_stream = None
        self._keep_host = bool(keep_host_records)
        self._host_budget = int(host_record_budget or int(os.environ.get("DNET_COMPUTE_HOST_RECORDS", "0") or 0))
        self._host_lru: List[int] = []
        self.repack_dir = None                # per-layer files (repack fast path), set below
        self.record_reads = {"sequential": 0, "per-tensor": 0, "checkpoint": 0}
        if self._use_mxload_fastpath and self.source is None and isinstance(model_metadata.path, os.PathLike):
            try:
                from dnet_b200.utils.repack import ensure_repacked_for_layers

                t0 = time.perf_counter()
                self.repack_dir, did = ensure_repacked_for_layers(str(model_metadata.path), sorted(self.assigned_layers),
                                                                  md=model_metadata)
                logger.info("[REPACK] %s per-layer files in %s (%.1fs)", "wrote" if did else "reusing", self.repack_dir,
                            time.perf_counter() - t0)
            except Exception as e:
                logger.warning("repack fast path unavailable (%s); reading tensors from the checkpoint shards", e)
                self.repack_dir = None
        for lid in sorted(self.assigned_layers):
            self._build_layout(lid)

    # -- layout ---------------------------------------------------------------
    def _build_layout(self, layer_idx: int) -> None:
        off = 0
        ents: List[PackedEntry] = []
        for suffix in sorted(self.weight_info[layer_idx]):
            wt = self.weight_info[layer_idx][suffix]
            ents.append(PackedEntry(suffix, off, wt.size_bytes, tuple(wt.shape), safetensor_torch_dtype[wt.dtype]))
            off = (off + wt.size_bytes + _ALIGN - 1) // _ALIGN * _ALIGN
        self._layout[layer_idx] = ents
        self._layer_bytes[layer_idx] = max(off, _ALIGN)

    def layer_bytes(self, layer_idx: int) -> int:
        return self._layer_bytes[layer_idx]

    def max_layer_bytes(self) -> int:
        return max(self._layer_bytes.values()) if self._layer_bytes else 0

    # -- host staging ------------------------------------------------------------
    def _host_record(self, layer_idx: int) -> torch.Tensor:
        with self._host_lock:
            rec = self._host.get(layer_idx)
            if rec is not None:
                if layer_idx in self._host_lru:
                    self._host_lru.remove(layer_idx)
                self._host_lru.append(layer_idx)
                return rec
        nbytes = self._layer_bytes[layer_idx]
        rec = torch.empty(nbytes, dtype=torch.uint8)
        if torch.cuda.is_available():
            rec = rec.pin_memory()
        how = None
        if self.repack_dir is not None:
            from dnet_b200.utils.repack import layer_file_name, read_layer_record

            f = self.repack_dir / layer_file_name(layer_idx)
            if f.exists():
                try:
                    how = read_layer_record(f, self._layout[layer_idx], rec)
                except Exception as e:
                    logger.warning("could not read %s (%s); falling back to the checkpoint shards", f, e)
        if how is None:
            how = "checkpoint"
            for e in self._layout[layer_idx]:
                wt = self.weight_info[layer_idx][e.suffix]
                src = load_weight(wt, self.mapped_files, self.source)
                rec[e.offset:e.offset + e.nbytes].copy_(src.contiguous().view(torch.uint8).reshape(-1))
        self.record_reads[how] += 1
        with self._host_lock:
            self._host[layer_idx] = rec
            self._host_lru.append(layer_idx)
            if self._host_budget > 0:
                while len(self._host_lru) > self._host_budget:      # LRU: re-read from disk when needed again
                    self._host.pop(self._host_lru.pop(0), None)
        return rec

    def drop_host_record(self, layer_idx: int) -> None:
        with self._host_lock:
            self._host.pop(layer_idx, None)
            if layer_idx in self._host_lru:
                self._host_lru.remove(layer_idx)

    def stage_all_to_host(self) -> int:
        """safetensors -> pinned host, once (offload mode calls this at load time)."""
        total = 0
        for lid in sorted(self.assigned_layers):
            total += self._host_record(lid).numel()
        return total

    # -- madvise analogue ----------------------------------------------------------
    def prefetch_layer(self, layer_idx: int) -> bool:
        """Warm the host side for a layer (reference prefetch_layer: madvise WILLNEED).
        With pinned staging this packs the record if it is not packed yet."""
        if layer_idx not in self.assigned_layers:
            return False
        if self._prefetch_mode == "off":
            return True
        t0 = time.perf_counter()
        self._host_record(layer_idx)
        logger.debug("[PROFILE][PREFETCH] layer=%s ms=%.2f", layer_idx, (time.perf_counter() - t0) * 1e3)
        return True

    def async_prefetch(self, layer_idx: int) -> Future:
        return self.executor.submit(self.prefetch_layer, layer_idx)

    def release_layer(self, layer_idx: int) -> bool:
        """reference release_layer: madvise DONTNEED.  Pinned records stay (they ARE the
        backing store in offload mode); file pages are dropped via posix_fadvise."""
        if layer_idx not in self.assigned_layers:
            return False
        if self._prefetch_mode == "off":
            return True
        for wt in self.weight_info[layer_idx].values():
            mf = self.mapped_files.get(wt.filename)
            if mf is not None and hasattr(os, "posix_fadvise"):
                try:
                    os.posix_fadvise(mf.file.fileno(), wt.offset, wt.size_bytes, os.POSIX_FADV_DONTNEED)
                except OSError:
                    pass
        return True

    # -- device materialisation -------------------------------------------------
    def views(self, layer_idx: int, slot: torch.Tensor) -> Dict[str, torch.Tensor]:
        out: Dict[str, torch.Tensor] = {}
        for e in self._layout[layer_idx]:
            out[f"layers.{layer_idx}.{e.suffix}"] = slot[e.offset:e.offset + e.nbytes].view(e.dtype).view(e.shape)
        return out

    def load_layer_to_gpu(self, layer_idx: int, slot: Optional[torch.Tensor] = None,
                          stream: Optional[int] = None, wait_event=None) -> Dict[str, torch.Tensor]:
        """Materialise one layer into an HBM slot; returns ``layers.<abs>.<suffix>`` -> device
        tensor views (reference layer_manager.py:229-282) plus ``"_ready_event"`` (a dn_event
        recorded on the prefetch stream; the compute stream must wait on it)."""
        if layer_idx not in self.assigned_layers:
            raise RuntimeError(f"layer {layer_idx} not assigned to this node")
        from dnet_b200 import _cabi
        import ctypes as C

        lib = _cabi.load()
        nbytes = self._layer_bytes[layer_idx]
        if slot is None:
            slot = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        if isinstance(self.source, SyntheticSource) and not self._stage_host:
            data = self.views(layer_idx, slot)
            for e in self._layout[layer_idx]:
                self.source.fill_device(self.weight_info[layer_idx][e.suffix], data[f"layers.{layer_idx}.{e.suffix}"])
            data["_slot"] = slot
            data["_ready_event"] = None
            torch.cuda.current_stream().synchronize()   # generated on torch's stream, consumed on compute_stream
            return data
        rec = self._host_record(layer_idx)
        if stream is None:
            if self._prefetch_stream is None:
                s = C.c_void_p()
                _cabi.check(lib.dn_stream_create(C.byref(s), 0))
                self._prefetch_stream = s.value
            stream = self._prefetch_stream
        if wait_event is not None:
            _cabi.check(lib.dn_stream_wait_event(stream, wait_event))
        ev = C.c_void_p()
        _cabi.check(lib.dn_event_create(C.byref(ev), 0))
        _cabi.check(lib.dn_slot_prefetch(slot.data_ptr(), rec.data_ptr(), nbytes, stream, ev.value))
        data = self.views(layer_idx, slot)
        data["_slot"] = slot
        data["_ready_event"] = ev.value
        if not self._keep_host:
            # fit mode: the layer stays resident in HBM and is never reloaded, so its pinned staging record
            # (and, for quantised checkpoints, the dequantised copy in it) is released once the copy has run
            _cabi.check(lib.dn_event_sync(ev.value))
            self.drop_host_record(layer_idx)
        return data

    def close(self) -> None:
        self.executor.shutdown(wait=False, cancel_futures=True)
        for mf in self.mapped_files.values():
            try:
                mf.close()
            except Exception:
                pass
        self.mapped_files.clear()
