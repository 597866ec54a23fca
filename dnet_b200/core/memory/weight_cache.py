"""Layer weight cache with windowed HBM residency and LRU eviction
(reference src/dnet/core/memory/weight_cache.py:15-300).

Same contract: budget = min(#assigned, resident_windows * window_size); cache hit ->
LRU touch; miss -> evict the least-recent zero-reference layer if at budget, install an
in-flight Future, load outside the lock, resolve the future; refcount-gated eviction.
What changed underneath: a "load" enqueues one cudaMemcpyAsync (pinned host -> HBM slot)
on the prefetch stream and returns immediately with an event, evicted layers hand their
HBM slot back to a free list instead of being garbage-collected, and a slot is only
overwritten after the compute stream's last read of it (release events).
"""
from __future__ import annotations

import threading
import time
from concurrent.futures import Future
from typing import Any, Dict, List, Optional

from dnet_b200.utils.logger import logger
from dnet_b200.utils.model import ModelMetadata


class WeightCache:
    def __init__(self, assigned_layers: List[int], model_metadata: Optional[ModelMetadata],
                 window_size: Optional[int] = None, prefetch_threads: int = 2, *, resident_windows: int = 2,
                 use_mxload_fastpath: bool = False, prefetch_mode: str = "off", layer_manager=None,
                 stage_host: bool = True):
        self.assigned_layers = assigned_layers
        resident_windows = max(1, int(resident_windows))
        if window_size is not None and window_size > 0:
            self.max_weights = min(len(self.assigned_layers), max(1, resident_windows * int(window_size)))
        else:
            self.max_weights = len(self.assigned_layers)
        self.cache: Dict[int, tuple[Dict[str, Any], float]] = {}
        self.reference_counts: Dict[int, int] = {}
        if layer_manager is None:
            from dnet_b200.utils.layer_manager import LayerManager

            layer_manager = LayerManager(model_metadata, assigned_layers, thread_pool_size=int(prefetch_threads or 2),
                                         use_mxload_fastpath=bool(use_mxload_fastpath), prefetch_mode=prefetch_mode,
                                         stage_host=stage_host)
        self.layer_manager = layer_manager
        self.lock = threading.Lock()
        self.loading_futures: Dict[int, Future] = {}
        self.prefetch_futures: Dict[int, Future] = {}
        self.closed = False
        # HBM slot recycling
        self._free_slots: List[Any] = []           # (slot tensor, last_use_event or None)
        self._release_events: Dict[int, Any] = {}  # layer -> event recorded after its last compute use
        logger.info("WeightCache resident budget: max_weights=%d", self.max_weights)

    def shutdown(self):
        with self.lock:
            self.closed = True
        self.cancel_all_prefetch()

    # -- slot plumbing ---------------------------------------------------------------
    def _take_slot(self):
        if self._free_slots:
            return self._free_slots.pop()
        return None, None

    def _return_slot(self, layer_id: int, data: Dict[str, Any]) -> None:
        slot = data.get("_slot") if isinstance(data, dict) else None
        if slot is not None:
            self._free_slots.append((slot, self._release_events.pop(layer_id, None)))
            ev = data.get("_ready_event")
            if ev is not None:  # waits already enqueued on it stay valid after destroy
                self._destroy_event(ev)
                data["_ready_event"] = None

    @staticmethod
    def _destroy_event(ev) -> None:
        try:
            from dnet_b200 import _cabi

            _cabi.load().dn_event_destroy(ev)
        except Exception:
            pass

    def _load(self, layer_id: int):
        lm = self.layer_manager
        try:
            with self.lock:
                slot, ev = self._take_slot()
            data = lm.load_layer_to_gpu(layer_id, slot=slot, wait_event=ev)
            if ev is not None:
                self._destroy_event(ev)
            return data
        except TypeError:  # injected fake layer managers take only the layer id
            return lm.load_layer_to_gpu(layer_id)

    # -- reference API -----------------------------------------------------------------
    def get_weight(self, layer_id: int, *, inc_ref: bool = True) -> Optional[Dict[str, Any]]:
        if self.closed:
            return None
        with self.lock:
            if self.closed:
                return None
            if layer_id in self.cache:
                data, _ = self.cache[layer_id]
                self.cache[layer_id] = (data, time.time())
                if inc_ref:
                    self.reference_counts[layer_id] = self.reference_counts.get(layer_id, 0) + 1
                return data
            inflight = self.loading_futures.get(layer_id)
            if inflight is None:
                if len(self.cache) >= self.max_weights:
                    self._evict_lru()
                fut: Future = Future()
                self.loading_futures[layer_id] = fut
                inflight = fut
                creator = True
            else:
                creator = False

        if creator:
            try:
                t0 = time.perf_counter()
                data = self._load(layer_id)
                dt_ms = (time.perf_counter() - t0) * 1000.0
                try:
                    winfo = self.layer_manager.weight_info.get(layer_id, {})
                    total_bytes = sum(w.size_bytes for w in winfo.values())
                except Exception:
                    total_bytes = 0
                with self.lock:
                    if self.closed:
                        return None
                    self.cache[layer_id] = (data, time.time())
                    if inc_ref:
                        self.reference_counts[layer_id] = self.reference_counts.get(layer_id, 0) + 1
                    else:
                        self.reference_counts.setdefault(layer_id, 0)
                    fut2 = self.loading_futures.pop(layer_id, None)
                    if fut2 is not None and not fut2.done():
                        fut2.set_result(True)
                logger.info("[PROFILE][MATERIALIZE] layer=%s ms=%.2f bytes=%.2fMB", layer_id, dt_ms,
                            total_bytes / 1_048_576)
                return data
            except Exception as e:
                with self.lock:
                    fut2 = self.loading_futures.pop(layer_id, None)
                    if fut2 is not None and not fut2.done():
                        fut2.set_exception(e)
                if isinstance(e, OSError) and self.closed:
                    logger.warning("Ignored load error for layer %s during shutdown: %s", layer_id, e)
                    return None
                logger.exception("Failed to load weight %s: %s", layer_id, e)
                return None
        else:
            try:
                inflight.result()
            except Exception as e:
                logger.error("Wait for layer %s load failed: %s", layer_id, e)
                return None
            with self.lock:
                entry = self.cache.get(layer_id)
                if entry is None:
                    return None
                data, _ = entry
                if data is None:
                    return None
                self.cache[layer_id] = (data, time.time())
                if inc_ref:
                    self.reference_counts[layer_id] = self.reference_counts.get(layer_id, 0) + 1
                else:
                    self.reference_counts.setdefault(layer_id, 0)
                return data

    def decrease_reference(self, layer_id: int, release_event=None):
        """Decrease reference count; ``release_event`` (recorded on the compute stream after
        the layer's last kernel) gates any later overwrite of the layer's HBM slot."""
        with self.lock:
            if layer_id in self.reference_counts:
                self.reference_counts[layer_id] -= 1
            if release_event is not None:
                old = self._release_events.get(layer_id)
                if old is not None:
                    self._destroy_event(old)
                self._release_events[layer_id] = release_event

    def decrease_references(self, layer_ids) -> None:
        """decrease_reference for a whole run under one lock acquisition (decode hot path)."""
        with self.lock:
            rc = self.reference_counts
            for layer_id in layer_ids:
                if layer_id in rc:
                    rc[layer_id] -= 1

    def prefetch_to_ram(self, layer_id: int):
        try:
            if self.layer_manager._prefetch_mode == "off":
                return None
            f = self.prefetch_futures.get(layer_id)
            if f is not None and not f.done():
                return f
            f = self.layer_manager.async_prefetch(layer_id)
            self.prefetch_futures[layer_id] = f
            return f
        except Exception:
            return None

    def cancel_all_prefetch(self):
        with self.lock:
            for _, fut in list(self.prefetch_futures.items()):
                try:
                    if fut is not None and not fut.done():
                        fut.cancel()
                except Exception:
                    pass
            self.prefetch_futures.clear()

    def _evict_lru(self):
        candidates = [(lid, t) for lid, (_, t) in self.cache.items() if self.reference_counts.get(lid, 0) == 0]
        if candidates:
            candidates.sort(key=lambda x: x[1])
            layer_id = candidates[0][0]
            try:
                self.layer_manager.release_layer(layer_id)
            except Exception:
                pass
            data, _ = self.cache.pop(layer_id)
            self.reference_counts.pop(layer_id, None)
            self._return_slot(layer_id, data)
            logger.info("Evicted layer %s from cache", layer_id)

    def evict_layer(self, layer_id: int) -> bool:
        with self.lock:
            if self.reference_counts.get(layer_id, 0) != 0:
                return False
            if layer_id not in self.cache:
                return True
            try:
                self.layer_manager.release_layer(layer_id)
            except Exception:
                pass
            data, _ = self.cache.pop(layer_id)
            self.reference_counts.pop(layer_id, None)
            self._return_slot(layer_id, data)
            return True

    def evict_layers(self, layer_ids: List[int]) -> int:
        count = 0
        for lid in layer_ids:
            try:
                if self.evict_layer(lid):
                    count += 1
            except Exception:
                continue
        return count

    def get_resident_layers(self) -> List[int]:
        with self.lock:
            items = sorted(self.cache.items(), key=lambda kv: kv[1][1])
            return [lid for lid, _ in items]
