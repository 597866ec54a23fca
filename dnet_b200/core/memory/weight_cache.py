"""Layer weight cache with windowed HBM residency (row a16; contract of the reference's
src/dnet/core/memory/weight_cache.py:15-300 as pinned by its tests/test_weight_cache.py:23-231).

Contract kept, because policies and the reference's tests probe it directly:
  * ``max_weights = min(#assigned, resident_windows * window_size)`` (all layers without a window);
  * ``cache[layer] = (tensors, last_use_time)``, ``reference_counts[layer]``, ``loading_futures[layer]``
    (one in-flight load per layer; late callers wait on it), ``lock``;
  * ``get_weight`` = hit (touch, +ref) | join an in-flight load | become the loader (evict the least
    recently used *unreferenced* layer first when at budget -- never a referenced one, the cache
    overfills instead -- then load OUTSIDE the lock and publish);
  * ``decrease_reference``, ``evict_layer`` (refused while referenced, True when absent), ``evict_layers``,
    ``get_resident_layers`` (least recent first), ``prefetch_to_ram`` / ``cancel_all_prefetch``, ``shutdown``.

What is specific to this rebuild: a "load" only ENQUEUES one pinned-host -> HBM copy on the prefetch
stream and returns with a ready event; an evicted layer's HBM slot goes back to a free list together
with the event recorded after the compute stream's last read of it, and the next load into that slot
is ordered behind that event (``LayerManager.load_layer_to_gpu(slot=, wait_event=)``).
"""
from __future__ import annotations

import inspect
import threading
import time
from concurrent.futures import Future
from typing import Any, Dict, Iterable, List, Optional, Tuple

from dnet_b200.utils.logger import logger
from dnet_b200.utils.model import ModelMetadata

_HIT, _JOIN, _LOAD = 0, 1, 2


class WeightCache:
    def __init__(self, assigned_layers: List[int], model_metadata: Optional[ModelMetadata],
                 window_size: Optional[int] = None, prefetch_threads: int = 2, *, resident_windows: int = 2,
                 use_mxload_fastpath: bool = False, prefetch_mode: str = "off", layer_manager=None,
                 stage_host: bool = True, keep_host_records: bool = True):
        self.assigned_layers = assigned_layers
        n = len(assigned_layers)
        if window_size is not None and window_size > 0:
            self.max_weights = min(n, max(1, max(1, int(resident_windows)) * int(window_size)))
        else:
            self.max_weights = n
        self.cache: Dict[int, Tuple[Dict[str, Any], float]] = {}
        self.reference_counts: Dict[int, int] = {}
        self.loading_futures: Dict[int, Future] = {}
        self.prefetch_futures: Dict[int, Future] = {}
        self.lock = threading.Lock()
        self.closed = False
        if layer_manager is None:
            from dnet_b200.utils.layer_manager import LayerManager

            layer_manager = LayerManager(model_metadata, assigned_layers, thread_pool_size=int(prefetch_threads or 2),
                                         use_mxload_fastpath=bool(use_mxload_fastpath), prefetch_mode=prefetch_mode,
                                         stage_host=stage_host, keep_host_records=keep_host_records)
        self.layer_manager = layer_manager
        # does the manager's loader take an HBM slot to recycle?  (injected test managers take only the layer id)
        try:
            self._slot_aware = "slot" in inspect.signature(layer_manager.load_layer_to_gpu).parameters
        except (TypeError, ValueError):
            self._slot_aware = False
        self._free_slots: List[Tuple[Any, Any]] = []      # (HBM slot tensor, event after its last compute read | None)
        self._release_events: Dict[int, Any] = {}         # layer -> event recorded after its last compute use
        logger.info("WeightCache resident budget: max_weights=%d", self.max_weights)

    # ------------------------------------------------------------------ lookups
    def get_weight(self, layer_id: int, *, inc_ref: bool = True) -> Optional[Dict[str, Any]]:
        role, payload = self._claim(layer_id, inc_ref)
        if role == _HIT:
            return payload
        if role == _JOIN:
            return self._join(layer_id, payload, inc_ref)
        return self._load_and_publish(layer_id, payload, inc_ref) if role == _LOAD else None

    def _claim(self, layer_id: int, inc_ref: bool):
        """Under the lock: decide what this caller is for ``layer_id``."""
        with self.lock:
            if self.closed:
                return None, None
            entry = self.cache.get(layer_id)
            if entry is not None:
                self._touch(layer_id, entry[0], inc_ref, create=False)
                return _HIT, entry[0]
            fut = self.loading_futures.get(layer_id)
            if fut is not None:
                return _JOIN, fut
            if len(self.cache) >= self.max_weights:
                self._evict_lru()
            fut = self.loading_futures[layer_id] = Future()
            return _LOAD, fut

    def _touch(self, layer_id: int, data, inc_ref: bool, create: bool) -> None:
        self.cache[layer_id] = (data, time.time())
        if inc_ref:
            self.reference_counts[layer_id] = self.reference_counts.get(layer_id, 0) + 1
        elif create:
            self.reference_counts.setdefault(layer_id, 0)

    def _join(self, layer_id: int, fut: Future, inc_ref: bool) -> Optional[Dict[str, Any]]:
        try:
            fut.result()
        except Exception as e:
            logger.error("Wait for layer %s load failed: %s", layer_id, e)
            return None
        with self.lock:
            entry = self.cache.get(layer_id)
            if entry is None or entry[0] is None:
                return None
            self._touch(layer_id, entry[0], inc_ref, create=True)
            return entry[0]

    def _load_and_publish(self, layer_id: int, fut: Future, inc_ref: bool) -> Optional[Dict[str, Any]]:
        t0 = time.perf_counter()
        try:
            data = self._materialise(layer_id)
        except Exception as e:
            with self.lock:
                if self.loading_futures.pop(layer_id, None) is fut and not fut.done():
                    fut.set_exception(e)
            if isinstance(e, OSError) and self.closed:
                logger.warning("Ignored load error for layer %s during shutdown: %s", layer_id, e)
            else:
                logger.exception("Failed to load weight %s: %s", layer_id, e)
            return None
        with self.lock:
            if self.closed:
                return None
            self._touch(layer_id, data, inc_ref, create=True)
            if self.loading_futures.pop(layer_id, None) is fut and not fut.done():
                fut.set_result(True)
        try:
            nbytes = sum(w.size_bytes for w in self.layer_manager.weight_info.get(layer_id, {}).values())
        except Exception:
            nbytes = 0
        logger.info("[PROFILE][MATERIALIZE] layer=%s ms=%.2f bytes=%.2fMB", layer_id, (time.perf_counter() - t0) * 1e3,
                    nbytes / 1_048_576)
        return data

    def _materialise(self, layer_id: int):
        lm = self.layer_manager
        if not self._slot_aware:
            return lm.load_layer_to_gpu(layer_id)
        with self.lock:
            slot, last_read = self._free_slots.pop() if self._free_slots else (None, None)
        try:
            return lm.load_layer_to_gpu(layer_id, slot=slot, wait_event=last_read)
        except BaseException:
            if slot is not None:                 # the slot stays usable: hand it back with its gate
                with self.lock:
                    self._free_slots.append((slot, last_read))
                last_read = None
            raise
        finally:
            if last_read is not None:
                self._destroy_event(last_read)   # the wait on it is enqueued; destroying is legal

    # ------------------------------------------------------------------ references / eviction
    def decrease_reference(self, layer_id: int, release_event=None):
        """Drop one reference; ``release_event`` (recorded on the compute stream after the layer's last
        kernel was enqueued) gates any later overwrite of the layer's HBM slot."""
        with self.lock:
            if layer_id in self.reference_counts:
                self.reference_counts[layer_id] -= 1
            if release_event is not None:
                stale = self._release_events.get(layer_id)
                if stale is not None:
                    self._destroy_event(stale)
                self._release_events[layer_id] = release_event

    def decrease_references(self, layer_ids: Iterable[int]) -> None:
        """decrease_reference for a whole run under one lock acquisition (decode hot path)."""
        with self.lock:
            rc = self.reference_counts
            for lid in layer_ids:
                if lid in rc:
                    rc[lid] -= 1

    def _drop(self, layer_id: int) -> None:
        """lock held: forget a resident layer and recycle its HBM slot."""
        try:
            self.layer_manager.release_layer(layer_id)
        except Exception:
            pass
        data, _ = self.cache.pop(layer_id)
        self.reference_counts.pop(layer_id, None)
        slot = data.get("_slot") if isinstance(data, dict) else None
        if slot is None:
            return
        self._free_slots.append((slot, self._release_events.pop(layer_id, None)))
        ready = data.get("_ready_event")
        if ready is not None:
            self._destroy_event(ready)
            data["_ready_event"] = None

    def _evict_lru(self) -> None:
        idle = [(stamp, lid) for lid, (_, stamp) in self.cache.items() if self.reference_counts.get(lid, 0) == 0]
        if idle:
            _, victim = min(idle)
            self._drop(victim)
            logger.info("Evicted layer %s from cache", victim)

    def evict_layer(self, layer_id: int) -> bool:
        with self.lock:
            if self.reference_counts.get(layer_id, 0) != 0:
                return False
            if layer_id in self.cache:
                self._drop(layer_id)
            return True

    def evict_layers(self, layer_ids: List[int]) -> int:
        done = 0
        for lid in layer_ids:
            try:
                done += 1 if self.evict_layer(lid) else 0
            except Exception:
                pass
        return done

    def get_resident_layers(self) -> List[int]:
        with self.lock:
            return [lid for lid, _ in sorted(self.cache.items(), key=lambda kv: kv[1][1])]

    # ------------------------------------------------------------------ host prefetch / lifecycle
    def prefetch_to_ram(self, layer_id: int):
        try:
            if self.layer_manager._prefetch_mode == "off":
                return None
            fut = self.prefetch_futures.get(layer_id)
            if fut is None or fut.done():
                fut = self.prefetch_futures[layer_id] = self.layer_manager.async_prefetch(layer_id)
            return fut
        except Exception:
            return None

    def cancel_all_prefetch(self):
        with self.lock:
            pending, self.prefetch_futures = list(self.prefetch_futures.values()), {}
        for fut in pending:
            try:
                if fut is not None and not fut.done():
                    fut.cancel()
            except Exception:
                pass

    def shutdown(self):
        with self.lock:
            self.closed = True
        self.cancel_all_prefetch()

    @staticmethod
    def _destroy_event(ev) -> None:
        try:
            from dnet_b200 import _cabi

            _cabi.load().dn_event_destroy(ev)
        except Exception:
            pass
