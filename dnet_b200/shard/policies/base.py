"""Execution-strategy plug-in: registry + ComputePolicy ABC
(reference src/dnet/shard/policies/base.py:9-142).  Names, constructor, abstract methods
and helper semantics are the reference's; the weight "version" is the data pointer of the
layer's first tensor (the reference uses id() of the first array)."""
from __future__ import annotations

import bisect
from abc import ABC, abstractmethod
from typing import Any, Dict, List, Optional

from dnet_b200.core.types.messages import ActivationMessage
from dnet_b200.utils.logger import logger

POLICY_REGISTRY: dict[str, type["ComputePolicy"]] = {}


def register_policy(mode: str):
    def deco(cls: type[ComputePolicy]):
        POLICY_REGISTRY[mode] = cls
        return cls

    return deco


def make_policy(mode: str, runtime, resident_windows: int) -> "ComputePolicy":
    m = (mode or "fit").strip().lower()
    cls = POLICY_REGISTRY.get(m)
    if cls is None:
        raise ValueError(f"Unsupported compute mode: {m}")
    return cls(runtime, resident_windows)


class ComputePolicy(ABC):
    """Abstract compute policy for ShardRuntime"""

    def __init__(self, runtime, resident_windows: int):
        self.runtime = runtime
        self.weight_cache = None
        self._prepared_by_nonce: Dict[str, tuple[list[int], Any]] = {}
        self._resident_windows = resident_windows
        self._recent_windows: List[List[int]] = []
        self._defer_unload = True
        self._await_next_ready = False
        self._warmup_keep_flag = False
        self._warmup_completed = False
        self._bound_versions: Dict[int, int] = {}
        self._mode: Optional[str] = None
        self.window_size = 0

    @abstractmethod
    def process(self, req: ActivationMessage): ...

    @abstractmethod
    def configure_policy_for_model(self, req): ...

    @abstractmethod
    def clear(self): ...

    # ---- helpers shared by the fit / offload policies.  Behaviour per SURVEY.md section 8(a) row a5 and the
    #      reference's policy tests (tests/subsystems/test_shard_policy_impl.py:163-257); written against that
    #      behaviour, over this repo's data structures (device-pointer versions, HBM slot recycling).
    @staticmethod
    def _next_local_layers(s: List[int], after_layer: int, count: int) -> List[int]:
        """The first ``count`` layers of the sorted local list that come after ``after_layer``."""
        if count <= 0:
            return []
        lo = bisect.bisect_right(s, after_layer)
        return list(s[lo:lo + count])

    def _delta_swap_eviction(self, window_layers: List[int], resident: List[int]) -> int:
        """sliding_fit: make room for ``window_layers`` inside a budget of ``window_size`` resident layers.
        Of the resident layers that are not part of the new window, the most recent ones that still fit
        beside it stay; the older ones are evicted from the cache (when nothing references them), unbound
        from the model and forgotten as bound versions.  Returns how many layers were evicted."""
        budget = max(1, int(self.window_size or 1))
        incoming = set(window_layers)
        leftovers = [lid for lid in resident if lid not in incoming]       # oldest first, as the cache lists them
        room = max(0, budget - len(window_layers))
        victims = leftovers[:max(0, len(leftovers) - room)]
        if not victims or self.weight_cache is None:
            return 0
        gone = []
        for lid in victims:
            try:
                if self.weight_cache.evict_layer(lid):
                    gone.append(lid)
            except Exception:
                pass                                                       # still referenced / racing: keep it
        if gone:
            try:
                self.runtime.model.unload_layers(gone)
            except Exception:
                return len(gone)
            for lid in gone:
                self._bound_versions.pop(lid, None)
        return len(gone)

    def _bind_layer_weights(self, window_layers: List[int], msg) -> Optional[Dict[str, Any]]:
        """Which tensors must be (re)bound before ``window_layers`` can run.

        {}    nothing to do -- in particular the decode fast path: every local layer fits in the window and
              all requested layers are already bound, so the weight cache is not even consulted;
        dict  name -> tensor of every layer whose resident copy differs from the one the model is bound to
              (each consulted layer now holds one more cache reference: the caller drops it after compute);
        None  a layer could not be materialised: the message's input buffer has been released, give up."""
        everything_fits = len(self.runtime._assigned_sorted) <= self.window_size
        if everything_fits and self._bound_versions.keys() >= set(window_layers):
            return {}
        cache = self.weight_cache
        pending: Dict[str, Any] = {}
        for lid in window_layers:
            weights = cache.get_weight(lid) if cache else None
            if weights is None:
                logger.error("Weight cache not initialized" if not cache else f"Failed to load weights for layer {lid}")
                self.runtime.input_pool.release(msg.pool_id)
                return None
            version = self._get_weight_version(weights)
            if self._bound_versions.get(lid) != version:
                self._bound_versions[lid] = version
                pending.update(weights)
        return pending

    @staticmethod
    def _get_weight_version(weights: dict) -> int:
        """Identity of a layer's resident copy: the device address of its first tensor (a reload into another
        HBM slot changes it; the reference uses id() of the first array).  -1 for an empty record."""
        for name, tensor in (weights or {}).items():
            if name.startswith("_"):
                continue                      # bookkeeping entries (_slot, _ready_event)
            ptr = getattr(tensor, "data_ptr", None)
            return int(ptr()) if ptr is not None else id(tensor)
        return -1
