"""Fakes for the GPU-less host-logic tests (the reference's tests/fakes, mlx-free)."""
from __future__ import annotations

import threading
from concurrent.futures import Future
from typing import Any

import torch


class FakeWeightSize:
    def __init__(self, size_bytes: int):
        self.size_bytes = size_bytes


class FakeModelMetadata:
    def __init__(self, weight_info, num_layers: int = 4):
        self.weight_info = weight_info
        self.num_layers = num_layers
        self.source = None


class FakeLayerManagerForCache:
    def __init__(self, weight_info=None, prefetch_mode: str = "off"):
        self.weight_info = weight_info or {}
        self._prefetch_mode = prefetch_mode
        self._released: list[int] = []
        self.loads: list[int] = []

    def load_layer_to_gpu(self, layer_id: int):
        self.loads.append(layer_id)
        return {"w": torch.tensor([float(layer_id)])}

    def release_layer(self, layer_id: int) -> bool:
        self._released.append(layer_id)
        return True

    def async_prefetch(self, layer_id: int):
        f: Future[Any] = Future()
        return f


class FakePool:
    def __init__(self):
        self.released = []

    def release(self, pid):
        self.released.append(pid)


class FakeModel:
    def __init__(self):
        self.unloaded = []

    def unload_layers(self, layers):
        self.unloaded.append(list(layers))


class FakeRuntimeForPolicy:
    def __init__(self, assigned):
        self.shard_id = "s0"
        self.assigned_layers = list(assigned)
        self._assigned_sorted = sorted(assigned)
        self._assigned_set = set(assigned)
        self.input_pool = FakePool()
        self.output_pool = FakePool()
        self.model = FakeModel()
        self._model_lock = threading.Lock()
