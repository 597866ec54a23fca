"""Ring model registry (reference src/dnet/core/models/__init__.py:13-35)."""
from typing import Any, List, Optional

from .base import BaseRingModel, KVHandle
from .llama import LlamaRingModel, MixtralRingModel, Qwen2RingModel


def _subclasses(cls):
    for s in cls.__subclasses__():
        yield s
        yield from _subclasses(s)


def get_ring_model(model_type: str, model_config: Any, assigned_layers: Optional[List[int]] = None,
                   is_api_layer: bool = False, **kw) -> BaseRingModel:
    """subclass_where(BaseRingModel, model_type=...) of the reference (utils/loader.py:7-20)."""
    for c in _subclasses(BaseRingModel):
        if getattr(c, "model_type", None) == model_type:
            return c(model_config, assigned_layers=assigned_layers, is_api_layer=is_api_layer, **kw)
    raise ValueError(f"Unsupported model type: {model_type}")


__all__ = ["BaseRingModel", "KVHandle", "LlamaRingModel", "MixtralRingModel", "Qwen2RingModel", "get_ring_model"]
