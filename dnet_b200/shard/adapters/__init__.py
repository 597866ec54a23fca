from .base import TopologyAdapter
from .ring import RingAdapter

__all__ = ["TopologyAdapter", "RingAdapter"]
