"""Policy registry + plan_policy (reference src/dnet/shard/policies/__init__.py:1-77)."""
from dataclasses import dataclass
from typing import Type

from .base import ComputePolicy, POLICY_REGISTRY, make_policy, register_policy
from .noop import NoopPolicy
from . import fit_in_memory, offload  # noqa: F401  (registration side effects)
from .fit_in_memory import FitInMemoryPolicy
from .offload import OffloadPolicy


@dataclass
class PolicyPlan:
    mode: str
    window_size: int
    resident_windows: int
    is_sliding: bool
    policy_cls: Type[ComputePolicy]


def plan_policy(*, local_count: int, requested_w: int, residency_size: int, topology_config) -> PolicyPlan:
    requested_w = max(1, requested_w)
    n_residency = max(1, residency_size)
    if n_residency < requested_w:
        mode, sliding, resident_windows = "offload", True, 1
        window_size = max(1, min(n_residency, local_count))
    else:
        if requested_w >= local_count:
            mode, sliding, resident_windows = "fit", False, 9999
            window_size = local_count
        else:
            mode, sliding = "offload", False
            resident_windows = topology_config.resident_windows
            window_size = max(1, min(requested_w, local_count))
    cls = FitInMemoryPolicy if mode == "fit" else OffloadPolicy
    return PolicyPlan(mode=mode, window_size=window_size, resident_windows=resident_windows,
                      policy_cls=cls, is_sliding=sliding)


__all__ = ["make_policy", "register_policy", "ComputePolicy", "plan_policy", "PolicyPlan", "fit_in_memory",
           "offload", "NoopPolicy", "POLICY_REGISTRY", "FitInMemoryPolicy", "OffloadPolicy"]
