"""Transport plug-in seam (reference src/dnet/shard/adapters/base.py:12-70): an adapter takes a
topology-agnostic ShardRuntime and wires a transport in (network -> adapter -> runtime) and out
(runtime -> adapter -> network).  Same abstract surface, so `Shard` and the servicer do not care
which transport is underneath."""
from __future__ import annotations

import asyncio
from abc import ABC, abstractmethod


class TopologyAdapter(ABC):
    def __init__(self, runtime, discovery):
        self.runtime = runtime
        self.discovery = discovery
        self.running = False

    @property
    @abstractmethod
    def ingress_q(self) -> asyncio.Queue:
        """frames admitted from the network (filled by the gRPC servicer)"""

    @property
    @abstractmethod
    def activation_computed_queue(self) -> asyncio.Queue:
        """non-final results on their way to the next shard"""

    @property
    @abstractmethod
    def activation_token_queue(self) -> asyncio.Queue:
        """final tokens on their way to the API"""

    @abstractmethod
    async def start(self): ...

    @abstractmethod
    async def ingress(self): ...

    @abstractmethod
    async def egress(self): ...

    @abstractmethod
    async def configure_topology(self, req): ...

    @abstractmethod
    async def reset_topology(self): ...

    @abstractmethod
    async def shutdown(self) -> None: ...
