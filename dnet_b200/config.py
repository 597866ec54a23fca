"""Hot-path settings with the reference's env names (reference src/dnet/config.py:62-146).

Only the groups the shard forward reads are mirrored; values come from DNET_* env
variables so a reference .env keeps working.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from functools import lru_cache


def _env(name: str, default, cast=str):
    v = os.environ.get(name)
    if v is None or v == "":
        return default
    if cast is bool:
        return v.strip().lower() in {"1", "true", "yes", "on"}
    return cast(v)


@dataclass
class KVCacheSettings:  # reference config.py:62-77 (prefix DNET_KV_)
    mode: str = field(default_factory=lambda: _env("DNET_KV_MODE", "fp16"))
    bits: int = field(default_factory=lambda: _env("DNET_KV_BITS", 8, int))
    group_size: int = field(default_factory=lambda: _env("DNET_KV_GROUP_SIZE", 64, int))
    ttl_s: float = field(default_factory=lambda: _env("DNET_KV_TTL_S", 30.0, float))
    max_tokens: int = field(default_factory=lambda: _env("DNET_KV_MAX_TOKENS", 4096, int))
    pool_pages: int = field(default_factory=lambda: _env("DNET_KV_POOL_PAGES", 0, int))  # 0 = derive


@dataclass
class ComputeSettings:  # reference config.py:80-105 (prefix DNET_COMPUTE_)
    prefetch_mode: str = field(default_factory=lambda: _env("DNET_COMPUTE_PREFETCH_MODE", "off"))
    mxload_fastpath: bool = field(default_factory=lambda: _env("DNET_COMPUTE_MXLOAD_FASTPATH", False, bool))
    input_pool_mb: int = field(default_factory=lambda: _env("DNET_COMPUTE_INPUT_POOL_MB", 512, int))
    output_pool_mb: int = field(default_factory=lambda: _env("DNET_COMPUTE_OUTPUT_POOL_MB", 512, int))
    cuda_graphs: bool = field(default_factory=lambda: _env("DNET_COMPUTE_CUDA_GRAPHS", True, bool))
    pdl: bool = field(default_factory=lambda: _env("DNET_COMPUTE_PDL", False, bool))
    megakernel: bool = field(default_factory=lambda: _env("DNET_COMPUTE_MEGAKERNEL", True, bool))


@dataclass
class TransportSettings:  # reference config.py:108-127 (prefix DNET_TRANSPORT_)
    wire_dtype: str = field(default_factory=lambda: _env("DNET_TRANSPORT_WIRE_DTYPE", "fp16"))
    streaming: bool = True
    stream_idle_s: float = 2.0
    stream_backoff_s: float = 0.5
    # device-hop transport (dnet_b200): lanes = nonces in flight per ring, bulk slot = one prefill chunk
    hop_lanes: int = field(default_factory=lambda: _env("DNET_TRANSPORT_HOP_LANES", 16, int))
    hop_bulk_tokens: int = field(default_factory=lambda: _env("DNET_TRANSPORT_HOP_BULK_TOKENS", 512, int))
    sched_rounds_per_frame: int = field(default_factory=lambda: _env("DNET_TRANSPORT_SCHED_ROUNDS", 4, int))
    sched_frames_in_flight: int = field(default_factory=lambda: _env("DNET_TRANSPORT_SCHED_DEPTH", 3, int))
    lease_grace_s: float = field(default_factory=lambda: _env("DNET_TRANSPORT_LEASE_GRACE_S", 5e-4, float))
    # lm_head tensor-parallel over the ring during on-device decode: "auto" (rings of >= 4 shards), "on", "off"
    head_tp: str = field(default_factory=lambda: _env("DNET_TRANSPORT_HEAD_TP", "auto"))
    # extra schedule entries between a token's last layer and its head parts: 0 couples every shard to the last
    # shard's previous kernel (they all wait for its broadcast), 1 gives that broadcast a whole slot of slack
    head_tp_lag: int = field(default_factory=lambda: _env("DNET_TRANSPORT_HEAD_TP_LAG", 1, int))
    compress: bool = False
    compress_min_bytes: int = 65536


@dataclass
class TopologySettings:  # reference config.py:130-140 (prefix DNET_TOPOLOGY_)
    resident_windows: int = field(default_factory=lambda: _env("DNET_TOPOLOGY_RESIDENT_WINDOWS", 1, int))


@dataclass
class DnetSettings:
    kv_cache: KVCacheSettings = field(default_factory=KVCacheSettings)
    compute: ComputeSettings = field(default_factory=ComputeSettings)
    transport: TransportSettings = field(default_factory=TransportSettings)
    topology: TopologySettings = field(default_factory=TopologySettings)


@lru_cache(maxsize=1)
def get_settings() -> DnetSettings:
    return DnetSettings()
