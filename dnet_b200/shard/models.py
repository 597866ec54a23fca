"""Shard request/response models (reference src/dnet/shard/models.py:10-56)."""
from typing import Any, List, Literal, Optional

from pydantic import BaseModel, ConfigDict, Field


class ShardLoadModelRequest(BaseModel):
    """Request to load model with specified layers on shard."""

    model_config = ConfigDict(arbitrary_types_allowed=True, protected_namespaces=())

    # a local directory (reference: path or HF repo id) -- or, for synthetic benchmarks and
    # parity tests, a dnet_b200.utils.model.HostDictSource / SyntheticSource object
    model_path: Any = Field(..., description="Model path")
    total_layers: int = Field(..., description="Total number of layers in the model")
    layers: List[int] = Field(..., description="Layer indices to load on this shard")
    warmup: bool = Field(default=False)
    next_node: Optional[Any] = Field(default=None, description="Next shard in the ring")
    window_size: int = Field(..., description="Window size (computed from k)")
    residency_size: int = Field(..., description="Resident layers (n) allowed on GPU at once")
    kv_bits: Literal["4bit", "8bit", "fp16"] = Field(..., description="KV cache quantization")
    api_callback_address: str = Field(default="", description="API callback address (gRPC host:port)")


class ShardLoadModelResponse(BaseModel):
    success: bool
    message: str
    layers_loaded: List[int]
    load_time_ms: float


class ShardUnloadModelResponse(BaseModel):
    success: bool
    message: str
