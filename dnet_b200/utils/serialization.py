"""dtype maps + raw tensor (de)serialisation (reference src/dnet/utils/serialization.py:7-105).

The wire format is byte-identical to the reference: the raw little-endian element bytes
of the row-major tensor, dtype named by string.  numpy has no bfloat16, so -- exactly as
the reference does (serialization.py:30) -- "bfloat16" maps to uint16 on the numpy side.
"""
from __future__ import annotations

import numpy as np
import torch

dtype_map = {
    "float32": np.float32, "float16": np.float16, "bfloat16": np.uint16,
    "int32": np.int32, "int64": np.int64, "uint8": np.uint8, "int8": np.int8, "uint16": np.uint16,
    "tokens": np.int32,
}
torch_dtype_map = {
    "float32": torch.float32, "float16": torch.float16, "bfloat16": torch.bfloat16,
    "int32": torch.int32, "int64": torch.int64, "uint8": torch.uint8, "int8": torch.int8,
    "tokens": torch.int32,
}
# aliases the reference accepts ("mlx.core.float16"-style strings, serialization.py:13-98)
for _k in list(torch_dtype_map):
    torch_dtype_map["mlx.core." + _k] = torch_dtype_map[_k]
    torch_dtype_map["torch." + _k] = torch_dtype_map[_k]
    if _k in dtype_map:
        dtype_map["mlx.core." + _k] = dtype_map[_k]
        dtype_map["torch." + _k] = dtype_map[_k]
safetensor_dtype_map = {
    "F32": np.float32, "F16": np.float16, "BF16": np.uint16, "I32": np.int32, "I64": np.int64,
    "U8": np.uint8, "I8": np.int8,
}
safetensor_torch_dtype = {
    "F32": torch.float32, "F16": torch.float16, "BF16": torch.bfloat16, "I32": torch.int32,
    "I64": torch.int64, "U8": torch.uint8, "I8": torch.int8,
    "U32": torch.int32,      # MLX packed quantised weights: read as 32-bit words (bit pattern preserved)
}


def canonical_dtype(name: str) -> str:
    n = str(name)
    for p in ("mlx.core.", "torch."):
        if n.startswith(p):
            n = n[len(p):]
    return n


def tensor_to_bytes(t: torch.Tensor) -> bytes:
    """bytes(memoryview(array)) of the reference: raw row-major element bytes."""
    t = t.detach().contiguous()
    if t.is_cuda:
        t = t.cpu()
    return t.view(torch.uint8).numpy().tobytes() if t.dtype != torch.uint8 else t.numpy().tobytes()


def bytes_to_tensor(data: bytes, dtype: str, shape) -> torch.Tensor:
    td = torch_dtype_map[canonical_dtype(dtype)]
    t = torch.frombuffer(bytearray(data), dtype=torch.uint8).view(td)
    return t.reshape(tuple(shape))
