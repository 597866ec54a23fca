"""proto <-> ActivationMessage + pool staging (reference src/dnet/shard/codec.py:13-131).

Wire format byte-identical to the reference: ``Activation.data`` carries the raw row-major
element bytes, ``dtype`` names them ("tokens" = int32 ids, "bfloat16"/"float16" tensors).
The compressed branch (a dtype string containing a pipe character) is dead on the
reference's live path and is rejected loudly here.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from dnet_b200.core.types.messages import ActivationMessage
from dnet_b200.utils.logger import logger
from dnet_b200.utils.serialization import canonical_dtype, dtype_map, tensor_to_bytes, torch_dtype_map


class ActivationCodec:
    def __init__(self, runtime):
        self.runtime = runtime

    def deserialize(self, request) -> Optional[ActivationMessage]:
        if self.runtime.input_pool is None:
            logger.error("Shard %s: input pool not initialized", self.runtime.shard_id)
            return None
        activation = request.activation
        pool_id = None
        try:
            if "|" in activation.dtype:
                logger.error("compressed activations (dtype=%r) are not supported", activation.dtype)
                return None
            elif activation.dtype == "tokens":
                tokens = np.frombuffer(activation.data, dtype=np.int32)
                shp = (int(len(tokens)),)
                pool_id = self.runtime.input_pool.allocate_for_layer(layer_id=activation.layer_id, dtype=torch.int32, shape=shp)
                if pool_id is not None:
                    buffer = self.runtime.input_pool.get_buffer(pool_id)
                    buffer[: len(tokens)] = torch.from_numpy(tokens.copy())
                    msg = ActivationMessage.from_proto(request, pool_id)
                    msg.dtype = "tokens"
                    msg.shape = shp
                    return msg
            else:
                name = canonical_dtype(activation.dtype)
                expected = int(np.prod(activation.shape)) * np.dtype(dtype_map[name]).itemsize
                actual = len(activation.data)
                if expected != actual:
                    logger.error("Payload mismatch nonce=%s: exp=%d act=%d", request.nonce, expected, actual)
                    return None
                td = torch_dtype_map[name]
                pool_id = self.runtime.input_pool.allocate_for_layer(layer_id=activation.layer_id, dtype=td,
                                                                     shape=tuple(activation.shape))
                if pool_id is not None:
                    buffer = self.runtime.input_pool.get_buffer(pool_id)
                    src = torch.frombuffer(bytearray(activation.data), dtype=torch.uint8).view(td)
                    buffer[: src.numel()] = src
                    return ActivationMessage.from_proto(request, pool_id)
        except Exception as e:
            logger.error(f"Deserialization error for nonce {request.nonce}: {e}")
            if pool_id is not None:
                self.runtime.input_pool.release(pool_id)
            return None
        return None

    def tokens_message(self, nonce: str, ids, **kw) -> ActivationMessage:
        """What ``deserialize`` produces for a ``tokens`` frame, without the proto round trip: the ids staged
        in the (pinned) input pool.  For in-process drivers of ``policy.process`` (warm-up, calibration)."""
        ids = [int(t) for t in ids]
        n = len(ids)
        pid = self.runtime.input_pool.allocate_for_layer(layer_id=-1, dtype=torch.int32, shape=(n,))
        if pid is None:
            raise MemoryError("input pool exhausted")
        self.runtime.input_pool.get_buffer(pid)[:n] = torch.tensor(ids, dtype=torch.int32)
        more = bool(kw.pop("more", False))     # an intermediate chunk of a chunked prefill (policies/_cuda_common.py)
        return ActivationMessage(nonce=nonce, pool_id=pid, batch_size=0 if more else 1, shape=(n,), dtype="tokens", layer_id=-1,
                                 timestamp=0, node_origin=kw.pop("node_origin", "api"),
                                 callback_url=kw.pop("callback_url", "grpc://api:0"),
                                 temperature=kw.pop("temperature", 0.0), **kw)

    def serialize(self, msg: ActivationMessage, transport_config=None) -> bytes:
        """Device tensor -> wire bytes in the wire dtype (device->host copy + sync)."""
        shaped = msg.tensor
        if shaped is None:
            if self.runtime.output_pool is None:
                raise ValueError("No output pool and no tensor to serialize")
            output_buffer = self.runtime.output_pool.get_buffer(msg.pool_id)
            data_size = int(np.prod(msg.shape))
            shaped = output_buffer[:data_size].reshape(msg.shape)
        if shaped.dtype != self.runtime._wire_mx_dtype:
            shaped = shaped.to(self.runtime._wire_mx_dtype)
        if shaped.is_cuda:
            if getattr(msg, "ready_event", None) is not None:
                msg.ready_event.synchronize()
            elif getattr(self.runtime, "compute_stream", None) is not None:
                self.runtime.compute_stream.synchronize()
        data = tensor_to_bytes(shaped)
        msg.tensor = None
        return data
