"""ctypes binding of libdnet_b200.so (include/dnet_b200.h).

The product path has NO CPU fallback: importing this module without the built
library raises, and every call checks the return code and raises ``DnError`` with
``dn_last_error()``.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from pathlib import Path
from typing import Dict, List, Optional

_ROOT = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("DNET_B200_LIB", _ROOT / "lib" / "libdnet_b200.so"))
HEADER_PATH = _ROOT.parent / "include" / "dnet_b200.h"

DN_OK, DN_EINVAL, DN_ENOMEM, DN_ENOENT, DN_ECUDA, DN_ETIME, DN_ENOSPC = 0, -22, -12, -2, -5, -62, -28
(DN_W_Q, DN_W_K, DN_W_V, DN_W_O, DN_W_GATE, DN_W_UP, DN_W_DOWN, DN_W_LN1, DN_W_LN2,
 DN_W_QB, DN_W_KB, DN_W_VB, DN_W_COUNT) = range(13)

# suffix (reference utils/model.py:33-43 weight_info keys) -> slot in dn_bind_layer's array
SUFFIX_TO_SLOT: Dict[str, int] = {
    "self_attn.q_proj.weight": DN_W_Q,
    "self_attn.k_proj.weight": DN_W_K,
    "self_attn.v_proj.weight": DN_W_V,
    "self_attn.o_proj.weight": DN_W_O,
    "mlp.gate_proj.weight": DN_W_GATE,
    "mlp.up_proj.weight": DN_W_UP,
    "mlp.down_proj.weight": DN_W_DOWN,
    "input_layernorm.weight": DN_W_LN1,
    "post_attention_layernorm.weight": DN_W_LN2,
    "self_attn.q_proj.bias": DN_W_QB,
    "self_attn.k_proj.bias": DN_W_KB,
    "self_attn.v_proj.bias": DN_W_VB,
}


class DnError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libdnet_b200 error {code}: {msg}")
        self.code = code


class ModelCfg(C.Structure):
    _fields_ = [
        ("hidden", C.c_int32), ("n_heads", C.c_int32), ("n_kv_heads", C.c_int32), ("head_dim", C.c_int32),
        ("ffn", C.c_int32), ("vocab", C.c_int32), ("n_layers_total", C.c_int32), ("rms_eps", C.c_float),
        ("tie_embeddings", C.c_int32), ("dtype", C.c_int32), ("wire_dtype", C.c_int32),
        ("kv_page_tokens", C.c_int32), ("kv_pool_pages", C.c_int32), ("kv_bits", C.c_int32), ("kv_group", C.c_int32),
        ("n_experts", C.c_int32), ("top_k", C.c_int32),
    ]


class TpArgs(C.Structure):
    """dn_tp_args (include/dnet_b200.h): tensor-parallel lm_head arguments of dn_shard_step_tp"""
    _fields_ = [
        ("hp_x", C.c_void_p), ("hp_wait_flag", C.c_void_p), ("hp_seq", C.c_uint32),
        ("hp_dst", C.c_void_p), ("hp_dst_flag", C.c_void_p),
        ("bc_n", C.c_int32), ("bc_dst", C.c_void_p * 16), ("bc_flag", C.c_void_p * 16), ("bc_seq", C.c_uint32),
        ("mg_n", C.c_int32), ("mg_part", C.c_void_p), ("mg_flags", C.c_void_p), ("mg_seq", C.c_uint32),
        ("mg_kv", C.c_void_p), ("mg_token_out", C.c_void_p), ("mg_logprob_out", C.c_void_p),
        ("mg_slot", C.c_void_p), ("mg_slot_flag", C.c_void_p), ("mg_slot_seq", C.c_uint32),
    ]


def declared_symbols() -> List[str]:
    """Every function name include/dnet_b200.h declares (used by the export test)."""
    txt = HEADER_PATH.read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dn_[a-z0-9_]+)\s*\(", txt)))


_vp, _i, _u32, _sz = C.c_void_p, C.c_int, C.c_uint32, C.c_size_t
_PROTOS = {
    "dn_init": (_i, [_i]),
    "dn_last_error": (C.c_char_p, []),
    "dn_version": (C.c_char_p, []),
    "dn_set_option": (_i, [C.c_char_p, C.c_int64]),
    "dn_launch_count": (C.c_int64, []),
    "dn_device_sm_count": (_i, []),
    "dn_model_create": (_i, [C.POINTER(ModelCfg), C.POINTER(C.c_int32), _i, C.POINTER(C.c_float), C.POINTER(_vp)]),
    "dn_model_destroy": (_i, [_vp]),
    "dn_bind_layer": (_i, [_vp, _i, C.POINTER(_vp)]),
    "dn_bind_layer_experts": (_i, [_vp, _i, _vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), _i]),
    "dn_unbind_layer": (_i, [_vp, _i]),
    "dn_layer_is_bound": (_i, [_vp, _i]),
    "dn_bind_api": (_i, [_vp, _vp, _vp, _vp]),
    "dn_model_max_chunk": (_i, [_vp]),
    "dn_model_max_prefill_chunk": (_i, [_vp]),
    "dn_kv_create": (_i, [_vp, _i, C.POINTER(_vp)]),
    "dn_kv_free": (_i, [_vp]),
    "dn_kv_reset": (_i, [_vp, _vp]),
    "dn_kv_offset": (_i, [_vp]),
    "dn_kv_advance": (_i, [_vp, _i, _vp]),
    "dn_kv_seek": (_i, [_vp, _i, _vp]),
    "dn_kv_set_token": (_i, [_vp, C.c_int32, _vp]),
    "dn_kv_note_advance": (_i, [_vp, _i]),
    "dn_kv_token_ptr": (_vp, [_vp]),
    "dn_embed": (_i, [_vp, _vp, _i, _vp, _vp]),
    "dn_layer_forward": (_i, [_vp, _i, _vp, _i, _vp, _vp]),
    "dn_window_forward": (_i, [_vp, C.POINTER(C.c_int32), _i, _vp, _i, _vp, _vp]),
    "dn_shard_step": (_i, [_vp, C.POINTER(C.c_int32), _i, _vp, _vp, _i, _i, _vp, _vp, _vp, _i, _vp]),
    "dn_shard_step_hop": (_i, [_vp, C.POINTER(C.c_int32), _i, _vp, _vp, _i, _i, _vp, _vp, _i, _vp, _u32, _vp, _vp, _vp, _u32, _vp]),
    "dn_bind_head_slice": (_i, [_vp, _vp, _i, _i]),
    "dn_shard_step_tp": (_i, [_vp, C.POINTER(C.c_int32), _i, _vp, _vp, _i, _i, _vp, _u32, _vp, _vp, _vp, _u32,
                              C.POINTER(TpArgs), _vp]),
    "dn_step_error": (_i, [_vp, _vp]),
    "dn_step_error_clear": (_i, [_vp, _vp]),
    "dn_debug_scratch": (_i, [_vp, _i, _vp, _sz, _vp]),
    "dn_step_set_bounds": (_i, [_vp, C.POINTER(C.c_int32)]),
    "dn_step_debug": (_i, [_vp, C.POINTER(C.c_uint64), _sz, _vp]),
    "dn_layer_forward_timed": (_i, [_vp, _i, _vp, _i, _vp, _vp, C.POINTER(C.c_float)]),
    "dn_head_timed": (_i, [_vp, _vp, _i, _vp, C.POINTER(C.c_float)]),
    "dn_head_sample_greedy": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "dn_head_logits": (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    "dn_graph_begin": (_i, [_vp]),
    "dn_graph_end": (_i, [_vp, C.POINTER(_vp)]),
    "dn_graph_launch": (_i, [_vp, _vp]),
    "dn_graph_destroy": (_i, [_vp]),
    "dn_graph_num_nodes": (_i, [_vp]),
    "dn_hop_alloc": (_i, [_sz, C.POINTER(_vp)]),
    "dn_hop_free": (_i, [_vp]),
    "dn_hop_export": (_i, [_vp, C.POINTER(C.c_uint8)]),
    "dn_hop_import": (_i, [C.POINTER(C.c_uint8), C.POINTER(_vp)]),
    "dn_hop_close": (_i, [_vp]),
    "dn_enable_peer": (_i, [_i]),
    "dn_hop_send": (_i, [_vp, _vp, _sz, _vp, _u32, _vp]),
    "dn_hop_wait": (_i, [_vp, _u32, _u32, _vp, _vp]),
    "dn_hop_ring_probe": (_i, [_vp, _vp, _vp, _vp, _sz, _u32, _i, _i, _u32, _vp, _vp]),
    "dn_pinned_alloc": (_i, [_sz, C.POINTER(_vp)]),
    "dn_pinned_free": (_i, [_vp]),
    "dn_device_alloc": (_i, [_sz, C.POINTER(_vp)]),
    "dn_device_free": (_i, [_vp]),
    "dn_slot_prefetch": (_i, [_vp, _vp, _sz, _vp, _vp]),
    "dn_stream_create": (_i, [C.POINTER(_vp), _i]),
    "dn_stream_destroy": (_i, [_vp]),
    "dn_stream_sync": (_i, [_vp]),
    "dn_stream_wait_event": (_i, [_vp, _vp]),
    "dn_event_create": (_i, [C.POINTER(_vp), _i]),
    "dn_event_destroy": (_i, [_vp]),
    "dn_event_record": (_i, [_vp, _vp]),
    "dn_event_query": (_i, [_vp]),
    "dn_event_sync": (_i, [_vp]),
    "dn_event_elapsed_ms": (_i, [_vp, _vp, C.POINTER(C.c_float)]),
    "dn_memcpy_h2d": (_i, [_vp, _vp, _sz, _vp]),
    "dn_memcpy_d2h": (_i, [_vp, _vp, _sz, _vp]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """dlopen the library (no GPU needed for this) and attach prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m dnet_b200.build` "
            "(dnet_b200 has no CPU / PyTorch fallback for the shard forward)")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return (load().dn_last_error() or b"").decode("utf-8", "replace")


def check(rc: int) -> int:
    if rc < 0:
        raise DnError(rc, last_error())
    return rc


_inited_device: Optional[int] = None


def init(device: int = 0) -> None:
    """dn_init once per process (one shard process per GPU, like dnet-shard)."""
    global _inited_device
    lib = load()
    if _inited_device is not None:
        if _inited_device != device:
            raise DnError(DN_EINVAL, f"process already bound to cuda:{_inited_device}; one shard process per GPU")
        return
    check(lib.dn_init(device))
    _inited_device = device


def inited_device() -> Optional[int]:
    return _inited_device
