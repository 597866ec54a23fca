"""Per-SM calibration of the step kernel's row partition.

Measured on B200 (profiles/r01_step_phase_times.txt): with an equal split every SM streams the same
bytes, yet the time an SM needs for its share differs persistently by up to +-8% between SMs
(position relative to the two dies / HBM stacks), and every grid barrier waits for the slowest.
``calibrate`` runs a few decode steps with the kernel's phase stamps on, converts each SM's
consume time per phase into a speed, and re-partitions the rows proportionally (damped, 3 rounds).
The result is a static table (dn_step_set_bounds); outputs do not depend on it.
"""
from __future__ import annotations

import ctypes as C
from typing import List

import numpy as np

from dnet_b200 import _cabi

_STAMP_PAIRS = [(1, 2), (6, 7), (9, 10), (12, 13)]     # consume QKV, O, GATE/UP, DOWN (see MK_STAMP)


def _equal_bounds(rows: int, align: int, sms: int) -> np.ndarray:
    units = rows // align
    return np.array([(units * i) // sms * align for i in range(sms + 1)], dtype=np.int64)


def calibrate(rt, nonce: str = "__calib__", rounds: int = 3, steps_per_round: int = 2, damping: float = 0.6) -> List[np.ndarray]:
    """Calibrate ``rt.model``'s partition using decode steps on a scratch nonce (KV content is
    irrelevant).  Requires a loaded fit-mode runtime whose layers are bound."""
    lib = _cabi.load()
    model = rt.model
    cfg = model.config
    sms = int(lib.dn_device_sm_count())
    hd = 128
    nh, nkv = cfg["num_attention_heads"], cfg.get("num_key_value_heads", cfg["num_attention_heads"])
    rows = [(nh + 2 * nkv) * hd, cfg["hidden_size"], 2 * cfg["intermediate_size"], cfg["hidden_size"]]
    align = [2, 1, 2, 1]
    run = list(rt._assigned_sorted)
    L = len(run)
    ns = rt.get_or_make_kv(nonce)
    pol = rt.policy
    # make sure every layer is bound (first message normally does this)
    from dnet_b200.core.types.messages import ActivationMessage
    dummy = ActivationMessage(nonce=nonce, pool_id=-1, batch_size=1, shape=(1,), dtype="tokens", layer_id=-1, timestamp=0,
                              node_origin="", callback_url="")
    to_bind = pol._bind_layer_weights(run, dummy)
    if to_bind:
        from dnet_b200.shard.policies import _cuda_common as cc
        cc.wait_layers_ready(rt, pol.weight_cache, run)
        model.load_weights(list(to_bind.items()), strict=False)
    bounds = [_equal_bounds(rows[i], align[i], sms) for i in range(4)]
    arr = (C.c_int32 * len(run))(*run)
    s = rt.compute_stream_ptr
    buf = (C.c_uint64 * (sms * L * 16))()

    def step():
        _cabi.check(lib.dn_shard_step(model._h, arr, L, ns.x1.data_ptr(), ns.kv._h, 0, 0, None, None, None, 1, s))

    for _ in range(2):
        step()
    for _ in range(rounds):
        lib.dn_set_option(b"mk_debug", 1)
        dur = np.zeros((4, sms))
        for _ in range(steps_per_round):
            step()
            rt.compute_stream.synchronize()
            n = lib.dn_step_debug(model._h, buf, sms * L * 16, s)
            a = np.frombuffer(buf, dtype=np.uint64).reshape(sms, L, 16).astype(np.int64)
            lo = min(2, L - 1)
            for ph, (i0, i1) in enumerate(_STAMP_PAIRS):
                dur[ph] += (a[:, lo:, i1] - a[:, lo:, i0]).mean(axis=1)
        lib.dn_set_option(b"mk_debug", 0)
        flat = []
        for ph in range(4):
            if ph < 2:
                # QKV / O are short phases that start with a pre-filled ring: their time is not
                # proportional to rows, so only GATE/UP and DOWN (85% of the bytes) are re-partitioned
                flat.append(bounds[ph])
                continue
            units = np.diff(bounds[ph]) // align[ph]
            speed = units / np.maximum(dur[ph], 1.0)
            target = speed / speed.sum() * units.sum()
            new = (1.0 - damping) * units + damping * target
            iu = np.floor(new).astype(np.int64)
            rem = int(units.sum() - iu.sum())
            order = np.argsort(-(new - iu))
            iu[order[:rem]] += 1
            bounds[ph] = np.concatenate([[0], np.cumsum(iu)]) * align[ph]
            assert bounds[ph][-1] == rows[ph]
            flat.append(bounds[ph])
        tbl = np.concatenate(flat).astype(np.int32)
        _cabi.check(lib.dn_step_set_bounds(model._h, tbl.ctypes.data_as(C.POINTER(C.c_int32))))
    rt.release_nonce(nonce)
    return bounds
