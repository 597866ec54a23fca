"""Per-SM calibration of the step kernel's row partition.

With an equal split every SM streams the same bytes, yet the rate at which an SM can pull its share
out of HBM differs persistently between SMs (position relative to the two dies / HBM stacks:
profiles/r01_step_phase_times*.txt show the same SMs 15-20 % slower in every layer), and every grid
barrier waits for the slowest.  ``calibrate`` runs a few decode steps with the kernel's phase stamps
on (TMEM parking off, so every SM starts a phase with the same pre-filled bytes), converts each SM's
gate/up + down consume time into one speed per SM, re-partitions the rows of all four phases
proportionally (damped, a few rounds), and keeps the table only if a timed A/B says the step got
faster.  The result is static (dn_step_set_bounds); outputs do not depend on it.
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


_DBG_WORDS = 32          # MK_DBG_WORDS


def _time_steps(rt, step, n: int = 12) -> float:
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        step()
    with torch.cuda.stream(rt.compute_stream):
        e0.record()
        for _ in range(n):
            step()
        e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / n


def calibrate(rt, nonce: str = "__calib__", rounds: int = 3, steps_per_round: int = 2, damping: float = 0.7) -> List[np.ndarray]:
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
    equal = [_equal_bounds(rows[i], align[i], sms) for i in range(4)]
    bounds = [b.copy() for b in equal]
    arr = (C.c_int32 * len(run))(*run)
    s = rt.compute_stream_ptr
    buf = (C.c_uint64 * (sms * L * _DBG_WORDS))()

    def step():
        _cabi.check(lib.dn_shard_step(model._h, arr, L, ns.x1.data_ptr(), ns.kv._h, 0, 0, None, None, None, 1, s))

    def apply(bs):
        tbl = np.concatenate(bs).astype(np.int32)
        _cabi.check(lib.dn_step_set_bounds(model._h, tbl.ctypes.data_as(C.POINTER(C.c_int32))))

    apply(equal)
    t_equal = _time_steps(rt, step)
    weight = np.ones(sms)                       # relative share of each SM
    for _ in range(rounds):
        lib.dn_set_option(b"park", 0)
        lib.dn_set_option(b"mk_debug", 1)
        dur = np.zeros(sms)
        for _ in range(steps_per_round):
            step()
            rt.compute_stream.synchronize()
            lib.dn_step_debug(model._h, buf, sms * L * _DBG_WORDS, s)
            a = np.frombuffer(buf, dtype=np.uint64).reshape(sms, L, _DBG_WORDS).astype(np.int64)
            lo = min(2, L - 1)
            for (i0, i1) in _STAMP_PAIRS[2:]:                 # gate/up and down: 85 % of the bytes
                dur += (a[:, lo:, i1] - a[:, lo:, i0]).mean(axis=1)
        lib.dn_set_option(b"mk_debug", 0)
        lib.dn_set_option(b"park", 1)
        got = sum(np.diff(bounds[ph]) * (cfg["hidden_size"] if ph == 2 else cfg["intermediate_size"]) for ph in (2, 3)).astype(np.float64)
        speed = got / np.maximum(dur, 1.0)
        target = speed / speed.mean()
        weight = (1.0 - damping) * weight + damping * target
        weight /= weight.mean()
        for ph in range(4):
            total = rows[ph] // align[ph]
            new = weight / weight.sum() * total
            iu = np.floor(new).astype(np.int64)
            rem = int(total - iu.sum())
            order = np.argsort(-(new - iu))
            iu[order[:rem]] += 1
            bounds[ph] = np.concatenate([[0], np.cumsum(iu)]) * align[ph]
            assert bounds[ph][-1] == rows[ph]
        apply(bounds)
    t_cal = _time_steps(rt, step)
    if t_cal >= t_equal:                         # no gain on this GPU: keep the equal split
        apply(equal)
        bounds = equal
    rt.calibration = {"ms_equal": t_equal, "ms_calibrated": t_cal, "kept": bool(t_cal < t_equal),
                      "share_min": float(weight.min()), "share_max": float(weight.max())}
    rt.release_nonce(nonce)
    return bounds
