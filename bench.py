#!/usr/bin/env python
"""bench.py -- decode tok/s of the dnet pipelined-ring shard forward on B200.

  python bench.py --gpus 1 --steps K --warmup W            (our arm)
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...                      (CPU port of the reference path)

Workload (BASELINE.json configs[1] at N shards): Llama-3-8B dims, bf16, random-init weights
(synthetic), bs=1 decode after a 128-token synthetic prompt, temperature 0, kv fp16, wire
dtype bf16, contiguous equal layer split, k=1.  A "step" is one decoded token per in-flight
sequence.  Weights (15 GB/token) far exceed the 126 MB L2, so no L2 flush is needed between
iterations.  Prints ONE JSON line on stdout (rank 0); diagnostics go to stderr.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
os.environ.setdefault("DNET_TRANSPORT_WIRE_DTYPE", "bf16")

LLAMA3_8B = dict(hidden_size=4096, num_attention_heads=32, num_key_value_heads=8, head_dim=128,
                 intermediate_size=14336, vocab_size=128256, num_hidden_layers=32, rms_norm_eps=1e-5,
                 rope_theta=500000.0, model_type="llama", tie_word_embeddings=False, torch_dtype="bfloat16")
METRIC = "decode tok/s Llama-3-8B bs=1"
PROMPT_LEN = 128


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


def layer_bytes(c) -> int:
    H, F, qd, kd = c["hidden_size"], c["intermediate_size"], c["num_attention_heads"] * 128, c["num_key_value_heads"] * 128
    return 2 * (qd * H + 2 * kd * H + H * qd + 3 * F * H + 2 * H)


def token_bytes(c) -> int:
    """all layers + lm_head + final norm + one embed row."""
    H, V = c["hidden_size"], c["vocab_size"]
    return c["num_hidden_layers"] * layer_bytes(c) + 2 * V * H + 2 * H + 2 * H


def kernel_bytes(c) -> dict:
    """algorithmic bytes per launch of each kernel at T=1 (weights + vectors in/out)."""
    H, F, V = c["hidden_size"], c["intermediate_size"], c["vocab_size"]
    qd, kd = c["num_attention_heads"] * 128, c["num_key_value_heads"] * 128
    return {
        "qkv_rope_append": 2 * ((qd + 2 * kd) * H + 2 * H + qd + 2 * kd),
        "attention": None,  # context dependent: 2 * 2 * kd * n
        "o_proj_residual": 2 * (H * qd + qd + 2 * H),
        "gate_up_swiglu": 2 * (2 * F * H + 2 * H + F),
        "down_residual": 2 * (H * F + F + 2 * H),
        "head_argmax": 2 * (V * H + 2 * H + V),
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception as e:
            log("nvidia-smi unavailable:", e)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()

    def summary(self, t0: float, t1: float) -> dict:
        sm, mx, reasons = [], 0.0, set()
        for t, line in self.rows:
            if not (t0 - 0.25 <= t <= t1 + 0.25):
                continue
            p = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(p[1])); mx = max(mx, float(p[2]))
            except Exception:
                continue
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for nm, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path, timed on the host cores
# ------------------------------------------------------------------------------------------
def cpu_port_run(cfg: dict, weights: dict, sample_layers: int, steps: int, warmup: int, budget_s: float):
    """FitInMemoryPolicy.process semantics on torch-CPU for `sample_layers` layers + lm_head at a
    PROMPT_LEN-token context; returns (tok/s extrapolated to all layers, detail)."""
    import torch
    from oracle.llama_oracle import LlamaOracle, OracleConfig, OracleKV, sample_greedy

    oc = OracleConfig.from_dict(cfg)
    m = LlamaOracle(oc, weights)
    # use as many host threads as actually help: a bs=1 GEMV is DRAM-bound and oversubscribing a
    # big dual-socket box makes it slower, so probe a few counts on one projection and keep the best
    ncpu = os.cpu_count() or 1
    probe_x = torch.randn(1, cfg["hidden_size"]).to(torch.bfloat16)
    best_t, best_n = None, ncpu
    for n_thr in sorted({min(ncpu, c) for c in (8, 16, 32, 64, 128, ncpu)}):
        torch.set_num_threads(n_thr)
        m.linear(probe_x, "model.layers.0.mlp.gate_proj.weight")
        t0 = time.perf_counter()
        for _ in range(3):
            m.linear(probe_x, "model.layers.0.mlp.gate_proj.weight")
        dt = (time.perf_counter() - t0) / 3
        if best_t is None or dt < best_t:
            best_t, best_n = dt, n_thr
    torch.set_num_threads(best_n)
    cpu_port_run.threads = best_n
    kv = {l: OracleKV() for l in range(sample_layers)}
    g = torch.Generator().manual_seed(1234)
    ids = torch.randint(0, cfg["vocab_size"], (PROMPT_LEN,), generator=g, dtype=torch.int32)
    x = m.embed(ids)
    for l in range(sample_layers):
        x = m.apply_single_layer(l, x, kv[l])
    tok = 1
    t_layers, t_head, n = 0.0, 0.0, 0
    t_start = time.perf_counter()
    for i in range(warmup + steps):
        a = time.perf_counter()
        x = m.embed(torch.tensor([tok], dtype=torch.int32))
        for l in range(sample_layers):
            x = m.apply_single_layer(l, x, kv[l]).to(torch.bfloat16)   # per-layer cast to the wire dtype
        b = time.perf_counter()
        r = sample_greedy(m.lm_project(m.normalize(x))[0], True, 0)
        c = time.perf_counter()
        tok = r.token_id
        if i >= warmup:
            t_layers += b - a; t_head += c - b; n += 1
        if time.perf_counter() - t_start > budget_s and n >= 3:
            break
    per_layer = t_layers / n / sample_layers
    head = t_head / n
    L = cfg["num_hidden_layers"]
    tps = 1.0 / (L * per_layer + head)
    return tps, {"steps_timed": n, "ms_per_layer": per_layer * 1e3, "ms_head": head * 1e3,
                 "ms_per_token_extrapolated": (L * per_layer + head) * 1e3, "threads": best_n, "host_cpus": ncpu}


def cpu_weights_random(cfg: dict, sample_layers: int):
    import torch
    from oracle.llama_oracle import OracleConfig, make_weights

    t0 = time.perf_counter()
    oc = OracleConfig.from_dict(cfg)
    w = make_weights(oc, 0, layers=range(sample_layers), with_api=False)
    g = torch.Generator().manual_seed(0)
    H, V = cfg["hidden_size"], cfg["vocab_size"]
    w["model.embed_tokens.weight"] = torch.randn(V, H, generator=g).to(torch.bfloat16)
    w["lm_head.weight"] = (torch.randn(V, H, generator=g) * 0.02).to(torch.bfloat16)
    w["model.norm.weight"] = torch.ones(H, dtype=torch.bfloat16)
    log(f"cpu weights for {sample_layers} layers + head generated in {time.perf_counter() - t0:.1f}s")
    return w


def run_reference(args, rank: int, world: int) -> None:
    if rank != 0:
        return
    cfg = dict(LLAMA3_8B)
    sample_layers = 2
    w = cpu_weights_random(cfg, sample_layers)
    tps, detail = cpu_port_run(cfg, w, sample_layers, args.steps, args.warmup, budget_s=60.0)
    cores = detail["threads"]
    sample = (f"{sample_layers} of {cfg['num_hidden_layers']} layers + lm_head per step at a {PROMPT_LEN}-token context, "
              f"{detail['steps_timed']} decode steps, per-token time extrapolated to all layers")
    out = {
        "impl": "reference", "metric": METRIC, "value": tps, "unit": "tok/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": detail["ms_per_token_extrapolated"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "Llama-3-8B bf16 bs=1 decode (BASELINE configs[1]), reference path = CPU port of "
                               "FitInMemoryPolicy.process + mlx_lm llama block (mlx is not installable here)",
                   "prompt_len": PROMPT_LEN},
        "cpu_baseline": {"value": tps, "unit": "tok/s", "cores": cores, "kind": "port", "sample": sample,
                         "detail": detail},
        "e2e": {"value": tps, "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def run_ours(args, rank: int, local_rank: int, world: int) -> None:
    import torch

    torch.cuda.set_device(local_rank)
    from dnet_b200 import _cabi
    from dnet_b200.shard.models import ShardLoadModelRequest
    from dnet_b200.shard.runtime import ShardRuntime
    from dnet_b200.utils.model import SyntheticSource
    from tests.helpers import token_message

    _cabi.init(local_rank)
    lib = _cabi.load()
    if world > 1:
        import bench as _b                     # bench_ring imports `bench`; this file runs as __main__
        _b._REAL_STDOUT, _b.PROMPT_LEN = _REAL_STDOUT, PROMPT_LEN
        from bench_ring import run_ring
        return run_ring(args, rank, local_rank, world)

    cfg = dict(LLAMA3_8B)
    if args.layers:
        cfg["num_hidden_layers"] = args.layers
    L = cfg["num_hidden_layers"]
    K, W = args.steps, args.warmup
    need = PROMPT_LEN + W + K + 16
    rt = ShardRuntime(shard_id=0)
    rt.kv_cache_config.max_tokens = need
    os.environ["DNET_KV_POOL_PAGES"] = str(((need + 63) // 64) * 4)
    from dnet_b200.config import get_settings
    get_settings.cache_clear()
    t0 = time.perf_counter()
    rt.load_model_core(ShardLoadModelRequest(model_path=SyntheticSource(cfg, seed=0), total_layers=L,
                                             layers=list(range(L)), window_size=L, residency_size=L, kv_bits="fp16"))
    lib.dn_set_option(b"pdl", 1 if args.pdl else 0)
    lib.dn_set_option(b"l2_prefetch_kb", args.l2_prefetch_kb)
    rt.use_megakernel = bool(args.megakernel)
    if args.pf_depth >= 0:
        lib.dn_set_option(b"pf_depth", args.pf_depth)
    lib.dn_set_option(b"mk_flags", args.mk_flags)
    if args.attn_chunk:
        lib.dn_set_option(b"attn_chunk", args.attn_chunk)
    if args.inflight >= 0:
        lib.dn_set_option(b"inflight", args.inflight)
    if args.inflight_hi >= 0:
        lib.dn_set_option(b"inflight_hi", args.inflight_hi)
    if args.park >= 0:
        lib.dn_set_option(b"park", args.park)
    pol = rt.policy
    g = torch.Generator().manual_seed(1234)
    prompt = torch.randint(0, cfg["vocab_size"], (PROMPT_LEN,), generator=g).tolist()
    # materialise + bind all layers, prefill both nonces through the public policy API
    def prefill(nonce):
        pol.process(token_message(rt, nonce, prompt, req_logprobs=True))
        return rt.activation_send_queue.get_nowait()
    first = prefill("dev")
    torch.cuda.synchronize()
    if args.megakernel and args.calibrate:
        from dnet_b200.shard.calibrate import calibrate
        tcal = time.perf_counter()
        calibrate(rt)
        log(f"step-kernel partition calibrated in {time.perf_counter() - tcal:.2f}s: {getattr(rt, 'calibration', None)}")
    log(f"model ready + prefill in {time.perf_counter() - t0:.1f}s; first token {first.token_id} lp {first.logprob}")
    run = list(range(L))
    stream = rt.compute_stream

    # ---------------- value: device-resident decode loop (CUDA graph replay, no host in the loop)
    ns = rt.get_or_make_kv("dev")
    ns.kv.set_token(first.token_id, rt.compute_stream_ptr)
    for _ in range(W):
        pol._graph_step(ns, ns.x1, True, run, True)
    stream.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    l0 = lib.dn_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    tw0 = time.perf_counter()
    e0.record(stream)
    for _ in range(K):
        pol._graph_step(ns, ns.x1, True, run, True)
    e1.record(stream)
    stream.synchronize()
    torch.cuda.synchronize()
    tw1 = time.perf_counter()
    ms = e0.elapsed_time(e1)
    launches = int(lib.dn_launch_count() - l0)
    value = K / ms * 1e3
    dev_last_token = int(ns.result_i32[0].item())
    clocks = sampler.summary(tw0, tw1)
    log(f"value: {value:.1f} tok/s, {ms / K:.4f} ms/step, {launches} kernel launches, clocks {clocks}")

    # ---------------- e2e: the reference-facing call (policy.process) with HOST buffers every step
    e2e = None
    if not args.no_e2e:
        firste = prefill("e2e")
        tok = firste.token_id
        seq = []
        for i in range(W + K):
            if i == W:
                torch.cuda.synchronize()
                te0 = time.perf_counter()
            pol.process(token_message(rt, "e2e", [tok], req_logprobs=True))   # 4-byte id from pinned host memory
            res = rt.activation_send_queue.get_nowait()                      # token + logprob read back on the host
            tok = res.token_id
            seq.append(tok)
        torch.cuda.synchronize()
        te1 = time.perf_counter()
        e2e_v = K / (te1 - te0)
        e2e = {"value": e2e_v, "unit": "tok/s", "h2d_bytes_per_step": 4, "d2h_bytes_per_step": 8,
               "api": "FitInMemoryPolicy.process(ActivationMessage) -> runtime.activation_send_queue",
               "tokens_match_device_loop": bool(seq[-1] == dev_last_token)}
        log(f"e2e: {e2e_v:.1f} tok/s; last token device-loop {dev_last_token} vs e2e {seq[-1]}")
    sampler.stop()

    # ---------------- live per-kernel times -> roofline of the dominant kernel
    peak, peak_src = peaks()
    kb = kernel_bytes(cfg)
    names = ["qkv_rope_append", "attention", "o_proj_residual", "gate_up_swiglu", "down_residual"]
    nsp = rt.get_or_make_kv("prof")
    prefill("prof")
    acc = [0.0] * 5
    reps = 0
    out5 = (C.c_float * 5)()
    for rep in range(3):
        for l in range(L):
            _cabi.check(lib.dn_layer_forward_timed(rt.model._h, l, nsp.x1.data_ptr(), 1, nsp.kv._h, rt.compute_stream_ptr, out5))
            if rep > 0:
                for i in range(5):
                    acc[i] += out5[i]
                reps += 1
        nsp.kv.advance(1, rt.compute_stream_ptr)
    hm = C.c_float()
    head_ms = 0.0
    for rep in range(4):
        _cabi.check(lib.dn_head_timed(rt.model._h, nsp.x1.data_ptr(), 1, rt.compute_stream_ptr, C.byref(hm)))
        if rep > 0:
            head_ms += hm.value / 3
    kern = {}
    ctx = PROMPT_LEN + 3
    for i, nm in enumerate(names):
        t = acc[i] / max(1, reps)
        b = kb[nm] if kb[nm] is not None else 2 * 2 * cfg["num_key_value_heads"] * 128 * ctx
        kern[nm] = {"ms": t, "bytes": b, "gbs": b / t / 1e6 if t > 0 else None, "frac": (b / t / 1e6) / peak if t > 0 else None}
    kern["head_argmax"] = {"ms": head_ms, "bytes": kb["head_argmax"], "gbs": kb["head_argmax"] / head_ms / 1e6,
                           "frac": kb["head_argmax"] / head_ms / 1e6 / peak}
    tb = token_bytes(cfg)
    traffic = None
    tp = ROOT / "profiles" / "ncu_traffic.json"
    tj = {}
    if tp.exists():
        try:
            tj = json.loads(tp.read_text())
        except Exception:
            tj = {}
    if args.megakernel:
        # dominant kernel = the persistent step kernel: one launch per token, timed by the CUDA
        # events around the K timed launches on the compute stream
        traffic = tj.get("k_shard_step_dram_bytes_per_launch")
        ach = tb / (ms / K) / 1e6
        roofline = {"bound": "hbm", "kernel": "k_shard_step<4> (whole decode step: 32 layers + lm_head, one launch per token)",
                    "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": tb, "launch_ms": ms / K,
                    "step": {"algorithmic_bytes_per_token": tb, "achieved_gbs": tb * value / 1e9,
                             "frac": tb * value / 1e9 / peak, "roofline_tok_s": peak * 1e9 / tb},
                    "per_op_kernels_timed_alone": kern}
    else:
        dom = kern["gate_up_swiglu"]
        traffic = tj.get("gate_up_swiglu_dram_bytes_per_launch")
        roofline = {"bound": "hbm", "kernel": "k_gemv<1,OpGateUp> (RMSNorm + gate/up GEMV + SwiGLU)", "achieved": dom["gbs"],
                    "peak": peak, "unit": "GB/s", "frac": dom["frac"], "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": dom["bytes"], "launch_ms": dom["ms"],
                    "step": {"algorithmic_bytes_per_token": tb, "achieved_gbs": tb * value / 1e9,
                             "frac": tb * value / 1e9 / peak, "roofline_tok_s": peak * 1e9 / tb},
                    "kernels": kern}
    log("per-kernel:", json.dumps(kern))

    # ---------------- CPU baseline: oracle port on a bounded sample of the same workload
    cpu = None
    if not args.no_cpu:
        sample_layers = min(2, L)
        w = {}
        for l in range(sample_layers):
            for k, v in pol.weight_cache.cache[l][0].items():
                if not k.startswith("_"):
                    w["model." + k] = v.cpu()
        for k, v in rt._api_tensors.items():
            w[("model." if not k.startswith("lm_head") else "") + k] = v.cpu()
        tps, detail = cpu_port_run(cfg, w, sample_layers, 64, 2, budget_s=args.cpu_budget)
        cpu = {"value": tps, "unit": "tok/s", "cores": detail["threads"], "kind": "port",
               "sample": f"{sample_layers} of {L} layers + lm_head per step at a {PROMPT_LEN}-token context, "
                         f"{detail['steps_timed']} decode steps, per-token time extrapolated to all layers",
               "detail": detail}
        log("cpu_baseline:", json.dumps(cpu))

    out = {
        "metric": METRIC, "value": value, "unit": "tok/s", "n_gpus": 1, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": f"Llama-3-8B bf16 bs=1 decode, 1 shard x {L} layers (BASELINE configs[1] at 1 shard)",
                   "prompt_len": PROMPT_LEN, "kv": "fp16 paged (64-token pages)", "wire_dtype": "bf16",
                   "l2": "inputs larger than L2 (15.0 GB of weights per step vs 126 MB L2); no flush",
                   "pdl": bool(args.pdl), "l2_prefetch_kb": args.l2_prefetch_kb,
                   "step_kernel": "k_shard_step (one persistent cooperative kernel per token)" if args.megakernel
                   else "per-op kernels replayed as a CUDA graph",
                   "step_error": int(lib.dn_step_error(rt.model._h, rt.compute_stream_ptr)),
                   "partition": "per-SM calibrated (dnet_b200.shard.calibrate)" if (args.megakernel and args.calibrate) else "equal",
                   "sequences_in_flight": 1},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
        "check": {"nonce0_token_after_steps": W + K, "token": dev_last_token},
    }
    emit(out)
    rt.unload_model_core()


_REAL_STDOUT = None


def _guard_stdout():
    """Everything any library prints to fd 1 (e.g. NCCL's version banner) goes to stderr; the one
    JSON line is written to the real stdout by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(obj) -> None:
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def main():
    _guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--layers", type=int, default=0, help="debug: fewer layers (invalid as a bench number)")
    ap.add_argument("--pdl", type=int, default=int(os.environ.get("DNET_COMPUTE_PDL", "0")))
    ap.add_argument("--megakernel", type=int, default=int(os.environ.get("DNET_COMPUTE_MEGAKERNEL", "1")))
    ap.add_argument("--l2-prefetch-kb", type=int, default=64)
    ap.add_argument("--prompt-len", type=int, default=128)
    ap.add_argument("--mk-flags", type=int, default=0)
    ap.add_argument("--inflight-hi", type=int, default=-1)
    ap.add_argument("--park", type=int, default=-1, help="step kernel: TMEM parking of ready stages during grid barriers (-1 = library default)")
    ap.add_argument("--inflight", type=int, default=-1, help="step kernel: cap on ring stages with loads outstanding (-1 = library default)")
    ap.add_argument("--attn-chunk", type=int, default=0, help="step kernel: min tokens per attention split (0 = library default)")
    ap.add_argument("--calibrate", type=int, default=0, help="per-SM row-partition calibration of the step kernel")
    ap.add_argument("--fused-hop", type=int, default=1, help="N>1: wait+step+hop in one kernel")
    ap.add_argument("--pf-depth", type=int, default=-1, help="megakernel L2 prefetch look-ahead (ring stages); -1 = library default")
    ap.add_argument("--in-flight", type=int, default=0, help="sequences in flight at N>1 (default N)")
    ap.add_argument("--split", default="balanced", choices=["balanced", "equal"],
                    help="N>1: contiguous layer split balanced by streamed bytes (lm_head counted) or equal layer counts")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    args = ap.parse_args()
    global PROMPT_LEN
    PROMPT_LEN = args.prompt_len
    if args.warmup < 3:
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
