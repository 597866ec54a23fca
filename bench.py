#!/usr/bin/env python
"""bench.py -- decode tok/s of the dnet pipelined-ring shard forward on B200.

  python bench.py --gpus 1 --steps K --warmup W            (our arm)
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...                      (CPU port of the reference path, all layers)

Workload (BASELINE.json configs[1] at N shards): Llama-3-8B dims, bf16, random-init weights
(synthetic), bs=1 decode after a 128-token synthetic prompt, temperature 0, kv fp16, wire
dtype bf16, contiguous layer split, k=1, N sequences in flight.  A "step" is one decoded token per
in-flight sequence.  Weights (15 GB/token) far exceed the 126 MB L2, so no L2 flush is needed
between iterations.  Every N (1 included) runs through the product's public transport API
(ShardNode / RingAdapter / ApiNode, see bench_ring.py): `value` is device-timed (CUDA events on each
rank's compute stream, max over ranks), `e2e` is the wall clock on the API side of the same steps
with every token read on the host.  Prints ONE JSON line on stdout (rank 0); diagnostics go to stderr.
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


def tensor_peak_tf() -> float:
    """dense bf16 TFLOP/s: MEASURED_PEAKS.json when present, else the profiling recipe's fallback"""
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            for k in ("bf16_tflops_sustained", "bf16_tflops"):      # a prompt is a long step: the sustained figure applies
                if k in d:
                    return float(d[k])
        except Exception:
            pass
    return 1600.0


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path, timed on the host cores
# ------------------------------------------------------------------------------------------
def _cpu_threads(m, H: int, L: int) -> int:
    """use as many host threads as actually help: a bs=1 GEMV is DRAM-bound and oversubscribing a big
    dual-socket box makes it slower, so probe a few counts on one projection and keep the best"""
    import torch

    ncpu = os.cpu_count() or 1
    probe_x = torch.randn(1, H).to(torch.bfloat16)
    name = "model.layers.1.mlp.gate_proj.weight" if L > 1 else "model.layers.0.mlp.gate_proj.weight"
    best_t, best_n = None, ncpu
    for n_thr in sorted({min(ncpu, c) for c in (8, 16, 32, 64, 128, ncpu)}):
        torch.set_num_threads(n_thr)
        m.linear(probe_x, "model.layers.0.mlp.gate_proj.weight")
        ta = time.perf_counter()
        for _ in range(3):
            m.linear(probe_x, name)
        dt = (time.perf_counter() - ta) / 3
        if best_t is None or dt < best_t:
            best_t, best_n = dt, n_thr
    torch.set_num_threads(best_n)
    return best_n


def cpu_full_depth_run(cfg: dict, steps: int, warmup: int, budget_s: float, weights=None, prompt=None, forced=None):
    """The CPU port on the WHOLE model (all layers + lm_head, ~15 GB of bf16 weights streamed per token):
    nothing is extrapolated.  With ``weights`` (the very tensors the GPU arm runs, copied to the host) the
    greedy tokens it produces are returned too, so the same leg is the full-depth parity witness.  Without,
    one layer's tensors are generated once and cloned per layer (distinct memory, so every layer streams
    from DRAM like distinct weights would; values do not matter for timing).  With ``forced`` (the GPU arm's
    greedy tokens) the oracle is teacher-forced: step i is fed forced[i], so one near-tie does not end the
    comparison, and ``regret_ulps[i]`` says how far below the oracle's best logit the GPU's next token sits."""
    import torch
    from oracle.llama_oracle import LlamaOracle, OracleConfig, OracleKV, make_weights, sample_greedy

    t0 = time.perf_counter()
    oc = OracleConfig.from_dict(cfg)
    L, H, V = cfg["num_hidden_layers"], cfg["hidden_size"], cfg["vocab_size"]
    if weights is None:
        w0 = make_weights(oc, 0, layers=range(1), with_api=False)
        w = {}
        for l in range(L):
            for k, v in w0.items():
                w[k.replace("layers.0.", f"layers.{l}.")] = v.clone() if l else v
        g = torch.Generator().manual_seed(0)
        w["model.embed_tokens.weight"] = torch.randn(V, H, generator=g).to(torch.bfloat16)
        w["lm_head.weight"] = (torch.randn(V, H, generator=g) * 0.02).to(torch.bfloat16)
        w["model.norm.weight"] = torch.ones(H, dtype=torch.bfloat16)
    else:
        w = weights
    log(f"cpu weights for {L} layers + head ready in {time.perf_counter() - t0:.1f}s")
    m = LlamaOracle(oc, w)
    threads = _cpu_threads(m, H, L)
    kv = {l: OracleKV() for l in range(L)}
    if prompt is None:
        gp = torch.Generator().manual_seed(1234)
        prompt = torch.randint(0, V, (PROMPT_LEN,), generator=gp).tolist()
    x = m.embed(torch.tensor(prompt, dtype=torch.int32))
    for l in range(L):
        x = m.apply_single_layer(l, x, kv[l]).to(torch.bfloat16)
    first = sample_greedy(m.lm_project(m.normalize(x)[-1:])[0], True, 0)
    tok, n, t_tot = first.token_id, 0, 0.0
    tokens, gaps, regret = [tok], [], []
    t_start = time.perf_counter()
    for i in range(warmup + steps):
        a = time.perf_counter()
        if forced is not None and i < len(forced):
            tok = int(forced[i])
        x = m.embed(torch.tensor([tok], dtype=torch.int32))
        for l in range(L):
            x = m.apply_single_layer(l, x, kv[l]).to(torch.bfloat16)   # per-layer cast to the wire dtype
        logits = m.lm_project(m.normalize(x))[0]
        tok = sample_greedy(logits, True, 0).token_id
        b = time.perf_counter()
        top2 = torch.topk(logits.to(torch.float32), 2).values
        ulp = max(float(top2[0].abs()) * 2.0 ** -8, 1e-30)
        gaps.append(float(top2[0] - top2[1]) / ulp)   # top-1/top-2 gap in bf16 ulps
        if forced is not None and i + 1 < len(forced):
            regret.append(float(top2[0] - logits[int(forced[i + 1])].to(torch.float32)) / ulp)
        tokens.append(tok)
        if i >= warmup:
            t_tot += b - a
            n += 1
        if time.perf_counter() - t_start > budget_s and n >= 3:
            break
    return n / t_tot, {"steps_timed": n, "ms_per_token": t_tot / n * 1e3, "threads": threads, "host_cpus": os.cpu_count() or 1,
                       "layers": L, "extrapolated": False, "tokens": tokens, "gap_ulps": gaps, "regret_ulps": regret}


def cpu_baseline_leg(args, cfg: dict, rt=None, prompt=None, gpu_tokens=None):
    """cpu_baseline of our arm: the full-depth CPU port, bounded to ~cpu_budget seconds of stepping.  At N=1
    it runs on the GPU arm's own weights (copied device -> host) and prompt, which makes it the full-depth
    parity witness as well: its greedy tokens are compared with the GPU's, step by step until they part."""
    weights = None
    if rt is not None:
        try:
            weights = {}
            for l in range(cfg["num_hidden_layers"]):
                for k, v in rt.policy.weight_cache.cache[l][0].items():
                    if not k.startswith("_"):
                        weights["model." + k] = v.cpu()
            for k, v in rt._api_tensors.items():
                weights[("model." if not k.startswith("lm_head") else "") + k] = v.cpu()
        except Exception as e:
            log(f"could not copy the GPU arm's weights to the host ({e}); timing the CPU port on cloned layers")
            weights = None
    try:
        tps, detail = cpu_full_depth_run(cfg, 16, 0, budget_s=args.cpu_budget, weights=weights, prompt=prompt,
                                         forced=gpu_tokens if weights is not None else None)
    except MemoryError as e:   # a box without ~17 GB of free host RAM
        log(f"cpu baseline skipped: {e}")
        return None, None
    toks, gaps, regret = detail.pop("tokens"), detail.pop("gap_ulps"), detail.pop("regret_ulps")
    cpu = {"value": tps, "unit": "tok/s", "cores": detail["threads"], "kind": "port",
           "sample": f"all {detail['layers']} layers + lm_head per step at a {PROMPT_LEN}-token context (nothing extrapolated), "
                     f"{detail['steps_timed']} decode steps" + (" on the GPU arm's weights" if weights is not None else ""),
           "detail": detail}
    log("cpu_baseline:", json.dumps(cpu))
    parity = None
    if weights is not None and gpu_tokens:
        # toks[0] = the oracle's first token after its own prefill; toks[i+1] = its argmax after being FED gpu_tokens[i]
        n = min(len(toks), len(gpu_tokens))
        agree = [toks[i] == gpu_tokens[i] for i in range(n)]
        miss = [i for i in range(n) if not agree[i]]
        parity = {"what": "greedy token ids, GPU (k_shard_step through the ring transport) vs the CPU oracle on the same weights, "
                          f"full depth ({detail['layers']} layers), prompt {PROMPT_LEN} tokens; the oracle is teacher-forced with the "
                          "GPU's tokens, so every step is compared on the same context",
                  "compared": n, "identical": sum(agree), "mismatch_steps": miss,
                  # for a mismatching step: how far below the oracle's best logit the GPU's choice sits (bf16 ulps of that logit)
                  "regret_ulps_at_mismatch": [round(regret[i - 1], 3) for i in miss if 0 < i <= len(regret)],
                  "max_regret_ulps": round(max(regret), 3) if regret else None,
                  "min_oracle_top2_gap_ulps": round(min(gaps), 3) if gaps else None,
                  "criterion": "every GPU token is the oracle's argmax or within 2 bf16 ulps of it",
                  "pass": bool((not regret or max(regret) <= 2.0) and (agree[0] if n else True))}
        log("full-depth parity:", json.dumps(parity))
    return cpu, parity


def run_reference(args, rank: int, world: int) -> None:
    if rank != 0:
        return
    cfg = dict(LLAMA3_8B)
    if args.layers:
        cfg["num_hidden_layers"] = args.layers
    tps, detail = cpu_full_depth_run(cfg, max(args.steps, 3), min(args.warmup, 2), budget_s=90.0)
    cores = detail["threads"]
    sample = (f"all {detail['layers']} layers + lm_head per step at a {PROMPT_LEN}-token context (nothing extrapolated), "
              f"{detail['steps_timed']} decode steps, bounded to 90 s")
    out = {
        "impl": "reference", "metric": METRIC, "value": tps, "unit": "tok/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": detail["ms_per_token"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "Llama-3-8B bf16 bs=1 decode (BASELINE configs[1]), reference path = CPU port of "
                               "FitInMemoryPolicy.process + mlx_lm llama block (mlx is not installable here)",
                   "prompt_len": PROMPT_LEN, "mlx_importable": _mlx_probe()},
        "cpu_baseline": {"value": tps, "unit": "tok/s", "cores": cores, "kind": "port", "sample": sample,
                         "detail": detail},
        "e2e": {"value": tps, "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)


def _mlx_probe() -> dict:
    """if the run box had mlx + mlx_lm, the real reference block could be a third witness."""
    out = {}
    for mod in ("mlx", "mlx.core", "mlx_lm"):
        try:
            __import__(mod)
            out[mod] = True
        except Exception as e:
            out[mod] = f"no ({type(e).__name__})"
    return out


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def measure_hop(rt, ad, barrier, rank: int, world: int, iters: int = 2000):
    """Ring-hop latency, measured directly: an 8 KiB activation + flag token travels `iters` times around the
    ring on a spare lane (dn_hop_ring_probe); the origin's globaltimer covers all trips, so no cross-GPU clock
    offset enters.  Returns microseconds per hop (payload stores + system-scope flag release -> acquire)."""
    import torch
    from dnet_b200 import _cabi

    lib = _cabi.load()
    hop = ad.hop
    lane = hop.n_lanes - 1                      # never handed to a request here (lanes are taken from 0 upwards)
    out = torch.zeros(1, dtype=torch.int64).pin_memory()
    res = {}
    base = 0
    for nbytes, name in ((hop.hidden * 2, "activation_8k"), (16, "token_16b")):
        rt.compute_stream.synchronize()
        barrier()
        _cabi.check(lib.dn_hop_ring_probe(hop.rx.slot(lane), hop.rx.flag(lane), hop.tx_slot(lane), hop.tx_flag(lane), nbytes,
                                          base, iters, 1 if rank == 0 else 0, 5000, out.data_ptr(), rt.compute_stream_ptr))
        rt.compute_stream.synchronize()
        barrier()
        base += iters
        if rank == 0:
            ns = int(out[0].item())
            res[name] = ns / (iters * world) / 1e3 if ns > 0 else None
    if rank == 0:
        res["how"] = (f"{iters} trips of (payload + flag) around the {world}-GPU ring on a spare lane, timed on rank 0's "
                      "globaltimer; us per hop = elapsed / (trips x ring size)")
        return res
    return None


def single_gpu_extras(args, rt, cfg: dict, K: int, ms: float) -> dict:
    """N=1 only: every per-op kernel timed alone (CUDA events between the five launches of a layer) for the
    per-kernel view, and the ncu-measured DRAM traffic of the dominant kernel."""
    import torch
    from dnet_b200 import _cabi
    from dnet_b200.shard.codec import ActivationCodec

    lib = _cabi.load()
    peak, _ = peaks()
    kb = kernel_bytes(cfg)
    L = cfg["num_hidden_layers"]
    names = ["qkv_rope_append", "attention", "o_proj_residual", "gate_up_swiglu", "down_residual"]
    g = torch.Generator().manual_seed(1234)
    prompt = torch.randint(0, cfg["vocab_size"], (PROMPT_LEN,), generator=g).tolist()
    # (the adapter's egress worker takes the emitted first token; the local:// sink drops it)
    rt.policy.process(ActivationCodec(rt).tokens_message("prof", prompt, req_logprobs=True, callback_url="local://"))
    rt.compute_stream.synchronize()
    nsp = rt.get_or_make_kv("prof")
    acc, reps = [0.0] * 5, 0
    out5 = (C.c_float * 5)()
    for rep in range(3):
        for l in range(L):
            _cabi.check(lib.dn_layer_forward_timed(rt.model._h, l, nsp.x1.data_ptr(), 1, nsp.kv._h, rt.compute_stream_ptr, out5))
            if rep > 0:
                for i in range(5):
                    acc[i] += out5[i]
                reps += 1
        nsp.kv.advance(1, rt.compute_stream_ptr)
    hm, head_ms = C.c_float(), 0.0
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
    log("per-kernel:", json.dumps(kern))
    traffic, src = None, None
    tp = ROOT / "profiles" / "ncu_traffic.json"
    if tp.exists():
        try:
            tj = json.loads(tp.read_text())
            traffic = tj.get("k_shard_step_dram_bytes_per_launch")
            src = tj.get("k_shard_step_source", "profiles/ncu_traffic.json (dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full)")
        except Exception:
            pass
    G = cfg["num_attention_heads"] // cfg["num_key_value_heads"]
    return {"kernels": kern, "traffic": traffic, "traffic_source": src,
            "kernel_name": f"k_shard_step<{G}> (whole decode step: {L} layers + lm_head, one launch per token)"}


def run_ours(args, rank: int, local_rank: int, world: int) -> None:
    import torch

    torch.cuda.set_device(local_rank)
    from dnet_b200 import _cabi

    _cabi.init(local_rank)
    import bench as _b                     # bench_ring imports `bench`; this file runs as __main__
    _b._REAL_STDOUT, _b.PROMPT_LEN = _REAL_STDOUT, PROMPT_LEN
    if args.config == "swap":
        from bench_swap import run_swap
        return run_swap(args, rank, local_rank, world)
    if args.config == "moe":
        from bench_moe import run_moe
        return run_moe(args, rank, local_rank, world)
    if args.config == "prefill":
        from bench_prefill import run_prefill
        return run_prefill(args, rank, local_rank, world)
    from bench_ring import run_ring
    return run_ring(args, rank, local_rank, world)


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
    ap.add_argument("--l2-prefetch-kb", type=int, default=64)
    ap.add_argument("--prompt-len", type=int, default=128)
    ap.add_argument("--mk-flags", type=int, default=0)
    ap.add_argument("--inflight-hi", type=int, default=-1)
    ap.add_argument("--park", type=int, default=-1, help="step kernel: TMEM parking of ready stages during grid barriers (-1 = library default)")
    ap.add_argument("--inflight", type=int, default=-1, help="step kernel: cap on ring stages with loads outstanding (-1 = library default)")
    ap.add_argument("--attn-chunk", type=int, default=0, help="step kernel: min tokens per attention split (0 = library default)")
    ap.add_argument("--pf-depth", type=int, default=-1, help="megakernel L2 prefetch look-ahead (ring stages); -1 = library default")
    ap.add_argument("--in-flight", type=int, default=0, help="sequences in flight at N>1 (default N)")
    ap.add_argument("--split", default="balanced", choices=["balanced", "equal"],
                    help="N>1: contiguous layer split balanced by streamed bytes (lm_head counted) or equal layer counts")
    ap.add_argument("--no-single", action="store_true", help="skip the single-sequence latency view")
    ap.add_argument("--head-tp", default="auto", choices=["auto", "on", "off"],
                    help="N>1: lm_head tensor-parallel over the ring's shards (auto: rings of >= 4 shards)")
    ap.add_argument("--sched-rounds", type=int, default=8, help="decode rounds per schedule frame (head shard's RingAdapter)")
    ap.add_argument("--sched-depth", type=int, default=4, help="schedule frames in flight")
    ap.add_argument("--config", default="decode", choices=["decode", "swap", "prefill", "moe"],
                    help="decode: BASELINE configs[1] (the bench contract); swap: configs[3], Llama-3-70B layer swap (bench_swap.py)")
    ap.add_argument("--prefill-len", type=int, default=32768, help="--config prefill: prompt tokens")
    ap.add_argument("--prefill-chunk", type=int, default=512, help="--config prefill: tokens per chunk frame")
    ap.add_argument("--swap-window", type=int, default=4, help="--config swap: window_size = residency_size (HBM layer slots per window)")
    ap.add_argument("--swap-resident-windows", type=int, default=1)
    ap.add_argument("--attn-tc", type=int, default=-1, help="step kernel: tensor-core attention phase at long contexts (-1 = library default: on)")
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
