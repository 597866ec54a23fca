"""The bench line contract (driver-facing): checked on the committed lines under profiles/, which are
verbatim stdout of `bench.py` runs on B200 (no GPU needed here)."""
import json
from pathlib import Path

import pytest

PROFILES = Path(__file__).resolve().parent.parent / "profiles"
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"}


def _load(name):
    p = PROFILES / name
    if not p.exists():
        pytest.skip(f"{name} not committed")
    return json.loads(p.read_text().strip().splitlines()[-1])


@pytest.mark.parametrize("name,n", [("r01_bench_n1.json", 1), ("r01_bench_n2_ring.json", 2), ("r01_bench_n4_ring.json", 4),
                                    ("r01_bench_n8_ring.json", 8)])
def test_our_arm_line(name, n):
    d = _load(name)
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["metric"] == "decode tok/s Llama-3-8B bs=1" and d["unit"] == "tok/s" and d["n_gpus"] == n
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "bf16" and d["data"] == "synthetic" and d["warmup"] >= 3
    assert "workload" in d["config"] and "model" not in d["config"]
    assert abs(d["value"] - n_seq(d) * 1e3 / d["ms_per_step"]) / d["value"] < 1e-6
    assert d["gpu_launches"] >= d["steps"]                       # one step kernel per token per shard at least
    e = d["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e) and e["h2d_bytes_per_step"] > 0
    assert 0 < e["value"] <= d["value"] * 1.01 and e["value"] != d["value"]
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm"
    assert 0.5 < r["frac"] < 1.1
    c = d["clocks"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(c)
    assert not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"]))
    if n == 1:
        assert abs(r["achieved"] / r["peak"] - r["frac"]) < 1e-9
        assert r["traffic"] and 0.99 < r["traffic"] / r["algorithmic_bytes_per_launch"] < 1.05
        b = d["cpu_baseline"]
        assert {"value", "unit", "cores", "kind", "sample"} <= set(b) and b["kind"] in ("port", "reference") and b["cores"] >= 1


def n_seq(d):
    return d["config"].get("sequences_in_flight", 1)


def test_reference_arm_line():
    d = _load("r01_bench_reference_arm.json")
    assert d["impl"] == "reference" and d["metric"] == "decode tok/s Llama-3-8B bs=1" and d["unit"] == "tok/s"
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    b = d["cpu_baseline"]
    assert b["value"] == d["value"] and b["kind"] in ("port", "reference") and b["cores"] >= 1 and b["sample"]


def test_same_token_at_every_shard_count():
    toks = set()
    for name in ("r01_bench_n2_ring.json", "r01_bench_n4_ring.json", "r01_bench_n8_ring.json"):
        d = _load(name)
        toks.add((d["check"]["nonce0_token_after_steps"], d["check"]["token"]))
    assert len(toks) == 1, toks


# ---- round 2: every N runs through the product transport (ShardNode / RingAdapter / ApiNode)
@pytest.mark.parametrize("name,n", [("r02_bench_n1.json", 1), ("r02_bench_n2.json", 2), ("r02_bench_n4.json", 4), ("r02_bench_n8.json", 8)])
def test_r02_our_arm_line(name, n):
    d = _load(name)
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["metric"] == "decode tok/s Llama-3-8B bs=1" and d["unit"] == "tok/s" and d["n_gpus"] == n
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "bf16" and d["data"] == "synthetic" and d["warmup"] >= 3
    ns = d["config"]["sequences_in_flight"]
    assert abs(d["value"] - ns * 1e3 / d["ms_per_step"]) / d["value"] < 1e-6
    assert d["gpu_launches"] >= d["steps"] * ns                      # one fused step kernel per (sequence, step) per shard
    e = d["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e)
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] >= 0.9 * d["value"], "the API-side rate must follow the device-timed one (round 1: it fell with N)"
    assert d["config"]["step_error"] == 0 and d["check"]["all_tokens_valid"] is True
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm" and 0.5 < r["frac"] < 1.1
    assert not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"]))
    b = d["cpu_baseline"]
    if b is not None:
        assert b["kind"] == "port" and b["cores"] >= 1 and "extrapolated" not in b["sample"].replace("nothing extrapolated", "")
    if n == 1:
        assert r["traffic"] and 0.99 < r["traffic"] / r["algorithmic_bytes_per_launch"] < 1.05
        od = d["check"]["oracle_full_depth"]
        # full depth (32 layers, the GPU arm's own weights), the oracle teacher-forced with the GPU's tokens:
        # every GPU token is the oracle's argmax or within 2 bf16 ulps of the oracle's best logit
        assert od and od["compared"] >= 8 and od["pass"] is True and od["max_regret_ulps"] <= 2.0
        assert od["identical"] >= od["compared"] - len(od["mismatch_steps"])
    else:
        assert d["ring_hop_us"]["activation_8k"] > 0


def test_r02_same_token_at_every_shard_count():
    toks = set()
    # the driver's settings (--steps 20 --warmup 5) at every N: request 0's token after 25 decode steps
    for name in ("r02_bench_n1_k20.json", "r02_bench_n2.json", "r02_bench_n4.json", "r02_bench_n8.json"):
        d = _load(name)
        toks.add((d["check"]["nonce0_token_after_steps"], d["check"]["token"]))
    assert len(toks) == 1, toks
