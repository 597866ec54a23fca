"""Measured argument for the wire-dtype restriction.

The reference's default wire dtype is fp16 (src/dnet/config.py:108-111) and the policy casts the residual
stream to it after every layer (shard/policies/fit_in_memory.py:105-107).  With bf16 weights MLX promotes
fp16 x bf16 to fp32, so under the DEFAULT settings every layer of the reference computes in fp32 (norms,
projections, RoPE, attention, SwiGLU, residuals) and rounds ONCE, to fp16, at the layer boundary; the KV
cache holds fp32.  That is a different numerical pipeline from "bf16 activations, bf16 wire", not a cast:
this script runs the oracle both ways on the same weights / prompt and reports how far the logits are apart.
CPU only (oracle), writes one JSON line."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle.llama_oracle import LlamaOracle, OracleConfig, OracleKV, make_weights  # noqa: E402
from tests.helpers import load_golden, oracle_weights, rel_inf  # noqa: E402


def run(cfgd, w, prompt, steps, compute_dtype, wire):
    oc = OracleConfig.from_dict(cfgd)
    m = LlamaOracle(oc, w, dtype=compute_dtype, exact_linear=True)
    kv = {l: OracleKV() for l in range(oc.num_hidden_layers)}
    ids = torch.tensor(prompt, dtype=torch.int32)
    out, toks = [], []
    for _ in range(steps):
        x = m.embed(ids).to(wire)
        for l in range(oc.num_hidden_layers):
            x = m.apply_single_layer(l, x, kv[l]).to(wire)        # the policy's per-layer cast to the wire dtype
        lf = m.lm_project(m.normalize(x[-1:]), return_fp32=True)[0]
        out.append(lf.double())
        t = int(torch.argmax(lf if compute_dtype == torch.float32 else lf.to(torch.bfloat16).float()))
        toks.append(t)
        ids = torch.tensor([t], dtype=torch.int32)
    return out, toks


def main():
    res = {}
    for name in ("tiny_llama", "tiny_qwen2_tied"):
        g = load_golden(name)
        w = oracle_weights(g["config"], g["wseed"])
        steps = int(g["steps"])
        a, ta = run(g["config"], w, g["prompt"].tolist(), steps, torch.bfloat16, torch.bfloat16)   # what dnet_b200 runs
        b, tb = run(g["config"], w, g["prompt"].tolist(), steps, torch.float32, torch.float16)     # the reference's default
        c, tc = run(g["config"], w, g["prompt"].tolist(), steps, torch.float32, torch.float32)     # exact fp32 pipeline
        same = next((i for i in range(steps) if ta[i] != tb[i]), steps)
        n = max(1, same)
        res[name] = {"steps": steps, "tokens_equal_prefix": same,
                     "logits_rel_err_bf16_vs_fp16wire": max(rel_inf(a[i], b[i]) for i in range(n)),
                     "logits_rel_err_fp16wire_vs_fp32": max(rel_inf(b[i], c[i]) for i in range(min(n, next((i for i in range(steps) if tb[i] != tc[i]), steps)) or 1)),
                     "noise_floor_bf16_pipeline": float(g["noise_floor"].max())}
    print(json.dumps({"what": "max over steps of max|a-b|/max|b| on the fp32 last-position logits, same weights and prompt; "
                              "bf16 = bf16 activations + bf16 wire (dnet_b200); fp16wire = fp32 compute + fp16 cast per layer "
                              "(what the reference's defaults do with bf16 weights under MLX type promotion)", **res}))


if __name__ == "__main__":
    main()
