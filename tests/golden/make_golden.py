"""Generate tests/golden/*.npz from the CPU oracle (oracle/llama_oracle.py).

The reference holds no golden vectors for this path ("parity
unpinned"), and mlx cannot be imported here, so the fixtures pin the ORACLE: a later
change to the oracle or to make_weights() that moves any number fails
tests/test_oracle.py::test_oracle_matches_golden.  Run from the repo root:

    python tests/golden/make_golden.py

Greedy ids are only bit-exact-comparable where the argmax is decided by more than
rounding noise, so prompts are searched (seed scan) until every step's top-1/top-2 gap of
the fp32 logits exceeds MARGIN_ULPS bf16 ulps of the top logit; the gaps are stored.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.llama_oracle import LlamaOracle, OracleConfig, OracleKV, make_weights, sample_greedy  # noqa: E402

TINY = dict(hidden_size=512, num_attention_heads=4, num_key_value_heads=2, head_dim=128, intermediate_size=1024,
            vocab_size=1024, num_hidden_layers=4, rms_norm_eps=1e-5, rope_theta=500000.0, model_type="llama",
            tie_word_embeddings=False, torch_dtype="bfloat16")
# GQA group 4 like Llama-3-8B, qwen2-style bias on q/k/v, llama3 rope scaling, tied head
TINY_B = dict(hidden_size=512, num_attention_heads=8, num_key_value_heads=2, head_dim=128, intermediate_size=768,
              vocab_size=2048, num_hidden_layers=3, rms_norm_eps=1e-6, rope_theta=10000.0, model_type="qwen2",
              tie_word_embeddings=True, attention_bias=True, torch_dtype="bfloat16",
              rope_scaling=dict(rope_type="llama3", factor=8.0, low_freq_factor=1.0, high_freq_factor=4.0,
                                original_max_position_embeddings=64))
# sparse MoE FFN (mixtral): 8 experts, top-2, GQA group 2
TINY_MOE = dict(hidden_size=512, num_attention_heads=4, num_key_value_heads=2, head_dim=128, intermediate_size=768,
                vocab_size=1024, num_hidden_layers=3, rms_norm_eps=1e-5, rope_theta=1000000.0, model_type="mixtral",
                tie_word_embeddings=False, torch_dtype="bfloat16", num_local_experts=8, num_experts_per_tok=2)
MARGIN_ULPS = 3.0


def bf16_ulp(x: float) -> float:
    import math
    return 2.0 ** (math.floor(math.log2(abs(x))) - 7) if x != 0 else 0.0


def run(cfgd: dict, wseed: int, pseed: int, prompt_len: int, steps: int, f64: bool = False, force=None):
    cfg = OracleConfig.from_dict(cfgd)
    w = make_weights(cfg, wseed)
    m = LlamaOracle(cfg, w, torch.bfloat16, exact_linear=True, f64_linear=f64)
    g = np.random.Generator(np.random.PCG64([pseed, 77]))
    prompt = g.integers(0, cfg.vocab_size, size=prompt_len).astype(np.int32)
    kv = {l: OracleKV() for l in range(cfg.num_hidden_layers)}
    ids = torch.from_numpy(prompt)
    hidden_prefill, hidden_all = [], []
    toks, lps, gaps, logits_f32, logits_bf16 = [], [], [], [], []
    for step in range(steps):
        x = m.embed(ids)
        if step == 0:
            hidden_all.append(x.view(torch.int16).numpy().copy())
        for l in range(cfg.num_hidden_layers):
            x = m.apply_single_layer(l, x, kv[l])
            if step == 0:
                hidden_prefill.append(x[-1].view(torch.int16).numpy().copy())
                hidden_all.append(x.view(torch.int16).numpy().copy())
        y = m.normalize(x[-1:])
        lf = m.lm_project(y, return_fp32=True)[0]
        lb = lf.to(torch.bfloat16)
        top2 = torch.topk(lf, 2).values
        gap = float(top2[0] - top2[1])
        gaps.append(gap / max(bf16_ulp(float(top2[0])), 1e-30))
        r = sample_greedy(lb, True, 0)
        toks.append(r.token_id)
        lps.append(r.logprob)
        logits_f32.append(lf.numpy().copy())
        logits_bf16.append(lb.view(torch.int16).numpy().copy())
        ids = torch.tensor([r.token_id if force is None else int(force[step])], dtype=torch.int32)
    return dict(hidden_all=np.stack(hidden_all), prompt=prompt, tokens=np.array(toks, np.int32), logprobs=np.array(lps, np.float32),
                gap_ulps=np.array(gaps, np.float32), logits_f32=np.stack(logits_f32),
                logits_bf16=np.stack(logits_bf16), hidden_prefill=np.stack(hidden_prefill))


def search(name: str, cfgd: dict, wseed: int, prompt_len: int, steps: int, max_tries: int = 400):
    for pseed in range(max_tries):
        out = run(cfgd, wseed, pseed, prompt_len, steps)
        if float(out["gap_ulps"].min()) >= MARGIN_ULPS:
            import json
            # the oracle's own summation-order sensitivity, teacher-forced on the same tokens
            alt = run(cfgd, wseed, pseed, prompt_len, steps, f64=True, force=out["tokens"])
            a, b = torch.from_numpy(alt["logits_f32"]).double(), torch.from_numpy(out["logits_f32"]).double()
            out["noise_floor"] = ((a - b).abs().amax(dim=1) / b.abs().amax(dim=1)).float().numpy()
            np.savez_compressed(Path(__file__).parent / f"{name}.npz", config=json.dumps(cfgd), wseed=wseed,
                                pseed=pseed, steps=steps, **out)
            print(name, "pseed", pseed, "min gap (bf16 ulps)", float(out["gap_ulps"].min()), "tokens", out["tokens"][:8],
                  "noise floor max", float(out["noise_floor"].max()))
            return
    raise SystemExit(f"no margin-safe prompt found for {name}")


if __name__ == "__main__":
    torch.set_num_threads(8)
    search("tiny_llama", TINY, wseed=11, prompt_len=7, steps=16)
    search("tiny_qwen2_tied", TINY_B, wseed=23, prompt_len=70, steps=10)
    search("tiny_mixtral", TINY_MOE, wseed=31, prompt_len=21, steps=10)
