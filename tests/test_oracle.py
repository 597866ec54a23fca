"""The oracle is pinned three ways (no GPU needed):
  1. its bf16 Linear against an exact float64 product (one correct rounding);
  2. its structure against an independent implementation (HF transformers, fp32);
  3. its numbers against the committed golden fixtures."""
import numpy as np
import pytest
import torch

from oracle.llama_oracle import (LlamaOracle, OracleConfig, OracleKV, OracleShard, greedy_generate, make_weights,
                                 rope_inv_freq, sample_greedy)
from tests.helpers import load_golden


def test_bf16_linear_is_single_rounded_fp32_accumulate():
    torch.manual_seed(0)
    W = (torch.randn(640, 1024) * 0.02).to(torch.bfloat16)
    x = torch.randn(3, 1024).to(torch.bfloat16)
    cfg = OracleConfig(1024, 8, 8, 128, 640, 16, 1)
    o = LlamaOracle(cfg, {"w": W})
    fast = o.linear(x, "w").float()
    exact = (x.double() @ W.double().T).float().to(torch.bfloat16).float()
    # identical except where the fp32 accumulation order crosses a bf16 rounding boundary
    assert (fast != exact).float().mean() < 2e-3
    assert torch.allclose(fast, exact, rtol=2 ** -7, atol=1e-6)


def test_structure_matches_hf_llama_fp32():
    transformers = pytest.importorskip("transformers")
    cfgd = dict(hidden_size=256, num_attention_heads=2, num_key_value_heads=1, head_dim=128, intermediate_size=512,
                vocab_size=320, num_hidden_layers=3, rms_norm_eps=1e-5, rope_theta=500000.0)
    cfg = OracleConfig.from_dict(cfgd)
    w = make_weights(cfg, 7, dtype=torch.float32)
    hc = transformers.LlamaConfig(**cfgd, max_position_embeddings=512, attention_bias=False, mlp_bias=False,
                                  tie_word_embeddings=False, attn_implementation="eager")
    m = transformers.LlamaForCausalLM(hc).eval().to(torch.float32)
    m.load_state_dict({k: w[k] for k in m.state_dict()})
    ids = torch.randint(0, 320, (1, 11), generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref = m(ids).logits[0]
    o = LlamaOracle(cfg, w, torch.float32)
    kv = {l: OracleKV() for l in range(3)}
    outs = []
    x = o.embed(ids[0, :6])                      # prefill 6, then decode 5: exercises offset>0
    for l in range(3):
        x = o.apply_single_layer(l, x, kv[l])
    outs.append(o.lm_project(o.normalize(x)))
    for t in range(6, 11):
        x = o.embed(ids[0, t:t + 1])
        for l in range(3):
            x = o.apply_single_layer(l, x, kv[l])
        outs.append(o.lm_project(o.normalize(x)))
    got = torch.cat(outs)
    assert float((got - ref).abs().max() / ref.abs().max()) < 1e-5


def test_moe_structure_matches_hf_mixtral_fp32():
    """The restated sparse-MoE block (router -> top-k -> softmax over the selected logits -> SwiGLU experts ->
    weighted sum) against an independent implementation of the same architecture."""
    transformers = pytest.importorskip("transformers")
    if not hasattr(transformers, "MixtralForCausalLM"):
        pytest.skip("this transformers build has no Mixtral")
    cfgd = dict(hidden_size=256, num_attention_heads=2, num_key_value_heads=1, head_dim=128, intermediate_size=384,
                vocab_size=320, num_hidden_layers=2, rms_norm_eps=1e-5, rope_theta=1000000.0, num_local_experts=8,
                num_experts_per_tok=2)
    cfg = OracleConfig.from_dict(cfgd)
    w = make_weights(cfg, 9, dtype=torch.float32)
    hc = transformers.MixtralConfig(**cfgd, max_position_embeddings=512, tie_word_embeddings=False, sliding_window=None,
                                    attn_implementation="eager", router_jitter_noise=0.0)
    m = transformers.MixtralForCausalLM(hc).eval().to(torch.float32)
    sd = m.state_dict()
    if any(".experts.gate_up_proj" in k for k in sd):       # newer transformers keep the experts stacked
        mapped = {}
        for k in sd:
            if k.endswith("mlp.experts.gate_up_proj") or k.endswith("block_sparse_moe.experts.gate_up_proj"):
                pre = k.rsplit("experts.", 1)[0].replace(".mlp.", ".block_sparse_moe.")
                g = torch.stack([w[f"{pre}experts.{e}.w1.weight"] for e in range(8)])
                u = torch.stack([w[f"{pre}experts.{e}.w3.weight"] for e in range(8)])
                t = torch.cat([g, u], dim=1)
                mapped[k] = t if t.shape == sd[k].shape else t.transpose(1, 2).contiguous()
            elif k.endswith("experts.down_proj"):
                pre = k.rsplit("experts.", 1)[0].replace(".mlp.", ".block_sparse_moe.")
                t = torch.stack([w[f"{pre}experts.{e}.w2.weight"] for e in range(8)])
                mapped[k] = t if t.shape == sd[k].shape else t.transpose(1, 2).contiguous()
            elif ".mlp.gate.weight" in k:
                mapped[k] = w[k.replace(".mlp.gate.", ".block_sparse_moe.gate.")]
            else:
                mapped[k] = w[k]
        m.load_state_dict(mapped)
    else:
        m.load_state_dict({k: w[k] for k in sd})
    ids = torch.randint(0, 320, (1, 9), generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        ref = m(ids).logits[0]
    o = LlamaOracle(cfg, w, torch.float32)
    kv = {l: OracleKV() for l in range(2)}
    x = o.embed(ids[0, :5])
    for l in range(2):
        x = o.apply_single_layer(l, x, kv[l])
    outs = [o.lm_project(o.normalize(x))]
    for t in range(5, 9):
        x = o.embed(ids[0, t:t + 1])
        for l in range(2):
            x = o.apply_single_layer(l, x, kv[l])
        outs.append(o.lm_project(o.normalize(x)))
    got = torch.cat(outs)
    assert float((got - ref).abs().max() / ref.abs().max()) < 1e-5


def test_llama3_rope_scaling_matches_hf():
    transformers = pytest.importorskip("transformers")
    from transformers.modeling_rope_utils import ROPE_INIT_FUNCTIONS
    rs = dict(rope_type="llama3", factor=8.0, low_freq_factor=1.0, high_freq_factor=4.0,
              original_max_position_embeddings=8192)
    cfgd = dict(hidden_size=512, num_attention_heads=4, num_key_value_heads=2, head_dim=128, intermediate_size=512,
                vocab_size=64, num_hidden_layers=1, rope_theta=500000.0, rope_scaling=rs)
    mine = rope_inv_freq(OracleConfig.from_dict(cfgd))
    hc = transformers.LlamaConfig(**{k: v for k, v in cfgd.items() if k != "rope_scaling"}, rope_scaling=rs,
                                  max_position_embeddings=131072)
    try:
        ref, _ = ROPE_INIT_FUNCTIONS["llama3"](hc, "cpu")
    except Exception as e:  # transformers API drift: not a failure of the oracle
        pytest.skip(f"HF rope init unavailable: {e}")
    assert torch.allclose(mine, ref.float(), rtol=1e-6, atol=0)


@pytest.mark.parametrize("name", ["tiny_llama", "tiny_qwen2_tied", "tiny_mixtral"])
def test_oracle_matches_golden(name):
    g = load_golden(name)
    cfg = OracleConfig.from_dict(g["config"])
    w = make_weights(cfg, g["wseed"])
    res = greedy_generate(cfg, w, g["prompt"].tolist(), g["steps"], exact_linear=True)
    assert [r.token_id for r in res] == g["tokens"].tolist()
    assert np.array_equal(np.array([r.logprob for r in res], np.float32), g["logprobs"])
    assert float(g["gap_ulps"].min()) >= 3.0


def test_two_shard_split_is_bit_identical_to_one():
    g = load_golden("tiny_llama")
    cfg = OracleConfig.from_dict(g["config"])
    w = make_weights(cfg, g["wseed"])
    one = greedy_generate(cfg, w, g["prompt"].tolist(), 6, exact_linear=True)
    two = greedy_generate(cfg, w, g["prompt"].tolist(), 6, splits=[[0, 1], [2, 3]], exact_linear=True)
    assert [(a.token_id, a.logprob) for a in one] == [(b.token_id, b.logprob) for b in two]


def test_sampler_semantics():
    v = torch.tensor([0.5, 2.0, 2.0, -1.0]).to(torch.bfloat16)
    r = sample_greedy(v, True, 3)
    assert r.token_id == 1                      # first maximal index
    lse = torch.logsumexp(v.float(), -1).to(torch.bfloat16)
    assert r.logprob == float((v.float()[1] - lse.float()).to(torch.bfloat16))
    assert list(r.top_logprobs)[0] in (1, 2) and len(r.top_logprobs) == 3


def test_oracle_shard_routes_like_fit_policy():
    g = load_golden("tiny_llama")
    cfg = OracleConfig.from_dict(g["config"])
    w = make_weights(cfg, g["wseed"])
    sh = OracleShard(LlamaOracle(cfg, w, exact_linear=True), [0, 1])
    kind, x, last = sh.process("n", torch.tensor(g["prompt"]), "tokens", -1)
    assert kind == "activation" and last == 1 and x.shape == (len(g["prompt"]), cfg.hidden_size)
    with pytest.raises(RuntimeError):
        sh.process("n", x, "bfloat16", 2)       # layer 3 is not hosted here


@pytest.mark.parametrize("bits", [4, 8])
def test_affine_kv_quantiser_properties(bits):
    """N4 groundwork (restated mx.quantize, parity unpinned): codes in range, the large-magnitude bound of
    every group is reproduced exactly (up to the bf16 storage of scale / bias), error within one step (the
    snapped scale can leave the far bound just outside the code range, where it clips), constant groups
    survive, 8 bits beat 4 bits."""
    from oracle.llama_oracle import mlx_affine_dequantize, mlx_affine_quantize
    g = torch.Generator().manual_seed(bits)
    w = (torch.randn(3, 5, 128, generator=g) * 0.7).to(torch.bfloat16)
    w[0, 0, :64] = 0.25                                            # constant group
    codes, sc, bi = mlx_affine_quantize(w, bits)
    assert codes.dtype == torch.uint8 and int(codes.max()) <= (1 << bits) - 1
    assert sc.shape == (3, 5, 2) and bi.shape == (3, 5, 2) and sc.dtype == torch.bfloat16
    d = mlx_affine_dequantize(codes, sc, bi)
    err = (d - w.float()).abs().reshape(3, 5, 2, 64)
    bound = sc.float().abs().unsqueeze(-1) * 1.0 + 2.0 ** -7 * w.float().abs().amax() + 1e-6
    assert bool((err <= bound).all())
    assert float(err[0, 0, 0].max()) <= 2.0 ** -8 * 0.25 + 1e-6
    if bits == 8:
        c4, s4, b4 = mlx_affine_quantize(w, 4)
        assert (mlx_affine_dequantize(c4, s4, b4) - w.float()).abs().mean() > err.mean()


def test_quantised_kv_cache_tracks_the_float_cache():
    """The oracle with an 8-bit KV cache stays close to the bf16-cache oracle, 4-bit is further away,
    offsets advance identically (mlx_lm QuantizedKVCache semantics)."""
    from oracle.llama_oracle import LlamaOracle, OracleConfig, OracleKV, OracleQuantKV, make_weights
    cfg = OracleConfig.from_dict(dict(hidden_size=256, num_attention_heads=2, num_key_value_heads=1, head_dim=128, intermediate_size=512,
                                      vocab_size=97, num_hidden_layers=2, rms_norm_eps=1e-5, rope_theta=10000.0, model_type="llama",
                                      tie_word_embeddings=False))
    w = make_weights(cfg, 11)
    orc = LlamaOracle(cfg, w)
    ids = torch.tensor([5, 17, 3, 88, 41, 2, 9], dtype=torch.int32)

    def run(mk):
        kv = {l: mk() for l in range(2)}
        x = orc.embed(ids)
        for l in range(2):
            x = orc.apply_single_layer(l, x, kv[l])
        x2 = orc.embed(torch.tensor([7], dtype=torch.int32))
        for l in range(2):
            x2 = orc.apply_single_layer(l, x2, kv[l])
        assert kv[0].offset == 8
        return orc.lm_project(orc.normalize(x2), return_fp32=True)[0]

    ref = run(OracleKV)
    e8 = float((run(lambda: OracleQuantKV(8)) - ref).abs().max() / ref.abs().max())
    e4 = float((run(lambda: OracleQuantKV(4)) - ref).abs().max() / ref.abs().max())
    assert e8 < 0.05 and e8 < e4
