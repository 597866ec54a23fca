"""-m gpu: the affine-quantised KV cache (kv_bits "8bit" / "4bit" -- the reference API's defaults,
src/dnet/api/models.py:316,342; cache built by src/dnet/utils/model.py:505-554) against the oracle's
restatement of mlx quantize + mlx_lm's quantised attention (oracle/llama_oracle.py, PARITY UNPINNED).

Every decode path is covered: the persistent step kernel (on-the-fly quantisation of the new row + two-pass
attention with a cross-CTA (max, sum) exchange), the per-op kernels (small chunks / offload decode) and the
tensor-core prefill chunks, all through the policy API.  Steps are teacher-forced on the oracle's tokens so
a near-tie cannot cascade; tokens must agree wherever the oracle's top-1/top-2 margin exceeds 3 bf16 ulps."""
import numpy as np
import pytest
import torch

from tests.helpers import load_golden, make_runtime, oracle_weights, rel_inf, token_message
from tests.test_gpu_parity import _teacher_forced_logits

pytestmark = pytest.mark.gpu


def _oracle_run(cfgd, w, prompt, steps, bits):
    """greedy tokens of the oracle with a quantised cache + its teacher-forced fp32 logits and margins"""
    from oracle.llama_oracle import LlamaOracle, OracleConfig, OracleQuantKV, sample_greedy

    oc = OracleConfig.from_dict(cfgd)
    m = LlamaOracle(oc, w, exact_linear=True)
    kv = {l: OracleQuantKV(bits) for l in range(oc.num_hidden_layers)}
    ids = torch.tensor(list(prompt), dtype=torch.int32)
    toks, logits, gaps = [], [], []
    for _ in range(steps):
        x = m.embed(ids)
        for l in range(oc.num_hidden_layers):
            x = m.apply_single_layer(l, x, kv[l]).to(torch.bfloat16)
        lf = m.lm_project(m.normalize(x[-1:]), return_fp32=True)[0]
        top2 = torch.topk(lf, 2).values
        gaps.append(float(top2[0] - top2[1]) / max(float(top2[0].abs()) * 2.0 ** -8, 1e-30))
        t = sample_greedy(lf.to(torch.bfloat16), False, 0).token_id
        toks.append(t)
        logits.append(lf.double())
        ids = torch.tensor([t], dtype=torch.int32)
    return toks, logits, gaps


@pytest.mark.parametrize("bits", [8, 4])
@pytest.mark.parametrize("name,mk", [("tiny_llama", True), ("tiny_llama", False), ("tiny_qwen2_tied", True)])
def test_quantised_kv_decode_and_prefill_against_oracle(cuda_lib, name, mk, bits):
    g = load_golden(name)
    cfgd = g["config"]
    w = oracle_weights(cfgd, g["wseed"])
    steps = int(g["steps"])
    prompt = g["prompt"].tolist()
    toks, ref, gaps = _oracle_run(cfgd, w, prompt, steps, bits)
    # yardstick: the oracle's own summation-order sensitivity WITH this cache (fp32- vs float64-accumulated linears)
    alt = _teacher_forced_logits(cfgd, w, prompt, toks, f64=True, kv_bits=bits)
    floor = max(rel_inf(ref[i], alt[i]) for i in range(steps))
    tol = max(2e-3, 4.0 * floor)
    rt = make_runtime(cfgd, w, range(cfgd["num_hidden_layers"]), megakernel=mk, cuda_graphs=False, kv_bits=f"{bits}bit")
    try:
        assert rt.model.kv_bits == bits and rt.kv_cache_config.mode == f"{bits}bit"
        ids = prompt
        worst, bad = 0.0, []
        for step in range(steps):
            rt.policy.process(token_message(rt, "q", ids, req_logprobs=True))
            res = rt.activation_send_queue.get_nowait()
            ns = rt._kv_by_nonce["q"]
            f32, _ = rt.model.head_logits(ns.x_view(len(ids)))
            torch.cuda.synchronize()
            worst = max(worst, rel_inf(f32.cpu(), ref[step]))
            if res.token_id != toks[step] and gaps[step] >= 3.0:
                bad.append((step, res.token_id, toks[step], gaps[step]))
            ids = [toks[step]]
        assert not bad, f"greedy ids differ from the oracle at margin-safe steps: {bad}"
        assert worst <= tol, f"logits rel err {worst:.3e} > {tol:.3e} (oracle order-sensitivity {floor:.3e})"
        assert cuda_lib.dn_step_error(rt.model._h, rt.compute_stream_ptr) == 0
        print(f"{name} kv{bits} mk={mk}: worst logits rel err {worst:.3e} (tol {tol:.3e})")
    finally:
        rt.unload_model_core()


@pytest.mark.parametrize("bits", [8, 4])
def test_quantised_kv_long_context_splits_and_page_boundaries(cuda_lib, bits):
    """Contexts long enough that a head is split over several CTAs in the step kernel (the (max, sum) exchange),
    crossing page boundaries, prefilled in tensor-core chunks: the step kernel and the per-op path must both
    match the oracle, and each other within the same envelope."""
    g = load_golden("tiny_llama")
    cfgd = g["config"]
    w = oracle_weights(cfgd, g["wseed"])
    prompt = np.random.Generator(np.random.PCG64(5)).integers(0, cfgd["vocab_size"], size=700).tolist()
    steps = 4
    toks, ref, gaps = _oracle_run(cfgd, w, prompt, steps, bits)
    alt = _teacher_forced_logits(cfgd, w, prompt, toks, f64=True, kv_bits=bits)
    tol = max(2e-3, 4.0 * max(rel_inf(ref[i], alt[i]) for i in range(steps)))
    for mk in (True, False):
        rt = make_runtime(cfgd, w, range(cfgd["num_hidden_layers"]), megakernel=mk, cuda_graphs=False, kv_bits=f"{bits}bit",
                          max_tokens=1024)
        try:
            cuda_lib.dn_set_option(b"attn_chunk", 32)
            ids = prompt
            for step in range(steps):
                rt.policy.process(token_message(rt, "long", ids))
                res = rt.activation_send_queue.get_nowait()
                ns = rt._kv_by_nonce["long"]
                f32, _ = rt.model.head_logits(ns.x_view(len(ids)))
                torch.cuda.synchronize()
                r = rel_inf(f32.cpu(), ref[step])
                assert r <= tol, f"mk={mk} step {step}: logits rel err {r:.3e} > {tol:.3e}"
                assert res.token_id == toks[step] or gaps[step] < 3.0
                ids = [toks[step]]
            assert cuda_lib.dn_step_error(rt.model._h, rt.compute_stream_ptr) == 0
        finally:
            rt.unload_model_core()


def test_quantised_kv_pool_is_smaller(cuda_lib):
    """bytes per cached token and layer: 4096 (bf16) -> 2176 (8 bit) -> 1152 (4 bit) for 8 KV heads"""
    from dnet_b200 import _cabi
    import ctypes as C

    assert C.sizeof(_cabi.ModelCfg) == 17 * 4
    unit = {8: 64 * 128 + 64 * 8, 4: 64 * 64 + 64 * 8}
    assert unit[8] * 2 * 8 / 64 == 2176 and unit[4] * 2 * 8 / 64 == 1152
