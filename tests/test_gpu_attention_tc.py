"""-m gpu: the long-context attention phase of the step kernel on mma.sync with cp.async-staged K/V tiles
(mk_attention_tc, dn_megakernel.cuh) against the oracle, and against the CUDA-core attention phase it replaces.
Option attn_tc_min is lowered so the path engages at test-sized contexts: tiles with ragged tails, page
boundaries, multi-CTA splits, GQA groups of 2 and 4."""
import numpy as np
import pytest
import torch

from tests.helpers import load_golden, make_runtime, oracle_weights, rel_inf, token_message
from tests.test_gpu_parity import _teacher_forced_logits

pytestmark = pytest.mark.gpu


def _oracle_tokens(cfgd, w, prompt, steps):
    from oracle.llama_oracle import LlamaOracle, OracleConfig, OracleKV, sample_greedy

    oc = OracleConfig.from_dict(cfgd)
    m = LlamaOracle(oc, w, exact_linear=True)
    kv = {l: OracleKV() for l in range(oc.num_hidden_layers)}
    ids = torch.tensor(list(prompt), dtype=torch.int32)
    toks, logits, gaps = [], [], []
    for _ in range(steps):
        x = m.embed(ids)
        for l in range(oc.num_hidden_layers):
            x = m.apply_single_layer(l, x, kv[l]).to(torch.bfloat16)
        lf = m.lm_project(m.normalize(x[-1:]), return_fp32=True)[0]
        top2 = torch.topk(lf, 2).values
        gaps.append(float(top2[0] - top2[1]) / max(float(top2[0].abs()) * 2.0 ** -8, 1e-30))
        toks.append(sample_greedy(lf.to(torch.bfloat16), False, 0).token_id)
        logits.append(lf.double())
        ids = torch.tensor([toks[-1]], dtype=torch.int32)
    return toks, logits, gaps


@pytest.mark.parametrize("name,plen", [("tiny_llama", 37), ("tiny_llama", 1500), ("tiny_qwen2_tied", 130), ("tiny_qwen2_tied", 700),
                                       ("tiny_llama", 2100), ("tiny_llama", 8200)])
def test_tensor_core_decode_attention_against_oracle_and_cuda_core_path(cuda_lib, name, plen):
    g = load_golden(name)
    cfgd = g["config"]
    w = oracle_weights(cfgd, g["wseed"])
    prompt = np.random.Generator(np.random.PCG64(plen)).integers(0, cfgd["vocab_size"], size=plen).tolist()
    steps = 6
    toks, ref, gaps = _oracle_tokens(cfgd, w, prompt, steps)
    alt = _teacher_forced_logits(cfgd, w, prompt, toks, f64=True)
    tol = max(2e-3, 4.0 * max(rel_inf(ref[i], alt[i]) for i in range(steps)))
    got = {}
    try:
        for tc in (1, 0):
            cuda_lib.dn_set_option(b"attn_tc", tc)
            cuda_lib.dn_set_option(b"attn_tc_min", 16)
            rt = make_runtime(cfgd, w, range(cfgd["num_hidden_layers"]), megakernel=True, cuda_graphs=False, max_tokens=max(2048, plen + 64))
            try:
                ids, out = prompt, []
                for step in range(steps):
                    rt.policy.process(token_message(rt, "a", ids))
                    res = rt.activation_send_queue.get_nowait()
                    ns = rt._kv_by_nonce["a"]
                    f32, _ = rt.model.head_logits(ns.x_view(len(ids)))
                    torch.cuda.synchronize()
                    out.append(f32.cpu().double())
                    if step > 0:        # step 0 is the prefill (prefill kernels); decode steps run the step kernel
                        assert res.token_id == toks[step] or gaps[step] < 3.0, f"tc={tc} step {step}"
                    ids = [toks[step]]
                assert cuda_lib.dn_step_error(rt.model._h, rt.compute_stream_ptr) == 0
                got[tc] = out
            finally:
                rt.unload_model_core()
        for tc in (1, 0):
            worst = max(rel_inf(got[tc][i], ref[i]) for i in range(steps))
            assert worst <= tol, f"{name} ctx {plen} attn_tc={tc}: logits rel err {worst:.3e} > {tol:.3e}"
        between = max(rel_inf(got[1][i], got[0][i]) for i in range(steps))
        assert between <= tol
        print(f"{name} ctx {plen}: vs oracle tc {max(rel_inf(got[1][i], ref[i]) for i in range(steps)):.2e} "
              f"cuda-core {max(rel_inf(got[0][i], ref[i]) for i in range(steps)):.2e}; between paths {between:.2e} (tol {tol:.2e})")
    finally:
        cuda_lib.dn_set_option(b"attn_tc", 1)
        cuda_lib.dn_set_option(b"attn_tc_min", 12288)
