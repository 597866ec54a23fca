"""-m gpu: sparse MoE FFN layers (mixtral; BASELINE.json configs[4]) through the reference-facing policy / model API
against the CPU oracle's restatement (oracle/llama_oracle.py moe_block) and the committed fixture
tests/golden/tiny_mixtral.npz (8 experts, top-2, GQA group 2, 3 layers).  Tolerances as in test_gpu_parity.py:
per operator on identical inputs the output agrees up to the last bf16 rounding, greedy ids are bit-exact."""
import ctypes as C

import pytest
import torch

from tests.helpers import load_golden, make_runtime, oracle_weights, rel_inf, ring_generate, token_message
from tests.test_gpu_parity import LOGIT_TOL, _bf16, e2e_tol, ulp_diff

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def moe(cuda_lib):
    g = load_golden("tiny_mixtral")
    return g, oracle_weights(g["config"], g["wseed"])


def test_moe_layers_against_oracle_on_identical_inputs(moe):
    """Every layer is fed the oracle's input for that layer, all 21 prompt positions at once (the tensor-core
    prefill path for the attention half, the expert FFN token by token) and in T = 4 + 2 + 1 chunks (the GEMV path)."""
    g, w = moe
    cfgd = g["config"]
    L = cfgd["num_hidden_layers"]
    rt = make_runtime(cfgd, w, range(L), cuda_graphs=False, megakernel=False)
    try:
        assert rt.use_megakernel is False and rt.model.n_experts == 8
        m = rt.model
        msg = token_message(rt, "probe", g["prompt"].tolist())
        to_bind = rt.policy._bind_layer_weights(list(range(L)), msg)
        torch.cuda.synchronize()
        m.load_weights(list(to_bind.items()))
        for l in range(L):
            ref = _bf16(g["hidden_all"][l + 1])
            xin_all = _bf16(g["hidden_all"][l]).cuda()
            for chunks in ([21], [4, 4, 4, 4, 2, 2, 1]):
                ns = rt.get_or_make_kv(f"probe{l}-{len(chunks)}")
                outs, p0 = [], 0
                for n in chunks:
                    x = xin_all[p0:p0 + n].clone().unsqueeze(0)       # the layer runs in place
                    outs.append(m.apply_single_layer(l, x, ns.kv)[0].clone())
                    ns.kv.advance(n)          # same (default) stream as apply_single_layer
                    p0 += n
                torch.cuda.synchronize()
                got = torch.cat(outs).cpu()
                frac = float((ulp_diff(got, ref) > 0).float().mean())
                worst = float((got.float() - ref.float()).abs().max() / ref.float().abs().max())
                assert frac < 0.03 and worst <= 2 * 2.0 ** -8, f"layer {l} chunks {chunks}: {frac:.4f} mismatching, worst {worst:.2e}"
    finally:
        rt.unload_model_core()


@pytest.mark.parametrize("graphs", [True, False])
def test_moe_greedy_generation_matches_golden(moe, graphs):
    g, w = moe
    cfgd = g["config"]
    rt = make_runtime(cfgd, w, range(cfgd["num_hidden_layers"]), cuda_graphs=graphs, megakernel=True)   # megakernel asked for, refused
    try:
        assert rt.use_megakernel is False, "MoE models must fall back to the per-op path"
        out = ring_generate([rt], "moe", g["prompt"].tolist(), int(g["steps"]))
        assert [t for t, _, _ in out] == g["tokens"].tolist()
        f32, _ = rt.model.head_logits(rt._kv_by_nonce["moe"].x1)
        torch.cuda.synchronize()
        assert rel_inf(f32.cpu(), torch.from_numpy(g["logits_f32"][int(g["steps"]) - 1])) <= e2e_tol(g)
        lps = torch.tensor([lp for _, lp, _ in out])
        assert float((lps - torch.from_numpy(g["logprobs"])).abs().max()) <= 0.13      # bf16 logprobs near ln(V)
    finally:
        rt.unload_model_core()


def test_moe_two_shards_equal_one(moe):
    g, w = moe
    cfgd = g["config"]
    L = cfgd["num_hidden_layers"]
    a = make_runtime(cfgd, w, range(0, 1), shard_id="a")
    b = make_runtime(cfgd, w, range(1, L), shard_id="b")
    try:
        out = ring_generate([a, b], "split", g["prompt"].tolist(), int(g["steps"]))
        assert [t for t, _, _ in out] == g["tokens"].tolist()
    finally:
        a.unload_model_core()
        b.unload_model_core()


def test_step_kernel_refuses_moe(moe, cuda_lib):
    from dnet_b200 import _cabi

    g, w = moe
    cfgd = g["config"]
    L = cfgd["num_hidden_layers"]
    rt = make_runtime(cfgd, w, range(L))
    try:
        ring_generate([rt], "r", g["prompt"].tolist(), 1)
        ns = rt._kv_by_nonce["r"]
        arr = (C.c_int32 * L)(*range(L))
        rc = cuda_lib.dn_shard_step(rt.model._h, arr, L, ns.x1.data_ptr(), ns.kv._h, 1, 1, None, None, None, 1,
                                    rt.compute_stream_ptr)
        assert rc < 0 and "MoE" in _cabi.last_error()
    finally:
        rt.unload_model_core()
