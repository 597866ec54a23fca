"""Bisect the per-op path against the oracle inside one layer: feed layer 0 the oracle's input for the first prompt
token (T=1, position 0) and compare q (after RoPE), attention output, h and the SwiGLU output with the oracle's."""
import ctypes as C
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from dnet_b200 import _cabi  # noqa: E402
from oracle.llama_oracle import LlamaOracle, OracleConfig, OracleKV  # noqa: E402
from tests.helpers import load_golden, make_runtime, oracle_weights  # noqa: E402


def ulps(a, b):
    def key(t):
        i = t.view(torch.int16).to(torch.int32)
        return torch.where(i < 0, -(i & 0x7FFF), i)
    return (key(a) - key(b)).abs()


def main(name="tiny_llama", npos=3):
    g = load_golden(name)
    cfgd = g["config"]
    w = oracle_weights(cfgd, g["wseed"])
    oc = OracleConfig.from_dict(cfgd)
    m = LlamaOracle(oc, w, exact_linear=True)
    lib = _cabi.load()
    rt = make_runtime(cfgd, w, range(oc.num_hidden_layers), megakernel=False, cuda_graphs=False)
    try:
        from tests.helpers import token_message
        msg = token_message(rt, "probe", g["prompt"].tolist())
        to_bind = rt.policy._bind_layer_weights(list(range(oc.num_hidden_layers)), msg)
        torch.cuda.synchronize()
        rt.model.load_weights(list(to_bind.items()))
        H, F, qd = oc.hidden_size, oc.intermediate_size, oc.num_attention_heads * oc.head_dim
        x_all = m.embed(torch.tensor(g["prompt"], dtype=torch.int32))
        for layer in range(2):
            kvo = OracleKV()
            ns = rt.get_or_make_kv(f"b{layer}")
            for pos in range(npos):
                x = x_all[pos:pos + 1]
                p = f"model.layers.{layer}."
                # oracle intermediates (same code as LlamaOracle.apply_single_layer)
                xn = m.rms_norm(x, p + "input_layernorm.weight")
                q = m.linear(xn, p + "self_attn.q_proj.weight").view(1, oc.num_attention_heads, oc.head_dim).transpose(0, 1)
                k = m.linear(xn, p + "self_attn.k_proj.weight").view(1, oc.num_key_value_heads, oc.head_dim).transpose(0, 1)
                v = m.linear(xn, p + "self_attn.v_proj.weight").view(1, oc.num_key_value_heads, oc.head_dim).transpose(0, 1)
                off = kvo.offset
                q, k = m.rope(q, off), m.rope(k, off)
                kk, vv = kvo.update_and_fetch(k, v)
                a = m.sdpa(q, kk, vv, off).transpose(0, 1).reshape(1, -1)
                r = m.linear(a, p + "self_attn.o_proj.weight")
                h = m.T(x.float() + r.float())
                hn = m.rms_norm(h, p + "post_attention_layernorm.weight")
                gg = m.linear(hn, p + "mlp.gate_proj.weight").float()
                uu = m.linear(hn, p + "mlp.up_proj.weight").float()
                sg = m.T(torch.sigmoid(gg)).float()
                act = m.T(m.T(gg * sg).float() * uu)
                out = m.T(h.float() + m.linear(act, p + "mlp.down_proj.weight").float())
                # GPU
                xin = x.cuda().unsqueeze(0).contiguous()
                got = rt.model.apply_single_layer(layer, xin, ns.kv)[0].cpu()
                ns.kv.advance(1, rt.compute_stream_ptr)
                torch.cuda.synchronize()
                bufs = {}
                for which, (nm, n) in enumerate((("q_rope", qd), ("attn", qd), ("h", H), ("act", F))):
                    t = torch.empty(n, dtype=torch.bfloat16)
                    _cabi.check(lib.dn_debug_scratch(rt.model._h, which, t.data_ptr(), n * 2, rt.compute_stream_ptr))
                    bufs[nm] = t
                refs = {"q_rope": q.transpose(0, 1).reshape(-1), "attn": a.reshape(-1), "h": h.reshape(-1), "act": act.reshape(-1)}
                line = f"layer {layer} pos {pos}: "
                for nm in ("q_rope", "attn", "h", "act"):
                    d = ulps(bufs[nm], refs[nm].to(torch.bfloat16))
                    line += f"{nm} {int((d > 0).sum())}/{d.numel()} (max {int(d.max())} ulp)  "
                d = ulps(got.reshape(-1), out.reshape(-1))
                line += f"out {int((d > 0).sum())}/{d.numel()} (max {int(d.max())} ulp)"
                print(line)
                x_all = x_all.clone()
            # next layer input = oracle's output of this layer for all positions
            kvt = OracleKV()
            x_all = m.apply_single_layer(layer, x_all, kvt)
    finally:
        rt.unload_model_core()


if __name__ == "__main__":
    main(*sys.argv[1:2])
