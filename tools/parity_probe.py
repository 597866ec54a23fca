"""Where does the GPU's end-to-end deviation from the oracle come from?  Prints err(GPU, f64 oracle) per step for
the tiny fixture with the prompt fed (a) as one message (small-chunk prefill kernels) and (b) token by token
(decode kernels only), for the step kernel and the per-op path."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from tests.helpers import load_golden, make_runtime, oracle_weights, rel_inf, token_message  # noqa: E402
from tests.test_gpu_parity import _teacher_forced_logits  # noqa: E402


def main(name="tiny_llama"):
    g = load_golden(name)
    cfgd = g["config"]
    w = oracle_weights(cfgd, g["wseed"])
    steps = int(g["steps"])
    toks = [int(t) for t in g["tokens"][:steps]]
    prompt = g["prompt"].tolist()
    ref64 = _teacher_forced_logits(cfgd, w, prompt, toks, f64=True)
    ref32 = _teacher_forced_logits(cfgd, w, prompt, toks, f64=False)
    print("yardstick err(fp32,f64):", " ".join(f"{rel_inf(ref32[i], ref64[i]):.1e}" for i in range(steps)))
    for mk in (True, False):
        for mode in ("one message", "token by token"):
            rt = make_runtime(cfgd, w, range(cfgd["num_hidden_layers"]), megakernel=mk, cuda_graphs=False)
            try:
                errs = []
                ids = prompt
                for step in range(steps):
                    feed = [ids] if (mode == "one message" or len(ids) == 1) else [[t] for t in ids]
                    for chunk in feed:
                        rt.policy.process(token_message(rt, "p", chunk))
                        rt.activation_send_queue.get_nowait()
                    ns = rt._kv_by_nonce["p"]
                    f32, _ = rt.model.head_logits(ns.x_view(len(feed[-1])))
                    torch.cuda.synchronize()
                    errs.append(rel_inf(f32.cpu(), ref64[step]))
                    ids = [toks[step]]
                print(f"step_kernel={mk} prompt {mode:>15}: " + " ".join(f"{e:.1e}" for e in errs))
            finally:
                rt.unload_model_core()


if __name__ == "__main__":
    main(*sys.argv[1:])
