"""Sampler (reference src/dnet/core/decoding/sampler.py:8-65).

temperature == 0 (the benchmark / parity setting) never reaches this class on the CUDA
policies: argmax + logsumexp are fused into the lm_head kernel (dn_head_sample_greedy).
This class covers the remaining cases on the bf16 logits vector the head kernel leaves
in HBM: stochastic sampling (mlx_lm.sample_utils.make_sampler order: top_p -> min_p ->
top_k -> categorical at 1/temperature) and top-k logprobs.  It runs as a handful of
torch ops on a 128K-element device vector -- control plane, not the hot loop.
"""
from __future__ import annotations

import torch

from dnet_b200.core.decoding.config import DecodingConfig
from dnet_b200.core.types.messages import TokenResult


class Sampler:
    @staticmethod
    def sample(logits: torch.Tensor, config: DecodingConfig, req_logprobs: bool = False,
               req_top_logprobs: int = 0, generator=None) -> TokenResult:
        if logits.dim() == 3:
            v = logits[:, -1, :][0]
        elif logits.dim() == 2:
            v = logits[-1]
        else:
            v = logits
        vf = v.to(torch.float32)
        if config.temperature == 0:
            token_id = int(torch.argmax(vf).item())
        else:
            lp = torch.log_softmax(vf, dim=-1)
            if 0.0 < config.top_p < 1.0:
                sp, si = torch.sort(lp.exp(), descending=False)
                cum = torch.cumsum(sp, dim=-1)
                keep = cum > 1.0 - config.top_p
                mask = torch.zeros_like(keep)
                mask[si] = keep
                lp = torch.where(mask, lp, torch.full_like(lp, float("-inf")))
            if config.min_p and config.min_p > 0.0:
                top = lp.max()
                keep = lp >= top + torch.log(torch.tensor(config.min_p, device=lp.device))
                k = max(1, int(config.min_tokens_to_keep))
                keep[torch.topk(lp, k).indices] = True
                lp = torch.where(keep, lp, torch.full_like(lp, float("-inf")))
            if config.top_k is not None and 0 < config.top_k < lp.numel():
                kth = torch.topk(lp, config.top_k).values[-1]
                lp = torch.where(lp >= kth, lp, torch.full_like(lp, float("-inf")))
            probs = torch.softmax(lp * (1.0 / config.temperature), dim=-1)
            token_id = int(torch.multinomial(probs, 1, generator=generator).item())
        logprob = 0.0
        top_logprobs: dict[int, float] = {}
        if req_logprobs or req_top_logprobs > 0:
            lse = torch.logsumexp(vf, dim=-1).to(v.dtype)
            log_probs = (vf - lse.to(torch.float32)).to(v.dtype)
            if req_logprobs:
                logprob = float(log_probs[token_id].item())
            if req_top_logprobs > 0:
                order = torch.argsort(vf, stable=True).flip(0)[:req_top_logprobs]
                for ii in order.tolist():
                    top_logprobs[int(ii)] = float(log_probs[int(ii)].item())
        return TokenResult(token_id=token_id, logprob=logprob, top_logprobs=top_logprobs)
