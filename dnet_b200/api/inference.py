"""The API node's token loop (reference src/dnet/api/inference.py:135-212), on token ids.

Tokenisation, chat templates and the HTTP surface are out of scope; what is
rebuilt is the loop that closes the ring: send the prompt, then either

* ``device_loop=False`` -- the reference's loop: await the token, send it back as a one-token
  ``tokens`` frame, repeat (two control hops + Python per token), or
* ``device_loop=True`` -- lease decode steps to the ring in chunks; the last shard's kernel hands
  each token straight to the first shard over NVLink and the API only *observes* the tokens
  (N2).  The lease is renewed ``lease_ahead`` tokens before it runs out; on a stop token the request
  is ended and at most the outstanding lease is wasted.
"""
from __future__ import annotations

import time
from typing import AsyncIterator, Iterable, Optional, Sequence

import numpy as np

from dnet_b200.core.decoding.config import DecodingConfig
from dnet_b200.core.types.messages import TokenResult


class InferenceManager:
    def __init__(self, adapter, callback_addr: str, request_timeout_s: float = 30.0):
        self.adapter = adapter
        self.callback_addr = callback_addr
        self.request_timeout_s = request_timeout_s

    def resolve_request(self, nonce: str, result: TokenResult) -> None:
        self.adapter.resolve_token(nonce, result)

    async def generate_stream(self, nonce: str, prompt_ids: Sequence[int], max_tokens: int, *,
                              decoding: Optional[DecodingConfig] = None, stop_ids: Iterable[int] = (),
                              logprobs: bool = False, device_loop: bool = True, lease_steps: int = 16,
                              lease_ahead: int = 8, prefill_chunk: int = 0,
                              metrics: Optional[dict] = None) -> AsyncIterator[TokenResult]:
        """``metrics``: a dict to fill with the reference's per-request profile (api/inference.py:216-233; the keys
        of ``ChatResponseModel.metrics``): total_ms, ttfb_ms, token_gen_ms, tokens_generated, tps_overall, tps_decoding."""
        stop = set(int(t) for t in stop_ids)
        ad = self.adapter
        dec = decoding or DecodingConfig(temperature=0.0)
        device_loop = device_loop and float(dec.temperature) == 0.0     # the fused step samples greedily
        t_start = time.perf_counter()
        ids = np.asarray(list(prompt_ids), np.int32)
        step = int(prefill_chunk) if prefill_chunk and prefill_chunk > 0 else max(1, len(ids))
        # chunked prefill: the chunks stream through the ring back to back (shard r works on chunk k while
        # shard r+1 works on chunk k-1); only the last one is sampled
        for c0 in range(0, max(1, len(ids)), step):
            await ad.send_tokens(nonce, ids[c0:c0 + step].tobytes(), self.callback_addr, logprobs=logprobs,
                                 decoding_config=dec, more=c0 + step < len(ids))
        produced = 0
        leased = 0
        t_first = None
        try:
            while produced < max_tokens:
                res = await ad.await_token(nonce, self.request_timeout_s)
                if t_first is None:
                    t_first = time.perf_counter()
                produced += 1
                yield res
                if res.token_id in stop or res.token_id < 0 or produced >= max_tokens:
                    break
                if device_loop:
                    # `leased` counts decode steps granted after the prompt's own first token
                    if leased - (produced - 1) <= lease_ahead:
                        n = min(lease_steps, max_tokens - 1 - leased)
                        if n > 0:
                            await ad.lease(nonce, n, self.callback_addr)
                            leased += n
                else:
                    await ad.send_tokens(nonce, np.asarray([res.token_id], np.int32).tobytes(), self.callback_addr,
                                         logprobs=logprobs, decoding_config=dec)
        finally:
            if metrics is not None:
                t_end = time.perf_counter()
                total_s = max(t_end - t_start, 1e-9)
                gen_s = max(t_end - (t_first or t_start), 1e-9)
                metrics.update({
                    "total_ms": round(total_s * 1e3, 3), "ttfb_ms": round(((t_first or t_end) - t_start) * 1e3, 3),
                    "token_gen_ms": round(gen_s * 1e3, 3), "tokens_generated": produced,
                    "tps_overall": round(produced / total_s if produced else 0.0, 4),
                    "tps_decoding": round(produced / gen_s if produced else 0.0, 4)})
            await ad.end_request(nonce)
