"""Chunked prefill, host side: the API splits a long prompt into ``tokens`` frames, all but
the last flagged ``Activation.batch_size == 0``; the flag survives the proto round trip and the policies' helper reads it."""
import asyncio

import numpy as np

from dnet_b200.api.inference import InferenceManager
from dnet_b200.api.strategies.ring import RingApiAdapter
from dnet_b200.core.types.messages import ActivationMessage, TokenResult
from dnet_b200.shard.policies._cuda_common import more_chunks_follow


class _FakeAdapter:
    def __init__(self):
        self.sent = []

    async def send_tokens(self, nonce, tokens, callback, logprobs=False, top_logprobs=0, decoding_config=None, more=False):
        self.sent.append((np.frombuffer(tokens, np.int32).tolist(), more))

    async def await_token(self, nonce, timeout_s):
        return TokenResult(token_id=7, logprob=0.0, top_logprobs={})

    async def lease(self, *a, **k):
        pass

    async def end_request(self, nonce):
        self.sent.append(("end", False))


def _run(prompt, chunk):
    ad = _FakeAdapter()
    mgr = InferenceManager(ad, "local://")

    async def go():
        return [r async for r in mgr.generate_stream("n", prompt, 1, prefill_chunk=chunk)]
    out = asyncio.run(go())
    assert [r.token_id for r in out] == [7]
    return ad.sent


def test_prompt_is_split_and_only_the_last_chunk_is_sampled():
    prompt = list(range(10))
    sent = _run(prompt, 4)
    assert sent == [([0, 1, 2, 3], True), ([4, 5, 6, 7], True), ([8, 9], False), ("end", False)]
    # exact multiple: the last full chunk is the sampled one
    assert _run(list(range(8)), 4)[:2] == [([0, 1, 2, 3], True), ([4, 5, 6, 7], False)]
    # no chunking asked, or a prompt shorter than one chunk: one ordinary frame
    assert _run(prompt, 0)[0] == (prompt, False)
    assert _run(prompt, 64)[0] == (prompt, False)


def test_flag_travels_in_the_unchanged_proto():
    ad = RingApiAdapter.__new__(RingApiAdapter)
    for more in (False, True):
        req = RingApiAdapter._request(ad, "n", "tokens", np.arange(3, dtype=np.int32).tobytes(), "local://", False, 0, None,
                                      more=more)
        assert req.activation.batch_size == (0 if more else 1)
        back = ActivationMessage.from_proto(req, pool_id=-1)
        assert more_chunks_follow(back) is more


def test_request_metrics_have_the_reference_keys():
    """reference api/inference.py:216-233 (ChatResponseModel.metrics when profile=true)"""
    ad = _FakeAdapter()
    mgr = InferenceManager(ad, "local://")
    m = {}

    async def go():
        return [r async for r in mgr.generate_stream("n", [1, 2, 3], 4, device_loop=False, metrics=m)]
    out = asyncio.run(go())
    assert len(out) == 4
    assert set(m) == {"total_ms", "ttfb_ms", "token_gen_ms", "tokens_generated", "tps_overall", "tps_decoding"}
    assert m["tokens_generated"] == 4 and m["total_ms"] >= m["ttfb_ms"] >= 0 and m["tps_decoding"] >= m["tps_overall"] > 0
