"""-m gpu: the ring-hop data plane and layer-swap plumbing of the C ABI on one device."""
import ctypes as C

import pytest
import torch

from dnet_b200 import _cabi

pytestmark = pytest.mark.gpu


def test_hop_send_wait_and_bounded_timeout(cuda_lib):
    lib = cuda_lib
    slot, flags = C.c_void_p(), C.c_void_p()
    _cabi.check(lib.dn_hop_alloc(8192, C.byref(slot)))
    _cabi.check(lib.dn_hop_alloc(256, C.byref(flags)))
    s_tx, s_rx = torch.cuda.Stream(), torch.cuda.Stream()
    src = torch.arange(4096, dtype=torch.int16, device="cuda")
    dst_view = torch.empty(4096, dtype=torch.int16).pin_memory()
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()          # tensors above were made on torch's stream
    # receiver first: the wait kernel spins on seq 1 while the sender's hop kernel runs on another stream
    _cabi.check(lib.dn_hop_wait(flags.value, 1, 2000, err.data_ptr(), s_rx.cuda_stream))
    _cabi.check(lib.dn_hop_send(slot.value, src.data_ptr(), 8192, flags.value, 1, s_tx.cuda_stream))
    # (enqueued after the send: a copy queued behind the spinning wait would otherwise share a
    #  copy-engine queue with any cudaMemcpy-based send -- the reason dn_hop_send is a kernel)
    _cabi.check(lib.dn_memcpy_d2h(dst_view.data_ptr(), slot.value, 8192, s_rx.cuda_stream))
    s_rx.synchronize()
    assert torch.equal(dst_view, src.cpu()) and int(err.item()) == 0
    # large payloads take the cudaMemcpyAsync + flag path
    big = torch.arange(1 << 20, dtype=torch.int16, device="cuda")
    bslot = C.c_void_p()
    _cabi.check(lib.dn_hop_alloc(2 << 20, C.byref(bslot)))
    torch.cuda.synchronize()
    _cabi.check(lib.dn_hop_send(bslot.value, big.data_ptr(), 2 << 20, flags.value + 64, 7, s_tx.cuda_stream))
    _cabi.check(lib.dn_hop_wait(flags.value + 64, 7, 2000, err.data_ptr(), s_rx.cuda_stream))
    bdst = torch.empty(1 << 20, dtype=torch.int16).pin_memory()
    _cabi.check(lib.dn_memcpy_d2h(bdst.data_ptr(), bslot.value, 2 << 20, s_rx.cuda_stream))
    s_rx.synchronize()
    assert torch.equal(bdst, big.cpu()) and int(err.item()) == 0
    _cabi.check(lib.dn_hop_free(bslot.value))
    # a flag that never arrives: the wait kernel gives up after the timeout and reports it
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(s_rx)
    _cabi.check(lib.dn_hop_wait(flags.value, 99, 30, err.data_ptr(), s_rx.cuda_stream))
    t1.record(s_rx)
    s_rx.synchronize()
    assert int(err.item()) == 1 and 20 <= t0.elapsed_time(t1) < 500
    h = (C.c_uint8 * 64)()
    _cabi.check(lib.dn_hop_export(slot.value, h))
    assert any(h)
    _cabi.check(lib.dn_hop_free(slot.value))
    _cabi.check(lib.dn_hop_free(flags.value))


def test_slot_prefetch_pinned_to_hbm_with_event(cuda_lib):
    lib = cuda_lib
    host, dev, ev, st = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
    n = 1 << 20
    _cabi.check(lib.dn_pinned_alloc(n, C.byref(host)))
    _cabi.check(lib.dn_device_alloc(n, C.byref(dev)))
    _cabi.check(lib.dn_event_create(C.byref(ev), 0))
    _cabi.check(lib.dn_stream_create(C.byref(st), 0))
    C.memset(host.value, 0x5A, n)
    _cabi.check(lib.dn_slot_prefetch(dev.value, host.value, n, st.value, ev.value))
    _cabi.check(lib.dn_event_sync(ev.value))
    assert lib.dn_event_query(ev.value) == 1
    back = torch.empty(n, dtype=torch.uint8).pin_memory()
    _cabi.check(lib.dn_memcpy_d2h(back.data_ptr(), dev.value, n, st.value))
    _cabi.check(lib.dn_stream_sync(st.value))
    assert int(back.min()) == 0x5A and int(back.max()) == 0x5A
    for f, p in ((lib.dn_event_destroy, ev), (lib.dn_stream_destroy, st), (lib.dn_device_free, dev), (lib.dn_pinned_free, host)):
        _cabi.check(f(p.value))


def test_launch_counter_counts_graph_nodes(cuda_lib):
    assert cuda_lib.dn_launch_count() >= 0 and cuda_lib.dn_device_sm_count() >= 100


def test_fused_hop_step_matches_unfused(cuda_lib):
    """dn_shard_step_hop on one device (self-loop slots): waits on a pre-set flag, computes the
    same step as dn_shard_step, and publishes activation + flag / token + flag."""
    from tests.helpers import load_golden, make_runtime, oracle_weights, token_message
    lib = cuda_lib
    g = load_golden("tiny_llama")
    cfgd = g["config"]
    w = oracle_weights(cfgd, g["wseed"])
    H = cfgd["hidden_size"]
    a = make_runtime(cfgd, w, [0, 1], shard_id="a")
    b = make_runtime(cfgd, w, [2, 3], shard_id="b")
    try:
        # prefill both shards through the policy (unfused), take the first token
        a.policy.process(token_message(a, "n", g["prompt"].tolist()))
        mid = a.activation_send_queue.get_nowait()
        b.policy.process(mid)
        first = b.activation_send_queue.get_nowait()
        assert first.is_final and first.token_id == int(g["tokens"][0])
        nsa, nsb = a._kv_by_nonce["n"], b._kv_by_nonce["n"]
        slot, flags = C.c_void_p(), C.c_void_p()
        _cabi.check(lib.dn_hop_alloc(H * 2 + 64, C.byref(slot)))     # activation slot for b (+ token line)
        _cabi.check(lib.dn_hop_alloc(256, C.byref(flags)))
        tokslot = torch.tensor([first.token_id, 0, 0, 0], dtype=torch.int32, device="cuda")
        tok_out = torch.zeros(4, dtype=torch.int32, device="cuda")
        lp_out = torch.zeros(4, dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        sa, sb = a.compute_stream_ptr, b.compute_stream_ptr
        la = (C.c_int32 * 2)(0, 1)
        lb = (C.c_int32 * 2)(2, 3)
        toks = []
        for step in range(1, 6):
            # shard a: token_in = tokslot, no wait; sends its activation into `slot`, flag[0] = step
            _cabi.check(lib.dn_shard_step_hop(a.model._h, la, 2, nsa.x1.data_ptr(), nsa.kv._h, 1, 0, None, None, 1,
                                              None, 0, tokslot.data_ptr(), slot.value, flags.value, step, sa))
            # shard b: waits for flag[0] >= step, computes in place in the slot, samples, sends the token back
            _cabi.check(lib.dn_shard_step_hop(b.model._h, lb, 2, slot.value, nsb.kv._h, 0, 1, tok_out.data_ptr(),
                                              lp_out.data_ptr(), 1, flags.value, step, None, tokslot.data_ptr(),
                                              flags.value + 64, step, sb))
            b.compute_stream.synchronize()
            a.compute_stream.synchronize()
            toks.append(int(tok_out[0].item()))
            assert int(tokslot[0].item()) == toks[-1]
        assert toks == g["tokens"][1:6].tolist()
        assert lib.dn_step_error(a.model._h, sa) == 0 and lib.dn_step_error(b.model._h, sb) == 0
        _cabi.check(lib.dn_hop_free(slot.value))
        _cabi.check(lib.dn_hop_free(flags.value))
    finally:
        a.unload_model_core()
        b.unload_model_core()
