"""``fit`` policy: everything resident (reference src/dnet/shard/policies/fit_in_memory.py:15-236),
driving libdnet_b200 instead of MLX.

Differences that are deliberate and result-identical:
  * the per-layer Python loop + ``mx.eval`` per window is one ``dn_window_forward`` call
    (or, for single-token messages, one CUDA-graph replay of the whole shard step);
  * the per-layer cast to the wire dtype is the bf16 store of the down-proj epilogue;
  * the end shard projects the last position only and, at temperature 0, the argmax /
    logsumexp are fused into the lm_head kernel.
Conventions kept: runs on the compute thread under ``runtime._model_lock``; never raises;
releases the input buffer; emits exactly one output message per input on success.
"""
from __future__ import annotations

import ctypes as C

from dnet_b200 import _cabi
from dnet_b200.core.memory.weight_cache import WeightCache
from dnet_b200.core.types.messages import ActivationMessage, TokenResult
from dnet_b200.utils.logger import logger
from .. import frames as fr
from . import _cuda_common as cc
from .base import ComputePolicy, register_policy


@register_policy("fit")
class FitInMemoryPolicy(ComputePolicy):
    """Everything fits - no offloading needed"""

    def configure_policy_for_model(self, req) -> None:
        self._mode = "fit"
        self._run_arrays = {}            # tuple(run) -> ctypes int32 array handed to dn_shard_step
        self.sched_entries_done = 0      # decode steps launched from schedule frames (progress signal for drivers)
        self.sched_marks = {}            # real-entry count -> torch.cuda.Event recorded on the compute stream right BEFORE
                                         # the entry with that index is launched (drivers time a window of entries with it)
        self.sched_host_s = 0.0          # host seconds spent inside _process_sched, and the entries it covered
        self.sched_host_entries = 0
        from collections import deque
        self._tp_hist = deque()          # tensor-parallel head: recent schedule entries (None = bubble) ...
        self._tp_hist_base = 0           # ... the ring-wide index of hist[0] ...
        self._tp_index = 0               # ... and of the next entry
        local_count = max(1, len(self.runtime.assigned_layers))
        requested_w = max(1, int(req.window_size))
        self.window_size = min(local_count, requested_w)
        self.weight_cache = WeightCache(
            self.runtime.assigned_layers,
            self.runtime.model_metadata,
            window_size=self.window_size,
            prefetch_threads=self.runtime.prefetch_threads,
            resident_windows=self._resident_windows,
            use_mxload_fastpath=self.runtime.compute_config.mxload_fastpath,
            prefetch_mode=self.runtime.compute_config.prefetch_mode,
            stage_host=self.runtime.stage_host,
            keep_host_records=False,       # everything stays resident in HBM: staging records are dropped after the copy
        )

    # -- CUDA-graph fast path for single-token messages ----------------------------------
    def _graph_step(self, ns, x, is_tokens: bool, run: list[int], fused_head: bool) -> None:
        """Replay (capturing on first use) [embed] + window + [norm/head/argmax] + advance.
        The KV offset advances once per token: on the visit that computes the shard's last
        local layer (with k>1 rounds a shard is visited k times per token, api/utils.py:62-131)."""
        rt = self.runtime
        lib = _cabi.load()
        s = rt.compute_stream_ptr
        adv = 1 if run[-1] == rt._assigned_sorted[-1] else 0
        if rt.use_megakernel:
            # one persistent cooperative kernel for the whole step (dn_megakernel.cuh)
            key = tuple(run)
            arr = self._run_arrays.get(key)
            if arr is None:
                arr = self._run_arrays[key] = (C.c_int32 * len(run))(*run)
            _cabi.check(lib.dn_shard_step(rt.model._h, arr, len(run), x.data_ptr(), ns.kv._h, int(is_tokens),
                                          int(fused_head), ns.result_token_ptr if fused_head else None,
                                          ns.result_logprob_ptr if fused_head else None, None, adv, s))
            return
        key = (run[0], is_tokens, fused_head, adv)
        g = ns.graphs.get(key)
        if g is None:
            rt.compute_stream.synchronize()
            _cabi.check(lib.dn_graph_begin(s))
            gp = C.c_void_p()
            try:
                if is_tokens:
                    _cabi.check(lib.dn_embed(rt.model._h, ns.kv.token_ptr, 1, x.data_ptr(), s))
                arr = (C.c_int32 * len(run))(*run)
                _cabi.check(lib.dn_window_forward(rt.model._h, arr, len(run), x.data_ptr(), 1, ns.kv._h, s))
                if fused_head:
                    _cabi.check(lib.dn_head_sample_greedy(rt.model._h, x.data_ptr(), 1, ns.kv._h,
                                                          ns.result_token_ptr, ns.result_logprob_ptr, s))
                if adv:
                    _cabi.check(lib.dn_kv_advance(ns.kv._h, 1, s))
            finally:
                rc = lib.dn_graph_end(s, C.byref(gp))
            _cabi.check(rc)
            g = gp.value
            ns.graphs[key] = g
        _cabi.check(lib.dn_graph_launch(g, s))
        if adv:
            ns.kv.note_advance(1)

    # -- device-closed decode: one fused wait + step + hop kernel per scheduled (lane, seq) -----
    def _process_sched(self, msg: ActivationMessage) -> None:
        """A ``b200.sched`` frame (shard/frames.py): the head shard's ordered decode schedule.  Every shard
        launches the entries in exactly this order on its one compute stream, so kernels that spin on a
        predecessor's flag can never wait behind a kernel that (transitively) waits for them.
        Per entry: wait for flag ``seq`` of the lane (the token on the head shard, the activation
        elsewhere), run this shard's layers (in place in the lane slot), store the result into the
        successor's slot and release its flag -- all inside ``dn_shard_step_hop``; the finalising shard
        hands the token to the head's slot with ``seq + 1`` and to the host through the TokenTap."""
        rt = self.runtime
        hop = rt.hop
        ticket = msg.sched_done
        import time as _time
        _t0 = _time.perf_counter()
        try:
            with rt._model_lock:
                if not cc.model_ready(rt) or hop is None or not hop.connected:
                    logger.error("Runtime %s: schedule frame without a model / hop link", rt.shard_id)
                    return
                lib = _cabi.load()
                run = list(rt._assigned_sorted)
                if run != list(range(run[0], run[0] + len(run))):
                    logger.error("on-device decode needs one contiguous run of local layers (k=1); got %s", run)
                    return
                first = run[0] == 0
                last = run[-1] + 1 >= rt.model_metadata.num_layers
                to_bind = self._bind_layer_weights(run, msg)
                if to_bind is None:
                    return
                if to_bind:
                    cc.wait_layers_ready(rt, self.weight_cache, run)
                    rt.model.load_weights(list(to_bind.items()), strict=False)
                key = tuple(run)
                arr = self._run_arrays.get(key)
                if arr is None:
                    arr = self._run_arrays[key] = (C.c_int32 * len(run))(*run)
                s = rt.compute_stream_ptr
                now = None
                if getattr(rt, "tp_head", None):
                    self._launch_sched_tp(msg.sched, run, arr, first, last)
                    return
                for lane, seq in msg.sched:
                    nonce = rt.lane_nonce.get(lane)
                    ns = rt._kv_by_nonce.get(nonce) if nonce is not None else None
                    if ns is None:
                        logger.error("schedule entry for lane %d: no request holds that lane on shard %s", lane, rt.shard_id)
                        continue
                    tok_ptr = lp_ptr = None
                    if last:
                        tok_ptr, lp_ptr = rt.token_tap.post(lane, (nonce, seq, ns.params))
                    if self.sched_marks:
                        self._hit_mark()
                    _cabi.check(lib.dn_shard_step_hop(
                        rt.model._h, arr, len(run), ns.x1.data_ptr() if first else hop.rx.slot(lane), ns.kv._h,
                        1 if first else 0, 1 if last else 0, tok_ptr, lp_ptr, 1,
                        hop.rx.flag(lane), seq, hop.rx.slot(lane) if first else None,
                        hop.tx_slot(lane), hop.tx_flag(lane), seq + 1 if last else seq, s))
                    if now is None:
                        import time
                        now = time.perf_counter()
                    rt._kv_last_seen[nonce] = now
                    self.sched_entries_done += 1
                if not (len(rt._assigned_sorted) <= self.window_size):
                    self.weight_cache.decrease_references(run)
        except Exception as e:
            logger.exception("Error launching scheduled decode steps: %s", e)
        finally:
            self.sched_host_s += _time.perf_counter() - _t0      # host time spent launching (drivers report it per entry)
            self.sched_host_entries += len(msg.sched or ())
            if ticket is not None:
                ev = None
                try:
                    import torch
                    ev = torch.cuda.Event()
                    ev.record(rt.compute_stream)
                except Exception:
                    ev = None
                ticket.record(ev)

    def _hit_mark(self) -> None:
        ev = self.sched_marks.pop(self.sched_entries_done, None)
        if ev is not None:
            ev.record(self.runtime.compute_stream)

    def _launch_sched_tp(self, entries, run, arr, first: bool, last: bool) -> None:
        """Schedule entries with the lm_head tensor-parallel over the ring.

        Entry i of the ring-wide schedule makes shard r (of S) launch ONE kernel that
          * first serves the head part of the request of entry i - (S - r): by then the last shard has broadcast
            that request's final hidden state to every shard (it ran entry i - (S - r) one kernel earlier in
            wall-clock terms -- the pipeline skew), so all shards compute their vocabulary slices in the same
            slot and store the partial (max, sum-exp, argmax) into the head shard's table;
          * then runs its layers for entry i's own request (nothing for a bubble entry);
          * on the last shard, broadcasts the resulting final hidden state;
          * on the head shard, finally merges the S partials of the head part it started with into the token,
            hands it to the request's next step (own lane slot + flag) and to the host (TokenTap).
        The history of entries is all a shard needs; it is identical on every shard because the schedule is."""
        import time

        rt = self.runtime
        hop, mesh = rt.hop, rt.hop.mesh
        lib = _cabi.load()
        S, r, lag = int(rt.tp_head["S"]), int(rt.tp_head["r"]), int(rt.tp_head.get("lag", 0))
        hist = self._tp_hist
        s = rt.compute_stream_ptr
        now = time.perf_counter()
        for lane, seq in entries:
            real = lane != fr.BUBBLE
            idx = self._tp_index
            self._tp_index += 1
            hist.append((lane, seq) if real else None)
            if len(hist) > 64:
                hist.popleft()
                self._tp_hist_base += 1
            j = idx - (S - r) - lag - self._tp_hist_base
            due = hist[j] if 0 <= j < len(hist) else None
            tp = _cabi.TpArgs()
            if due is not None:
                dl, dq = due
                head_d, head_f = mesh.peers[0]
                tp.hp_x, tp.hp_wait_flag, tp.hp_seq = mesh.x_slot(dl), mesh.x_flag(dl), dq
                tp.hp_dst, tp.hp_dst_flag = mesh.partial(dl, r, head_d), mesh.p_flag(dl, r, head_f)
                if r == 0:
                    dn = rt.lane_nonce.get(dl)
                    dns = rt._kv_by_nonce.get(dn) if dn is not None else None
                    tp.mg_n, tp.mg_part, tp.mg_flags, tp.mg_seq = S, mesh.partial(dl, 0), mesh.p_flag(dl, 0), dq
                    if dns is not None:
                        tok_ptr, lp_ptr = rt.token_tap.post(dl, (dn, dq, dns.params))
                        tp.mg_kv, tp.mg_token_out, tp.mg_logprob_out = dns.kv._h, tok_ptr, lp_ptr
                    tp.mg_slot, tp.mg_slot_flag, tp.mg_slot_seq = hop.rx.slot(dl), hop.rx.flag(dl), dq + 1
            if not real:
                if due is not None:
                    _cabi.check(lib.dn_shard_step_tp(rt.model._h, None, 0, None, None, 0, 0, None, 0, None, None, None, 0,
                                                     C.byref(tp), s))
                continue
            nonce = rt.lane_nonce.get(lane)
            ns = rt._kv_by_nonce.get(nonce) if nonce is not None else None
            if ns is None:
                logger.error("schedule entry for lane %d: no request holds that lane on shard %s", lane, rt.shard_id)
                continue
            if self.sched_marks:
                self._hit_mark()
            if last:
                tp.bc_n, tp.bc_seq = S, seq
                for d in range(S):
                    pd, pf = mesh.peers[d]
                    tp.bc_dst[d], tp.bc_flag[d] = mesh.x_slot(lane, pd), mesh.x_flag(lane, pf)
            _cabi.check(lib.dn_shard_step_tp(
                rt.model._h, arr, len(run), ns.x1.data_ptr() if first else hop.rx.slot(lane), ns.kv._h, 1 if first else 0, 1,
                hop.rx.flag(lane), seq, hop.rx.slot(lane) if first else None,
                None if last else hop.tx_slot(lane), None if last else hop.tx_flag(lane), seq, C.byref(tp), s))
            rt._kv_last_seen[nonce] = now
            self.sched_entries_done += 1

    def _process_seed(self, msg: ActivationMessage) -> None:
        """A seeded ``b200.lease``: put ``msg.token_id`` into the head shard's own lane slot and publish
        the lane's next decode sequence number (used when a lease follows host-driven steps)."""
        rt = self.runtime
        try:
            with rt._model_lock:
                ns = rt._kv_by_nonce.get(msg.nonce)
                if ns is None or rt.hop is None or msg.lane < 0:
                    logger.error("seeded lease for unknown nonce %s", msg.nonce)
                    return
                lib = _cabi.load()
                ns.kv.set_token(int(msg.token_id), rt.compute_stream_ptr)
                _cabi.check(lib.dn_hop_send(rt.hop.rx.slot(msg.lane), ns.kv.token_ptr, 4, rt.hop.rx.flag(msg.lane),
                                            msg.seq0, rt.compute_stream_ptr))
        except Exception as e:
            logger.exception("Error seeding lease: %s", e)

    def process(self, msg: ActivationMessage) -> None:
        if msg.dtype == fr.SCHED_DTYPE:
            return self._process_sched(msg)
        if msg.dtype == fr.LEASE_DTYPE:
            return self._process_seed(msg)
        rt = self.runtime
        ns = None
        try:
            with rt._model_lock:
                if not cc.model_ready(rt):
                    logger.error("Runtime %s: cannot process activation - model not loaded", rt.shard_id)
                    return
                # 1) per-nonce KV (+ activation buffer, captured graphs, pinned result)
                ns = rt.get_or_make_kv(msg.nonce)
                cc.note_lane(rt, msg, ns)
                T = cc.msg_tokens(rt, msg)
                current_layer = msg.layer_id + 1
                run = cc.local_run(rt, current_layer)
                if not run or T <= 0:
                    logger.error("layer %s not hosted on shard %s (or empty message)", current_layer, rt.shard_id)
                    rt.input_pool.release(msg.pool_id)
                    return
                if ns.kv.offset + T > ns.kv.max_tokens:
                    logger.error("KV capacity exceeded for nonce %s: %d + %d > %d", msg.nonce, ns.kv.offset, T,
                                 ns.kv.max_tokens)
                    rt.input_pool.release(msg.pool_id)
                    return
                last_layer = run[-1]
                is_end = last_layer + 1 >= rt.model_metadata.num_layers
                greedy = msg.temperature == 0 and msg.req_top_logprobs <= 0
                more = cc.more_chunks_follow(msg)      # a prompt chunk that is not the last: fill the KV, sample nothing

                # 2) bind (fast exit when everything is already bound)
                to_bind = self._bind_layer_weights(run, msg)
                if to_bind is None:
                    return
                rt._compute_busy.set()
                if to_bind:
                    cc.wait_layers_ready(rt, self.weight_cache, run)
                    rt.model.load_weights(list(to_bind.items()), strict=False)
                    for other in rt.all_nonce_states():
                        other.drop_graphs()

                # 3) stage x, 4) compute the run
                final = None
                use_graph = (rt.use_cuda_graphs or rt.use_megakernel) and T == 1 and self.window_size >= len(run)
                if use_graph:
                    is_tokens = msg.dtype == "tokens"
                    if is_tokens:
                        buf = rt.input_pool.get_buffer(msg.pool_id)
                        if buf is None:
                            logger.error("Failed to get input buffer %s", msg.pool_id)
                            return
                        ns.kv.set_token(int(buf[0]), rt.compute_stream_ptr)
                        x = ns.x_view(1)
                    else:
                        staged = cc.stage_input(rt, msg, ns)
                        if staged is None:
                            logger.error("Failed to get input buffer %s", msg.pool_id)
                            return
                        x = staged[0]
                    self._graph_step(ns, x, is_tokens, run, is_end and greedy and not more)
                    self.weight_cache.decrease_references(run)
                    if is_end and greedy and not more:
                        rt.compute_stream.synchronize()
                        if int(ns.result_np_i32[0]) <= -1000:
                            # a bounded in-kernel wait timed out: the step's results are invalid (sticky until cleared)
                            code = -int(ns.result_np_i32[0]) - 1000
                            rt.step_errors += 1
                            _cabi.load().dn_step_error_clear(rt.model._h, rt.compute_stream_ptr)
                            logger.error("step kernel error %d for nonce %s: request failed, no token emitted", code, msg.nonce)
                            cc.finish_input(rt, msg, ns)
                            return
                        final = TokenResult(token_id=int(ns.result_np_i32[0]),
                                            logprob=float(ns.result_np_f32[1]) if msg.req_logprobs else 0.0,
                                            top_logprobs={})
                else:
                    staged = cc.stage_input(rt, msg, ns)
                    if staged is None:
                        logger.error("Failed to get input buffer %s", msg.pool_id)
                        return
                    x = staged[0]
                    for w0 in range(0, len(run), self.window_size):
                        window_layers = run[w0:w0 + self.window_size]
                        rt.model.window_forward(window_layers, x, ns.kv, rt.compute_stream_ptr)
                        for lid in window_layers:
                            self.weight_cache.decrease_reference(lid)
                    if last_layer == rt._assigned_sorted[-1]:
                        ns.kv.advance(T, rt.compute_stream_ptr)
                if is_end and more:
                    cc.finish_input(rt, msg, ns)       # the last shard swallows an intermediate prompt chunk
                    return
                if is_end and final is None:
                    try:
                        final = cc.sample_end_shard(rt, msg, ns, x)
                    except Exception as e:
                        logger.error("End-shard sampling failed: %s", e)
                        rt.input_pool.release(msg.pool_id)
                        return
                output_msg = cc.build_output(rt, msg, x, last_layer, final)
                rt.emit_result(output_msg)
                cc.finish_input(rt, msg, ns)
                return
        except Exception as e:
            logger.exception("Error in fit policy process: %s", e)
            try:
                if rt.input_pool:
                    cc.finish_input(rt, msg, ns)
            except Exception:
                pass
        finally:
            try:
                rt._compute_busy.clear()
            except Exception:
                pass

    def clear(self):
        try:
            if self.weight_cache:
                self.weight_cache.cancel_all_prefetch()
        except Exception:
            pass
        for layer_id in list(self._bound_versions.keys()):
            try:
                self.weight_cache.evict_layer(layer_id)
            except Exception:
                pass
        try:
            self._bound_versions.clear()
        except Exception:
            self._bound_versions = {}
