"""``offload`` / ``sliding_fit`` policy: more layers than HBM slots
(reference src/dnet/shard/policies/offload.py:19-455), rebuilt on CUDA streams.

Reference behaviour kept: mode selection (residency < window -> sliding_fit), window walk,
bind-by-version, post-window eviction for resident_windows <= 1, delta-swap eviction for
sliding_fit, prefetch of the next local window (wrapping to the first window for the next
token) recorded in ``_prepared_by_nonce``.

What the B200 build changes underneath: layer records live in pinned host memory (packed
once at load), a "load" is one cudaMemcpyAsync on the prefetch stream into a recycled HBM
slot, and ordering is by events -- the compute stream waits for a layer's ready event, the
prefetch stream waits for the slot's release event -- so no host thread blocks on a copy
and the copy of window i+1 overlaps the kernels of window i inside one message.
"""
from __future__ import annotations

import asyncio
import time
from concurrent.futures import Future

from dnet_b200.core.memory.weight_cache import WeightCache
from dnet_b200.core.types.messages import ActivationMessage
from dnet_b200.utils.logger import logger
from .. import frames as fr
from . import _cuda_common as cc
from .base import ComputePolicy, register_policy


@register_policy("offload")
@register_policy("sliding_fit")
class OffloadPolicy(ComputePolicy):
    def configure_policy_for_model(self, req) -> None:
        local_count = max(1, len(self.runtime.assigned_layers))
        requested_w = max(1, int(req.window_size))
        n_residency = int(max(1, int(req.residency_size)))
        if n_residency < requested_w:
            self._mode = "sliding_fit"
            self.window_size = max(1, min(n_residency, local_count))
        else:
            self._mode = "offload"
            self.window_size = max(1, min(requested_w, local_count))
        self._resident_windows = int(self._resident_windows) if self._resident_windows else 1
        # reference: repack the assigned layers into per-layer files (utils/repack.py) and
        # mx.load them; here the per-layer record is packed into pinned host memory
        t0 = time.perf_counter()
        self.runtime.stage_host = True
        self.weight_cache = WeightCache(
            self.runtime.assigned_layers,
            self.runtime.model_metadata,
            window_size=self.window_size,
            prefetch_threads=self.runtime.prefetch_threads,
            resident_windows=self._resident_windows,
            use_mxload_fastpath=True,
            prefetch_mode="off",
            stage_host=True,
        )
        try:
            nbytes = self.weight_cache.layer_manager.stage_all_to_host()
            logger.info("[REPACK] shard=%s layers=%s pinned_bytes=%d ms=%.1f", self.runtime.shard_id,
                        len(self.runtime._assigned_sorted), nbytes, (time.perf_counter() - t0) * 1e3)
        except Exception as e:
            logger.warning("Runtime %s: host staging failed: %s", self.runtime.shard_id, e)
        logger.info("OffloadPolicy configured: mode=%s window=%d resident=%d", self._mode, self.window_size,
                    self._resident_windows)

    def _prepare_window_blocking(self, window_layers: list[int]) -> None:
        """Materialise a window's weights (enqueue their pinned->HBM copies)."""
        if not self.weight_cache:
            return
        for lid in window_layers:
            _ = self.weight_cache.get_weight(lid, inc_ref=False)

    def _schedule_prefetch(self, nonce: str, next_window: list[int]) -> None:
        rt = self.runtime
        loop = rt._loop
        if loop is not None:
            fut = loop.run_in_executor(rt.executor, self._prepare_window_blocking, next_window)
        else:
            fut = rt.executor.submit(self._prepare_window_blocking, next_window)
        self._prepared_by_nonce[nonce] = (next_window, fut)

    def _drop_window(self, layers) -> None:
        self.weight_cache.evict_layers(layers)
        self.runtime.model.unload_layers(layers)
        for lid in layers:
            self._bound_versions.pop(lid, None)

    def _retire_window(self, window_layers, did_early_swap: bool) -> None:
        """What stays resident after a window ran (behaviour of reference offload.py:253-312).
        offload, 1 resident window : the window that just ran is evicted at once (its slots take the next one)
        offload, n resident windows: windows are kept in arrival order; with eager unloading the oldest go as soon
                                     as more than n are remembered (deferred unloading leaves that to the cache's LRU)
        sliding_fit, 1 window      : delta swap -- of the previous window only what still fits beside the new one stays
                                     (already done up front when the policy swapped early)
        sliding_fit, n windows     : just remembered; the weight cache's budget does the rest"""
        curr = list(window_layers)
        single = int(self._resident_windows) <= 1
        if self._mode != "sliding_fit":
            self._recent_windows.append(curr)
            if single:
                self._drop_window(self._recent_windows.pop(0))
            elif not self._defer_unload:
                keep = max(1, int(self._resident_windows))
                while len(self._recent_windows) > keep:
                    self._drop_window(self._recent_windows.pop(0))
            return
        if not single:
            self._recent_windows.append(curr)
            return
        if did_early_swap:
            return
        if not self._recent_windows:
            self._recent_windows.append(curr)
            return
        prev = self._recent_windows.pop(0)
        self._delta_swap_eviction(curr, prev)
        room = max(0, max(1, int(self.window_size or 1)) - len(curr))
        survivors = [lid for lid in prev if lid not in curr]
        self._recent_windows.append((survivors[-room:] if room else []) + curr)

    def process(self, msg: ActivationMessage) -> None:
        rt = self.runtime
        if msg.dtype in (fr.SCHED_DTYPE, fr.LEASE_DTYPE):
            # on-device decode leases need every layer resident (one persistent kernel per step); a shard that
            # swaps layers serves decode through the per-message path and the API must drive tokens itself
            logger.error("shard %s runs in %s mode: on-device decode schedules are not supported here", rt.shard_id, self._mode)
            if msg.sched_done is not None:
                msg.sched_done.record(None)
            return
        if not cc.model_ready(rt):
            logger.error("Runtime %s: cannot process activation - model not loaded", rt.shard_id)
            return
        try:
            with rt._model_lock:
                if not cc.model_ready(rt):
                    logger.error("Runtime %s: cannot process activation - model not loaded", rt.shard_id)
                    return
                ns = rt.get_or_make_kv(msg.nonce)
                cc.note_lane(rt, msg, ns)
                T = cc.msg_tokens(rt, msg)
                if T <= 0 or ns.kv.offset + T > ns.kv.max_tokens:
                    logger.error("bad message size / KV capacity exceeded for nonce %s", msg.nonce)
                    rt.input_pool.release(msg.pool_id)
                    return
                staged = cc.stage_input(rt, msg, ns)
                if staged is None:
                    logger.error("Failed to get input buffer %s", msg.pool_id)
                    return
                x = staged[0]
                current_layer = msg.layer_id + 1
                last_layer = current_layer - 1
                while True:
                    did_early_swap = False
                    window_layers: list[int] = []
                    for i in range(self.window_size):
                        layer = current_layer + i
                        if layer not in rt._assigned_set:
                            break
                        window_layers.append(layer)
                    if not window_layers:
                        break

                    # wait for the prefetch task that was scheduled for this window (the task
                    # only ENQUEUES copies; stream ordering does the real waiting)
                    if self._mode == "offload":
                        prep = self._prepared_by_nonce.get(msg.nonce)
                        if prep is not None:
                            layers, fut = prep
                            if layers == window_layers and fut is not None:
                                try:
                                    if isinstance(fut, Future):
                                        fut.result(timeout=30)
                                    elif not fut.done():
                                        t_end = time.time() + 30
                                        while not fut.done() and time.time() < t_end:
                                            time.sleep(0.0002)
                                except Exception:
                                    pass

                    if self._mode == "sliding_fit" and int(self._resident_windows) <= 1:
                        try:
                            resident = self.weight_cache.get_resident_layers()
                        except Exception:
                            resident = []
                        if self._delta_swap_eviction(window_layers, resident) > 0:
                            did_early_swap = True

                    to_bind = self._bind_layer_weights(window_layers, msg)
                    if to_bind is None:
                        return
                    rt._compute_busy.set()
                    # order the compute stream after each layer's latest pinned->HBM copy
                    cc.wait_layers_ready(rt, self.weight_cache, window_layers)
                    if to_bind:
                        rt.model.load_weights(list(to_bind.items()), strict=False)

                    # overlap: while this window computes, start the next window's copies when the
                    # budget has room for it (resident_windows >= 2)
                    if self._mode == "offload" and int(self._resident_windows) >= 2:
                        nxt_w = self._next_local_layers(rt._assigned_sorted, window_layers[-1], self.window_size)
                        if nxt_w:
                            self._schedule_prefetch(msg.nonce, nxt_w)

                    rt.model.window_forward(window_layers, x, ns.kv, rt.compute_stream_ptr)
                    last_layer = window_layers[-1]
                    for lid in window_layers:
                        self.weight_cache.decrease_reference(lid, release_event=cc.release_event(rt))

                    try:
                        self._retire_window(window_layers, did_early_swap)
                    except Exception:
                        pass

                    nxt = last_layer + 1
                    if nxt in rt._assigned_set:
                        current_layer = nxt
                        # resident_windows <= 1: the slots of the window just evicted are free now;
                        # enqueue the next window's copies immediately so they overlap the tail of
                        # the kernels still running on the compute stream
                        if self._mode == "offload" and int(self._resident_windows) <= 1:
                            nw = [l for l in range(nxt, nxt + self.window_size) if l in rt._assigned_set]
                            self._prepare_window_blocking(nw)
                        continue
                    break

                if last_layer == rt._assigned_sorted[-1]:   # once per token even with k>1 rounds
                    ns.kv.advance(T, rt.compute_stream_ptr)
                final = None
                is_end = last_layer + 1 >= rt.model_metadata.num_layers
                if is_end and not cc.more_chunks_follow(msg):
                    try:
                        final = cc.sample_end_shard(rt, msg, ns, x)
                    except Exception as e:
                        logger.error("End-shard sampling failed: %s", e)
                        rt.input_pool.release(msg.pool_id)
                        return
                if is_end and final is None:
                    cc.finish_input(rt, msg, ns)       # intermediate prompt chunk on the last shard: nothing to emit
                else:
                    output_msg = cc.build_output(rt, msg, x, last_layer, final)
                    rt.emit_result(output_msg)
                    cc.finish_input(rt, msg, ns)

                # schedule prefetch of the next local window, or wrap to the first window so the
                # next token's first copies overlap the other shards' compute
                if self._mode == "offload":
                    next_window = self._next_local_layers(rt._assigned_sorted, last_layer, self.window_size)
                    if not next_window:
                        next_window = rt._assigned_sorted[: self.window_size]
                    self._schedule_prefetch(msg.nonce, next_window)
                return
        except Exception as e:
            logger.exception("Error in offload policy process: %s", e)
            try:
                if rt.input_pool:
                    rt.input_pool.release(msg.pool_id)
            except Exception:
                pass
        finally:
            try:
                rt._compute_busy.clear()
            except Exception:
                pass

    def clear(self):
        for _, fut in self._prepared_by_nonce.values():
            try:
                if fut and not fut.done():
                    fut.cancel()
            except Exception:
                pass
        self._prepared_by_nonce.clear()
        try:
            if self.weight_cache:
                self.weight_cache.shutdown()
        except Exception:
            pass
        self._bound_versions.clear()
        self._recent_windows.clear()
