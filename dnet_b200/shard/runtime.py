"""ShardRuntime: owns model, per-nonce KV, pools, policy and the local ingress -> compute ->
egress queues (reference src/dnet/shard/runtime.py:56-401), rebuilt over CUDA streams.

No ring, no gRPC, no discovery: submit(ActivationMessage) -> ActivationMessage.
Threading model preserved: one compute thread drains ``activation_recv_queue`` and calls
``policy.process``; a 4-thread executor serves deserialisation / prefetch helpers.
"""
from __future__ import annotations

import asyncio
import gc
import queue
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from queue import Queue
from typing import Any, Dict, List, Optional

import torch

from dnet_b200 import _cabi
from dnet_b200.config import get_settings
from dnet_b200.core.memory.memory_pool import LayerAwareMemoryPool
from dnet_b200.core.models import BaseRingModel, KVHandle, get_ring_model
from dnet_b200.core.types.messages import ActivationMessage
from dnet_b200.utils.logger import logger
from dnet_b200.utils.model import ModelMetadata, SyntheticSource, get_model_metadata, load_weight
from .models import ShardLoadModelRequest, ShardUnloadModelResponse
from .policies import ComputePolicy, NoopPolicy, PolicyPlan, make_policy, plan_policy


class RuntimeKVCacheConfig:
    def __init__(self, settings):
        self.mode: str = settings.mode
        self.bits: int = settings.bits
        self.group_size: int = settings.group_size
        self.kv_ttl_s: float = settings.ttl_s
        self.max_tokens: int = settings.max_tokens


class RuntimeComputeConfig:
    def __init__(self, settings):
        self.prefetch_mode: str = settings.prefetch_mode
        self.mxload_fastpath: bool = settings.mxload_fastpath
        self.input_pool_mb: int = settings.input_pool_mb
        self.output_pool_mb: int = settings.output_pool_mb


class NonceState:
    """Everything a nonce owns on this shard: paged KV handle, its HBM activation buffer,
    device token-id staging, the pinned (token, logprob) result and captured step graphs."""

    def __init__(self, model: BaseRingModel, max_tokens: int):
        self.kv = KVHandle(model, max_tokens)
        self._model = model
        self._x: Optional[torch.Tensor] = None
        self._ids: Optional[torch.Tensor] = None
        self.graphs: Dict[Any, int] = {}
        self.keepalive = None
        self.lane: int = -1                       # hop lane of the nonce (device-hop transport), -1 = none
        self.params: Dict[str, Any] = {}          # request parameters the token tap needs (callback url, logprobs)
        self.hop_sent_event = None                # comm-stream event: the activation buffer's last hop copy has run
        self.input_copy_event = None              # event behind the H2D copy out of the pinned input buffer
        res = torch.zeros(2, dtype=torch.int32).pin_memory()
        self.result_i32 = res
        self.result_f32 = res.view(torch.float32)
        self.result_np_i32 = res.numpy()                      # same pinned words, read without tensor indexing overhead
        self.result_np_f32 = self.result_np_i32.view("float32")
        self.result_token_ptr = res.data_ptr()
        self.result_logprob_ptr = res.data_ptr() + 4
        self.x1 = torch.empty(1, model.hidden_size, dtype=torch.bfloat16, device="cuda")  # stable graph address
        # (torch.empty: no fill kernel on the default stream racing the compute stream)

    def x_view(self, T: int) -> torch.Tensor:
        if T == 1:
            return self.x1
        if self._x is None or self._x.shape[0] < T:
            self._x = torch.empty(T, self._model.hidden_size, dtype=torch.bfloat16, device="cuda")
        return self._x[:T]

    def ids_view(self, T: int) -> torch.Tensor:
        if self._ids is None or self._ids.numel() < T:
            self._ids = torch.empty(max(T, 16), dtype=torch.int32, device="cuda")
        return self._ids[:T]

    def drop_graphs(self) -> None:
        lib = _cabi.load()
        for g in self.graphs.values():
            lib.dn_graph_destroy(g)
        self.graphs.clear()

    def free(self) -> None:
        self.drop_graphs()
        self.kv.free()


class ShardRuntime:
    """Topology-agnostic shard runtime."""

    def __init__(self, shard_id, queue_size: int = 128, device_prefetch_workers: int = 4, prefetch_threads: int = 2):
        self.shard_id = shard_id
        settings = get_settings()
        self._compute_settings = settings.compute
        self._transport_settings = settings.transport
        self._topology_settings = settings.topology
        self.kv_cache_config = RuntimeKVCacheConfig(settings.kv_cache)
        self._compute_config = RuntimeComputeConfig(settings.compute)
        self.policy: ComputePolicy = NoopPolicy(runtime=self, resident_windows=1)
        self._device_prefetch_workers = device_prefetch_workers
        self.prefetch_threads = prefetch_threads
        self._loop: Optional[asyncio.AbstractEventLoop] = None
        self.max_queue_size = queue_size
        self.activation_recv_queue: Queue[ActivationMessage] = Queue(maxsize=queue_size)
        self.activation_send_queue: Queue[ActivationMessage] = Queue(maxsize=queue_size)
        self.compute_thread: Optional[threading.Thread] = None
        self.running = False
        # CUDA device of this shard = the creating thread's current device; worker threads bind to it
        self._device_index: Optional[int] = int(torch.cuda.current_device()) if torch.cuda.is_available() else None
        self.executor = ThreadPoolExecutor(max_workers=int(self._device_prefetch_workers or 4),
                                           initializer=self._bind_thread_device)
        self.assigned_layers: List[int] = []
        self._assigned_sorted: List[int] = []
        self._assigned_set: set = set()
        self.model_metadata: Optional[ModelMetadata] = None
        self.model: Optional[BaseRingModel] = None
        self.cache: Optional[Any] = None
        self.model_path: Optional[Any] = None
        self.input_pool: Optional[LayerAwareMemoryPool] = None
        self.output_pool: Optional[LayerAwareMemoryPool] = None
        _wd = (self._transport_settings.wire_dtype or "fp16").strip().lower()
        self._wire_dtype_str = "bfloat16" if _wd in {"bf16", "bfloat16"} else "float16"
        self._wire_mx_dtype = torch.bfloat16 if self._wire_dtype_str == "bfloat16" else torch.float16
        self._compute_busy = threading.Event()
        self._mlx_lock = threading.Lock()  # name kept for drop-in parity; guards C-ABI calls
        self._model_lock = threading.Lock()
        self._kv_by_nonce: Dict[str, NonceState] = {}
        self._ns_pool: List[NonceState] = []     # recycled nonce states: creating one costs cudaMalloc + pinned alloc
        self._kv_last_seen: Dict[str, float] = {}
        self._kv_ttl_s: float = self.kv_cache_config.kv_ttl_s
        # CUDA plumbing
        self.compute_stream: Optional[torch.cuda.Stream] = None
        self.compute_stream_ptr: int = 0
        self.use_cuda_graphs: bool = bool(self._compute_settings.cuda_graphs)
        self.use_megakernel: bool = bool(self._compute_settings.megakernel)
        self.stage_host: bool = True
        self._api_tensors: Dict[str, torch.Tensor] = {}
        # device-hop transport state, owned by the topology adapter (shard/adapters/ring.py)
        self.comm_stream: Optional[torch.cuda.Stream] = None   # second stream: ring hops overlap the next nonce's compute
        self.hop = None                      # HopLink: this shard's receive lanes + the successor's
        self.hop_pending = None              # HopLink being set up (answers the predecessor's endpoint request)
        self.on_emit = None                  # adapter hook run by emit_result on the compute thread
        self.token_tap = None                # TokenTap on the finalising shard
        self.lane_nonce: Dict[int, str] = {}
        self._deferred_releases: List[Any] = []   # (pool_id, event) input buffers whose H2D copy is still queued
        self.step_errors = 0

    @property
    def compute_config(self):
        return self._compute_config

    @property
    def transport_config(self):
        return self._transport_settings

    @property
    def topology_config(self):
        return self._topology_settings

    def attach_loop(self, loop):
        self._loop = loop

    def queue_size(self) -> int:
        return self.activation_recv_queue.qsize()

    # -- C ABI shorthands used by the adapter hook --------------------------------------------
    @property
    def lib(self):
        return _cabi.load()

    @staticmethod
    def check(rc: int) -> int:
        return _cabi.check(rc)

    def _bind_thread_device(self) -> None:
        """The CUDA current device is per thread and defaults to 0: every worker thread (compute,
        executor / prefetch) binds to the shard's device before touching CUDA."""
        dev = self._device_index
        if dev is not None and torch.cuda.is_available():
            torch.cuda.set_device(dev)

    def emit_result(self, msg: ActivationMessage) -> None:
        hook = self.on_emit
        if hook is not None:
            try:
                hook(msg)      # device hop of the tensor / first token, before the frame is queued
            except Exception as e:
                logger.error("egress hook failed for nonce %s (%s); the frame falls back to bytes", msg.nonce, e)
        self.activation_send_queue.put_nowait(msg)

    # -- input buffers whose host->device copy is still queued ------------------------------------
    def release_input(self, pool_id: int, event=None) -> None:
        """input_pool.release that respects an in-flight cudaMemcpyAsync out of the pinned buffer: with an
        event the buffer only becomes FREE (reusable by codec.deserialize) once the copy has run."""
        if self.input_pool is None or pool_id is None or pool_id < 0:
            return
        if event is not None:
            self._deferred_releases.append((pool_id, event))
        else:
            self.input_pool.release(pool_id)
        self.reap_releases()

    def reap_releases(self, wait: bool = False) -> None:
        if not self._deferred_releases:
            return
        lib = _cabi.load()
        keep = []
        for pid, ev in self._deferred_releases:
            if wait:
                lib.dn_event_sync(ev)
            if wait or lib.dn_event_query(ev) == 1:
                lib.dn_event_destroy(ev)
                if self.input_pool is not None:
                    self.input_pool.release(pid)
            else:
                keep.append((pid, ev))
        self._deferred_releases = keep

    def all_nonce_states(self):
        return list(self._kv_by_nonce.values())

    def shutdown(self) -> None:
        self.running = False
        if self.compute_thread:
            try:
                self.compute_thread.join(timeout=5)
            except Exception:
                pass
            self.compute_thread = None
        self.executor.shutdown(wait=True, cancel_futures=True)
        for ns in list(self._kv_by_nonce.values()) + self._ns_pool:
            ns.free()
        self._kv_by_nonce.clear()
        self._ns_pool.clear()
        self._kv_last_seen.clear()

    # -- model load ------------------------------------------------------------------------
    def load_model_core(self, req: ShardLoadModelRequest) -> None:
        if not torch.cuda.is_available():
            raise RuntimeError("dnet_b200 needs a CUDA device: the shard forward has no CPU fallback")
        self._bind_thread_device()           # load may run on an executor thread (Shard.load_model)
        self._device_index = int(torch.cuda.current_device())
        _cabi.init(self._device_index)
        if self._wire_dtype_str != "bfloat16":
            raise ValueError("set DNET_TRANSPORT_WIRE_DTYPE=bf16: the wire dtype must equal the bf16 model dtype")
        lib = _cabi.load()
        lib.dn_set_option(b"pdl", 1 if self._compute_settings.pdl else 0)
        self.model_metadata = get_model_metadata(req.model_path)
        self.assigned_layers = list(req.layers)
        self._assigned_sorted = sorted(self.assigned_layers)
        self._assigned_set = set(self._assigned_sorted)
        self.model_path = req.model_path
        local_count = max(1, len(self.assigned_layers))
        plan: PolicyPlan = plan_policy(local_count=local_count, requested_w=int(req.window_size),
                                       residency_size=int(req.residency_size), topology_config=self._topology_settings)
        # kv_bits "4bit" / "8bit" (the reference API's defaults, api/models.py:316,342) select the affine group-64
        # quantised cache (reference runtime.py:204-214 -> utils/model.py:505-554); "fp16" the 16-bit cache
        kv = (req.kv_bits or "").strip().lower()
        kv_bits = {"4bit": 4, "8bit": 8}.get(kv, 0)
        self.kv_cache_config.mode = kv if kv_bits else "fp16"
        self.kv_cache_config.bits = kv_bits or self.kv_cache_config.bits
        if self.compute_stream is None:
            self.compute_stream = torch.cuda.Stream()
            self.compute_stream_ptr = int(self.compute_stream.cuda_stream)
            self.comm_stream = torch.cuda.Stream()
        # fit mode with synthetic weights generates straight into HBM; everything else stages
        # the packed layer records in pinned host memory first
        self.stage_host = not (isinstance(self.model_metadata.source, SyntheticSource) and plan.mode == "fit")
        self.policy = make_policy(plan.mode, self, plan.resident_windows)
        self.policy.window_size = plan.window_size
        self.policy.configure_policy_for_model(req)
        self.input_pool = LayerAwareMemoryPool(total_memory_mb=int(self._compute_settings.input_pool_mb), placement="pinned")
        self.output_pool = LayerAwareMemoryPool(total_memory_mb=int(self._compute_settings.output_pool_mb), placement="pinned")
        settings = get_settings()
        pages = int(settings.kv_cache.pool_pages) or max(1, (self.kv_cache_config.max_tokens + 63) // 64) * 8
        self.model = get_ring_model(self.model_metadata.model_type, self.model_metadata.model_config,
                                    assigned_layers=self.assigned_layers, is_api_layer=False,
                                    kv_pool_pages=pages, wire_dtype=self._wire_dtype_str, kv_bits=kv_bits,
                                    kv_group=int(self.kv_cache_config.group_size))
        self.model.apply_quantization_from_config(self.model_metadata.model_config, model_metadata=self.model_metadata)
        if not getattr(self.model, "step_kernel_ok", True):
            # e.g. sparse MoE: expert selection happens on the device, the layers run on the per-op path (CUDA graphs)
            self.use_megakernel = False
            self.use_cuda_graphs = True
        # embed / norm / head iff this shard owns layer 0 / the last layer (reference runtime.py:263-273)
        has_start = 0 in self.assigned_layers
        has_end = (self.model_metadata.num_layers - 1) in self.assigned_layers
        tied = bool(self.model_metadata.model_config.get("tie_word_embeddings", False))
        api: Dict[str, torch.Tensor] = {}

        _dev = self._api_tensor_to_device

        if has_start or (has_end and tied):
            api["embed_tokens.weight"] = _dev(self.model_metadata.embed_tokens["weight"])
        if has_end:
            api["norm.weight"] = _dev(self.model_metadata.norm["weight"])
            if not tied:
                w = _dev(self.model_metadata.lm_head["weight"])
                if w.shape[0] != self.model.vocab_size and w.shape[1] == self.model.vocab_size:
                    w = w.t().contiguous()  # transposed (hidden, vocab) head accepted (utils/model.py:341-346)
                api["lm_head.weight"] = w
        self._api_tensors = api
        if api:
            self.model.load_weights(list(api.items()), strict=False)
        torch.cuda.synchronize()   # loads ran on torch's current stream; compute runs on compute_stream

    def _api_tensor_to_device(self, info) -> torch.Tensor:
        src = self.model_metadata.source
        if isinstance(src, SyntheticSource):
            t = torch.empty(tuple(info.shape), dtype=torch.bfloat16, device="cuda")
            src.fill_device(info, t)
            return t
        return load_weight(info, {}, src).to("cuda", non_blocking=False).contiguous()

    def load_head_slice(self, position: int, ring_size: int) -> tuple:
        """Tensor-parallel lm_head over the ring: bind this shard's vocabulary rows
        [V*position/S, V*(position+1)/S) and the final norm (every shard applies it to the hidden state the
        last shard broadcasts).  The shard that owns the last layer keeps the full head (prefill samples with
        it) and its slice is a view into it; the others load only their rows."""
        self._bind_thread_device()
        md = self.model_metadata
        V = int(self.model.vocab_size)
        row0, row1 = V * position // ring_size, V * (position + 1) // ring_size
        tied = bool(md.model_config.get("tie_word_embeddings", False))
        with self._model_lock:
            api = self._api_tensors
            full = api.get("lm_head.weight") if not tied else api.get("embed_tokens.weight")
            if "norm.weight" not in api:
                api["norm.weight"] = self._api_tensor_to_device(md.norm["weight"])
            if full is None:
                info = (md.embed_tokens if tied else md.lm_head)["weight"]
                whole = self._api_tensor_to_device(info)          # temporary: only the slice is kept
                if whole.shape[0] != V and whole.shape[1] == V:
                    whole = whole.t().contiguous()
                sl = whole[row0:row1].clone()
                del whole
                torch.cuda.empty_cache()
            else:
                sl = full[row0:row1]
            api["_head_slice"] = sl
            self.model.load_weights([(k, v) for k, v in api.items() if not k.startswith("_")], strict=False)
            _cabi.check(_cabi.load().dn_bind_head_slice(self.model._h, sl.data_ptr(), row0, row1 - row0))
            torch.cuda.synchronize()
        return row0, row1

    def unload_model_core(self) -> ShardUnloadModelResponse:
        try:
            with self._model_lock:
                if self.model is None:
                    return ShardUnloadModelResponse(success=True, message="No model loaded")
                while not self.activation_recv_queue.empty():
                    try:
                        self.activation_recv_queue.get_nowait()
                    except queue.Empty:
                        break
                if self.compute_stream is not None:
                    self.compute_stream.synchronize()
                self.reap_releases(wait=True)
                for ns in list(self._kv_by_nonce.values()) + self._ns_pool:
                    ns.free()
                self._kv_by_nonce.clear()
                self._ns_pool.clear()
                self._kv_last_seen.clear()
                self.lane_nonce.clear()
                self.policy.clear()
                self.policy = NoopPolicy(runtime=self, resident_windows=1)
                self.model.destroy()
                self.model = None
                self.cache = None
                self.model_metadata = None
                self.assigned_layers = []
                self.model_path = None
                self._assigned_sorted = []
                self._assigned_set = set()
                self.input_pool = None
                self.output_pool = None
                self._api_tensors = {}
                gc.collect()
                torch.cuda.empty_cache()
            return ShardUnloadModelResponse(success=True, message="Model unloaded successfully")
        except Exception as e:
            logger.exception("Node %s: Error unloading model: %s", self.shard_id, e)
            return ShardUnloadModelResponse(success=False, message=f"Error unloading model: {str(e)}")

    def reset_cache(self):
        if not self.model:
            logger.warning("Node %s: Cannot reset cache - no model loaded", self.shard_id)
            return
        self.cache = None

    def compute(self, activation_msg: ActivationMessage) -> None:
        if not self.policy:
            logger.error("Runtime %s: no compute policy configured", self.shard_id)
            return
        self.policy.process(activation_msg)

    def _compute_worker(self) -> None:
        self._bind_thread_device()
        while self.running:
            try:
                activation_msg = self.activation_recv_queue.get(timeout=1.0)
                self.compute(activation_msg)
            except queue.Empty:
                continue
            except Exception as e:
                logger.error("Compute worker error: %s", e)

    def get_or_make_kv(self, nonce: str) -> NonceState:
        """Per-nonce KV for this shard's local layers, with the reference's TTL sweep
        (runtime.py:374-396)."""
        if not self.model:
            raise RuntimeError("Model not initialized")
        now = time.perf_counter()
        ttl = float(self._kv_ttl_s)
        for n, ts in list(self._kv_last_seen.items()):
            if (now - ts) > ttl and n != nonce:
                self._kv_last_seen.pop(n, None)
                old = self._kv_by_nonce.pop(n, None)
                if old is not None:
                    self._recycle(old)
        ns = self._kv_by_nonce.get(nonce)
        if ns is None:
            if self._ns_pool:
                ns = self._ns_pool.pop()
                ns.kv.reset(self.compute_stream_ptr)
            else:
                ns = NonceState(self.model, self.kv_cache_config.max_tokens)
            self._kv_by_nonce[nonce] = ns
        self._kv_last_seen[nonce] = time.perf_counter()
        return ns

    def note_new_nonce(self, nonce: str) -> None:
        """Transport saw the first frame of a nonce (reference adapters/ring.py:177-181 builds the KV right
        there; here the compute thread does, in policy.process, so that all CUDA state of a shard is created
        by the one thread bound to its device)."""
        self._kv_last_seen.setdefault(nonce, time.perf_counter())

    def _recycle(self, ns: NonceState) -> None:
        """An expired nonce's KV pages, buffers and pinned result are kept for the next nonce."""
        if ns.lane >= 0:
            for lane, n in list(self.lane_nonce.items()):
                if lane == ns.lane and self._kv_by_nonce.get(n) is None:
                    self.lane_nonce.pop(lane, None)
        ns.lane, ns.params, ns.hop_sent_event = -1, {}, None
        if ns.kv.max_tokens >= self.kv_cache_config.max_tokens and len(self._ns_pool) < 64:
            ns.drop_graphs()
            self._ns_pool.append(ns)
        else:
            if self.compute_stream is not None:
                self.compute_stream.synchronize()
            ns.free()

    def release_nonce(self, nonce: str) -> None:
        """Explicit end-of-request (the reference only has the TTL sweep)."""
        ns = self._kv_by_nonce.pop(nonce, None)
        self._kv_last_seen.pop(nonce, None)
        if ns is not None:
            self._recycle(ns)

    def release_nonce_deferred(self, nonce: str) -> None:
        """End of request from the transport: kernels of the nonce may still be queued on the compute
        stream, so its KV only goes back to the pool through the TTL sweep (a recycled NonceState is
        reset on the same stream, i.e. behind those kernels)."""
        self._kv_last_seen[nonce] = time.perf_counter() - float(self._kv_ttl_s) - 1.0

    def start(self):
        self.running = True
        self.compute_thread = threading.Thread(target=self._compute_worker, daemon=True)
        self.compute_thread.start()
